"""Global top-k region selection and the pool/label bookkeeping around it.

Reference behaviour being reproduced (file:line relative to the reference checkout):
  * ``sorted(scores, reverse=True)`` over ``(score, 'img,lbl,spx', id)`` tuples -- active_selection/base.py:37
  * ``RegionActiveDataset.expand_training_set`` walking that list until the cumulative cost exceeds the
    budget -- dataloader/region_active_dataset.py:16-73
Only the first ``budget + 1`` entries of the 6 M-entry sorted list can ever be consumed (every region
costs >= 1), so the device selects and sorts exactly those (``ops.topk_keys``) and the host never sees
the rest.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import dist as mdist
from . import ops


def image_ranks(im_idx: Sequence[Sequence[str]]) -> np.ndarray:
    """rank[i] = position of image i's joined path in ascending string order (ties in the tuple sort)."""
    joined = [",".join(k) for k in im_idx]
    order = sorted(range(len(joined)), key=joined.__getitem__)
    rank = np.empty(len(joined), dtype=np.int32)
    rank[np.asarray(order, dtype=np.int64)] = np.arange(len(joined), dtype=np.int32)
    return rank


def pool_mask(im_idx: Sequence[Sequence[str]], suppix: Dict[str, List[int]], nseg: int,
              lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
    """(hi-lo, nseg) uint8: 1 where the region is still in the unlabeled pool (my_bvsb.py:44-46)."""
    hi = len(im_idx) if hi is None else hi
    mask = np.zeros((hi - lo, nseg), dtype=np.uint8)
    for row, key in enumerate(im_idx[lo:hi]):
        ids = suppix.get(key[2])
        if ids:
            mask[row, np.asarray(ids, dtype=np.int64)] = 1
    return mask


def decode_keys(keys: np.ndarray, nseg: int, im_idx: Sequence[Sequence[str]], rank: np.ndarray):
    """uint64 keys -> [(score, 'img,lbl,spx', id)] in the given order (see ``mas_region_keys_dev``)."""
    keys = np.ascontiguousarray(keys).view(np.uint64)
    hi = (keys >> np.uint64(32)).astype(np.uint32)
    tie = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    bits = np.where(hi & np.uint32(0x80000000), hi & np.uint32(0x7FFFFFFF), ~hi).astype(np.uint32)
    scores = bits.view(np.float32).astype(np.float64).tolist()
    image_of_rank = np.empty(len(rank), dtype=np.int64)
    image_of_rank[rank] = np.arange(len(rank))
    imgs = image_of_rank[tie // nseg].tolist()
    ids = (tie % nseg).tolist()
    joined = {}
    out = []
    for s, i, r in zip(scores, imgs, ids):
        path = joined.get(i)
        if path is None:
            path = joined[i] = ",".join(im_idx[i])
        out.append((s, path, int(r)))
    return out


_pinned = {}


def _pinned_buffers(device, k: int):
    """Pinned host landing zone of a round's result (k keys + [n_take, worst]), allocated once per (device, k)."""
    key = (str(device), int(k))
    buf = _pinned.get(key)
    if buf is None:
        buf = _pinned[key] = (torch.empty(k, dtype=torch.int64).pin_memory(), torch.empty(2, dtype=torch.int32).pin_memory())
    return buf


def _ranked_to_host(best: torch.Tensor, take: torch.Tensor, worst: Optional[torch.Tensor], k: int):
    """ONE host sync per round: the ranked keys (first k slots) and the two control words land in pinned memory through
    two async copies on the current stream, then a single stream synchronize.  -> (keys view, n_take, worst)."""
    h_keys, h_meta = _pinned_buffers(best.device, k)
    meta = torch.stack([take.reshape(()), (worst if worst is not None else take).reshape(())])
    h_meta.copy_(meta, non_blocking=True)
    h_keys.copy_(best[:k], non_blocking=True)
    torch.cuda.current_stream(best.device).synchronize()
    return h_keys, int(h_meta[0]), int(h_meta[1])


def top_regions(scores: torch.Tensor, in_pool: torch.Tensor, image_rank_local: torch.Tensor, k: int, group=None,
                cost_by_tie: Optional[torch.Tensor] = None, budget: Optional[int] = None):
    """Sorted (descending) keys of the k best pool regions over ALL ranks, as a host uint64 array.

    scores / in_pool: this rank's (n_local, S) shard; image_rank_local: global ranks of its images.
    Per rank: key -> candidate superset of its k best with the count in the last slot (one message of
    ``sort_capacity(k) + 1`` slots) -> ONE ``all_gather_into_tensor`` -> counts reduced / cleared on the device ->
    select + sort k on every rank -> one host sync for the result.
    With ``cost_by_tie`` (uint8 device table of label costs indexed by image rank * S + id, see ``region_cost_table``)
    and ``budget`` the list is cut on the device where ``expand_training_set`` would stop (``mas_prefix_cut_dev``) and
    only that prefix is decoded on the host.
    """
    local_keys = ops.region_keys(scores, in_pool, image_rank_local)
    distributed = mdist.is_distributed(group)
    keys, worst = local_keys, None
    if distributed:
        gathered = mdist.gather_messages(ops.topk_candidates_msg(local_keys, k), group)
        worst = torch.empty(1, dtype=torch.int32, device=scores.device)
        ops.merge_counts(gathered, worst)                # -1 if the buffer overflowed on any rank
        keys = gathered.view(-1)
    best, count = ops.topk_sorted(keys, k)               # two bucket histograms + compaction + sort
    take = ops.prefix_cut(best, count, cost_by_tie, budget) if cost_by_tie is not None else count
    k_host = min(int(k), best.numel())
    host, n, bad = _ranked_to_host(best, take, worst, k_host)
    if n < 0 or bad < 0:
        # massively tied scores overflowed a candidate buffer somewhere: every rank redoes the exact radix select
        keys = local_keys
        if distributed:
            local, count = ops.topk_keys(local_keys, k, sort=False)
            keys = mdist.gather_candidates(local, count, k, group)
        best, count = ops.topk_keys(keys, k, sort=True)
        take = ops.prefix_cut(best, count, cost_by_tie, budget) if cost_by_tie is not None else count
        host, n, _ = _ranked_to_host(best, take, None, k_host)
    return host[:n].numpy().copy().view(np.uint64)


def region_cost_table(costs_by_image: np.ndarray, rank: np.ndarray, device) -> torch.Tensor:
    """(N, S) label costs in pool order -> flat uint8 device table indexed by image RANK * S + id (the low key word)."""
    n, s = costs_by_image.shape
    table = np.empty((n, s), dtype=np.uint8)
    table[np.asarray(rank, dtype=np.int64)] = np.minimum(costs_by_image, 255).astype(np.uint8)
    return torch.from_numpy(table.reshape(-1)).to(device)


def cumulative_cut(costs: np.ndarray, budget: int) -> int:
    """Length of the prefix expand_training_set consumes: stop AFTER the pick that makes the running
    cost exceed the budget (strict '>', region_active_dataset.py:66); the whole list if it never does."""
    running = np.cumsum(costs.astype(np.int64))
    over = np.nonzero(running > budget)[0]
    return int(over[0]) + 1 if over.size else int(len(costs))
