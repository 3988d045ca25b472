"""Multi-GPU plumbing: shard the pool by image, exchange only the small pool-wide quantities.

One process per GPU (``torch.distributed``, NCCL over NVLink on the GPU box; the same code runs over
``gloo`` on CPU tensors in the tests).  The logits never cross GPUs -- the path shards by image with
no data-path collective; what is exchanged:
  * per-image class-probability sums           (N x C' f64, all_gather)   -> class weights
  * min over non-zero / max of the scores      (2 floats, all_reduce)      -> normalisation
  * dominant-class histogram                   (C' int64, all_reduce)      -> clsbal weights
  * each rank's k best (score, region) keys    (k x 8 B, all_gather)       -> global top-k merge
Every helper is the identity when ``group`` is None and torch.distributed is not initialised.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as td


# pass as ``group`` to keep a call local to this process although torch.distributed is initialised (bench.py recomputes a
# sharded mini-round on ONE rank this way to check the multi-GPU merge against it)
SINGLE = "single-process"


def is_distributed(group=None) -> bool:
    if group is SINGLE:
        return False
    return td.is_available() and td.is_initialized() and td.get_world_size(group) > 1


def rank_world(group=None) -> Tuple[int, int]:
    if group is SINGLE:
        return 0, 1
    if td.is_available() and td.is_initialized():
        return td.get_rank(group), td.get_world_size(group)
    return 0, 1


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image range of ``rank``: the first ``n % world`` ranks get one extra image."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def all_gather_rows(local: torch.Tensor, group=None, counts: Optional[List[int]] = None) -> torch.Tensor:
    """Concatenate per-rank (n_r, ...) tensors along dim 0 in rank order (n_r may differ by rank).
    ``counts`` = the n_r of every rank when the caller knows them (``shard_sizes``): saves a collective and a host sync.
    ONE ``all_gather_into_tensor`` into a single (world, width, ...) buffer; with equal shards the result is a view of it."""
    if not is_distributed(group):
        return local
    world = td.get_world_size(group)
    if counts is None:
        seen = torch.zeros(world, dtype=torch.int64, device=local.device)
        seen[td.get_rank(group)] = local.shape[0]
        td.all_reduce(seen, op=td.ReduceOp.SUM, group=group)
        counts = seen.tolist()
    elif len(counts) != world or counts[td.get_rank(group)] != local.shape[0]:
        raise RuntimeError(f"all_gather_rows: counts {counts} do not match this rank's {local.shape[0]} rows")
    width = max(counts)
    rest = tuple(local.shape[1:])
    if local.shape[0] == width:
        padded = local.contiguous()
    else:
        padded = local.new_zeros((width,) + rest)
        padded[: local.shape[0]] = local
    out = local.new_empty((world * width,) + rest)       # concatenated layout (what gloo accepts as well as NCCL)
    td.all_gather_into_tensor(out, padded, group=group)
    if min(counts) == width:
        return out
    boxes = out.view((world, width) + rest)
    return torch.cat([boxes[r, :c] for r, c in enumerate(counts)], dim=0)


def all_reduce_minmax(minmax: torch.Tensor, group=None) -> torch.Tensor:
    """[min, max] pairs reduced over the ranks (+inf / -inf from an empty shard are neutral): one MAX all-reduce of
    [-min, max]."""
    if not is_distributed(group):
        return minmax
    both = torch.stack([-minmax[0], minmax[1]])
    td.all_reduce(both, op=td.ReduceOp.MAX, group=group)
    return torch.stack([-both[0], both[1]]).contiguous()


def all_reduce_min(t: torch.Tensor, group=None) -> torch.Tensor:
    if is_distributed(group):
        td.all_reduce(t, op=td.ReduceOp.MIN, group=group)
    return t


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    if is_distributed(group):
        td.all_reduce(t, op=td.ReduceOp.SUM, group=group)
    return t


def gather_messages(msg: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank's fixed-size message -> one (world, len(msg)) tensor on every rank: a single
    ``all_gather_into_tensor`` into one buffer (no per-rank list, no concatenation).  The top-k merge sends each rank's
    candidate keys with their count in the last slot this way (``mas_topk_candidates_msg_u64_dev``)."""
    if not is_distributed(group):
        return msg.view(1, -1)
    world = td.get_world_size(group)
    out = msg.new_empty(world * msg.numel())
    td.all_gather_into_tensor(out, msg.contiguous().view(-1), group=group)
    return out.view(world, msg.numel())


def gather_candidates(keys: torch.Tensor, count: torch.Tensor, k: int, group=None) -> torch.Tensor:
    """All ranks' candidate keys -> one flat int64 tensor on every rank (unused slots are 0 = 'no key').

    ``keys`` holds this rank's candidates in its first ``count`` slots, ``count`` <= k (anything after is ignored).
    Used by the exact fallback of the merge (``selection.top_regions``).
    """
    local = keys[:k].clone()
    slot = torch.arange(k, device=keys.device)
    local = torch.where(slot < count.to(torch.int64), local, torch.zeros_like(local))
    return gather_messages(local, group).reshape(-1)
