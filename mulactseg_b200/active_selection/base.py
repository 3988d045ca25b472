"""Drop-in ``RegionSelector`` base: same constructor, attributes and entry points as the reference's
``active_selection/base.py:13-38``, with the scoring and ranking done by the sm_100a kernels.

Plugin contract kept from the reference (``train_AL.py:29-32,65-68``):
  ``RegionSelector(args)`` reads ``val_batch_size, val_num_workers, nseg, active_method, num_classes``
  (+ ``ce_temp, cls_weight_coeff, method, save_scores``);
  ``calculate_scores(trainer, pool_set) -> [(score, 'img,lbl,spx', id)]`` for every region still in the pool;
  ``select_next_batch(trainer, active_set, selection_count)`` ranks and calls
  ``active_set.expand_training_set(sorted_list, selection_count, active_method)``.
``trainer`` supplies ``.net`` / ``.device`` / ``.model_save_dir`` / ``.selection_iter``; ``pool_set`` supplies
``.im_idx``, ``.suppix`` and items ``{'images', 'spx', ...}``.

Differences that are invisible to the caller: the pool is scored in ONE pass over the logits (the
reference's ``predclsbal`` selectors run the network twice), per-region results stay on the GPU, and
``select_next_batch`` hands ``expand_training_set`` only the prefix of the sorted list it can consume.
With ``torch.distributed`` initialised the pool is sharded by image across the ranks.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from .. import acquisition, dist as mdist, selection


def _collate(items):
    out = {}
    for key in ("images", "spx"):
        vals = [it[key] for it in items]
        vals = [torch.from_numpy(v) if isinstance(v, np.ndarray) else v for v in vals]
        out[key] = torch.stack(vals)
    return out


class PoolScores:
    """Scores of this rank's shard (device) + what is needed to rank them globally."""

    def __init__(self, scores, dominant, lo, hi, n_total, nseg):
        self.scores, self.dominant = scores, dominant
        self.lo, self.hi, self.n_total, self.nseg = lo, hi, n_total, nseg


class RegionSelector(object):
    method_name = None  # set by the concrete modules (== the reference module name)

    def __init__(self, args):
        self.args = args
        self.batch_size = args.val_batch_size
        self.num_workers = args.val_num_workers
        self.num_superpixels = args.nseg
        self.active_method = args.active_method
        self.num_class = args.num_classes
        self.eps = 1e-8
        self.temperature = getattr(args, "ce_temp", 1.0)
        name = self.method_name or self.active_method
        if name not in acquisition.SELECTORS:
            raise NotImplementedError(f"no B200 selector for active_method={name!r}")
        self.spec = acquisition.SELECTORS[name]
        self.group = None  # torch.distributed process group (None = default group / single process)

    # ------------------------------------------------------------------ scoring (device)
    def score_regions(self, trainer, pool_set) -> PoolScores:
        model = trainer.net
        model.eval()
        device = torch.device(trainer.device)
        if device.type != "cuda":
            raise RuntimeError("mulactseg_b200 selectors need a CUDA device (there is no CPU path)")
        rank, world = mdist.rank_world(self.group)
        n_total = len(pool_set)
        lo, hi = mdist.shard_range(n_total, rank, world)
        shard = torch.utils.data.Subset(pool_set, range(lo, hi)) if world > 1 else pool_set
        loader = torch.utils.data.DataLoader(dataset=shard, batch_size=self.batch_size, shuffle=False,
                                             num_workers=self.num_workers, collate_fn=_collate, pin_memory=True)
        predignore = "predignore" in getattr(self.args, "method", "")
        if self.spec.ban_ignore:
            assert predignore  # my_bvsb_banignore.py:35
        # opt-in: a net that exposes ``forward_lowres(images) -> (B, C', h, w)`` -- the head's output BEFORE
        # the final x4 ``F.interpolate`` of models/segmentation/utils.py:28-34 -- is scored from that tensor, the
        # interpolation to the id map's size happening inside the kernel (``--b200_lowres`` / args.b200_lowres)
        lowres = bool(getattr(self.args, "b200_lowres", False)) and hasattr(model, "forward_lowres")
        stats, first = None, 0
        with torch.no_grad():
            for batch in loader:
                images = batch["images"].to(device, dtype=torch.float32, non_blocking=True)
                spx = batch["spx"].to(torch.int32).to(device, non_blocking=True)
                preds = model.forward_lowres(images) if lowres else model(images)      # (B, C', H, W) -- stays PyTorch
                if preds.dtype not in (torch.float32, torch.bfloat16):
                    preds = preds.float()                  # fp16 / fp64 heads: the reference's ops take any float dtype
                b_, c_, h_, w_ = preds.shape
                if preds.stride(3) != 1 or preds.stride(2) != w_ or preds.stride(1) != h_ * w_:
                    preds = preds.contiguous()             # channels_last / sliced outputs: one copy, like .softmax would make
                if self.spec.slice_ignore and predignore:
                    preds = preds[:, :-1]                  # read in place through the image stride
                if stats is None:
                    stats = acquisition.RegionStats(hi - lo, self.num_superpixels, preds.shape[1], device,
                                                    need_prob=self.spec.weighting == "predclsbal")
                if lowres:
                    stats.add_batch_lowres(first, preds.contiguous(), spx, self.temperature)
                else:
                    stats.add_batch(first, preds, spx, self.temperature)
                first += preds.shape[0]
        if stats is None:
            raise RuntimeError("empty pool shard: fewer pool images than ranks")
        scores, dominant = acquisition.finalize(stats, self.spec, getattr(self.args, "cls_weight_coeff", 0.0),
                                                self.batch_size, self.group, mdist.shard_sizes(n_total, world))
        return PoolScores(scores, dominant, lo, hi, n_total, self.num_superpixels)

    # ------------------------------------------------------------------ reference-compatible entry points
    def gen_score_list_from_tensor(self, pool_set, scores_tensor):
        """my_bvsb.py:29-48, vectorised per image."""
        out = []
        host = scores_tensor.detach().to("cpu", torch.float32).numpy()
        for k, key in enumerate(pool_set.im_idx):
            ids = pool_set.suppix[key[2]]
            path = ",".join(key)
            vals = host[k][np.asarray(ids, dtype=np.int64)].astype(np.float64).tolist() if len(ids) else []
            out.extend(zip(vals, [path] * len(ids), ids))
        return out

    def calculate_scores(self, trainer, pool_set):
        """Full ``(score, path, id)`` list like the reference (slow host loop; ``select_next_batch`` avoids it)."""
        ps = self.score_regions(trainer, pool_set)
        scores = mdist.all_gather_rows(ps.scores, self.group)
        return self.gen_score_list_from_tensor(pool_set, scores)

    def ranked_prefix(self, ps: PoolScores, pool_set, k: int, cost_table=None, budget=None):
        """The first ``k`` entries of ``sorted(calculate_scores(...), reverse=True)`` without building the list; with a
        device cost table (``selection.region_cost_table``) and a budget, cut where ``expand_training_set`` stops."""
        device = ps.scores.device
        rank = selection.image_ranks(pool_set.im_idx)
        mask = selection.pool_mask(pool_set.im_idx, pool_set.suppix, ps.nseg, ps.lo, ps.hi)
        keys = selection.top_regions(ps.scores, torch.from_numpy(mask).to(device),
                                     torch.from_numpy(rank[ps.lo:ps.hi].copy()).to(device), k, self.group, cost_table, budget)
        return selection.decode_keys(keys, ps.nseg, pool_set.im_idx, rank)

    def _pool_costs(self, active_set, pool_set):
        """(N, S) multi-hot class count of every pool region when --fair_counting --or_labeling
        (region_active_dataset.py:56-61: ``multi_hot_cls[trg_index, suppix_id].sum()``), in pool order."""
        lab = active_set.trg_label_dataset
        rows = [lab.id_to_index[key[2].split("/")[-1].split(".")[0]] for key in pool_set.im_idx]
        return np.asarray(lab.multi_hot_cls)[np.asarray(rows, dtype=np.int64)].sum(axis=-1)

    def select_next_batch(self, trainer, active_set, selection_count):
        pool_set = active_set.trg_pool_dataset
        ps = self.score_regions(trainer, pool_set)

        if getattr(self.args, "save_scores", False):   # base.py:31-34 (needs the full list)
            gathered = mdist.all_gather_rows(ps.scores, self.group)       # collective: every rank takes part ...
            if mdist.rank_world(self.group)[0] == 0:                      # ... one rank writes
                full = self.gen_score_list_from_tensor(pool_set, gathered)
                fname = os.path.join(trainer.model_save_dir, "AL_record", "region_val_{}.json".format(trainer.selection_iter))
                with open(fname, "w") as f:
                    json.dump(full, f)

        n_pool = sum(len(v) for v in pool_set.suppix.values())
        fair = getattr(active_set.args, "fair_counting", False) and getattr(active_set.args, "or_labeling", False)
        k = min(int(selection_count) + 1, n_pool)
        cost_table = costs = None
        if fair:
            # label cost of every pool region as a device table: the ranked list is cut on the GPU where the walk of
            # expand_training_set stops, and only that prefix is decoded into python tuples
            costs = self._pool_costs(active_set, pool_set)
            cost_table = selection.region_cost_table(costs, selection.image_ranks(pool_set.im_idx), ps.scores.device)
        while True:
            prefix = self.ranked_prefix(ps, pool_set, k, cost_table, selection_count if fair else None)
            if k >= n_pool or not fair or len(prefix) < k:
                break
            index_of = {",".join(key): i for i, key in enumerate(pool_set.im_idx)}
            if sum(int(costs[index_of[joined], sid]) for _, joined, sid in prefix) > selection_count:
                break
            k = min(2 * k, n_pool)   # zero-cost regions: the walk needs a longer prefix
        active_set.expand_training_set(prefix, selection_count, self.active_method)
