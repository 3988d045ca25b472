"""Drop-in for the reference's ``dataloader/region_active_dataset.py:8-103`` (``RegionActiveDataset``): the same
pool -> label bookkeeping, ``*_selection_XX.pkl`` / ``datalist_XX.pkl`` files and wandb record, without the
per-pick list scans.

The reference walks the ranked list and per pick does ``list.__contains__`` over the labelled image list
(<= 2975 three-string lists), ``list.remove`` over the image's pool ids (<= 2048 ints) and, when an image runs
empty, ``list.remove`` over the pool image list -- O(n) each, 10-20 s for a 100 000-unit Cityscapes round
(SURVEY.md section 6).  Here the walk only counts costs and groups the picks by image; every container is then
rewritten once, reproducing the reference's element order exactly:
  * label ``im_idx``: images in order of first pick; label ``suppix[spx]``: ids in pick order (a first pick of an
    image that is not yet in ``im_idx`` REPLACES an existing list, like the reference's ``= [suppix_id]``, :37);
  * pool ``suppix[spx]``: remaining ids in their old order; emptied images leave ``suppix`` and ``im_idx``.
Host-side Python by design: the ranked prefix it consumes comes from the GPU top-k.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List


def _is_writer() -> bool:
    """Under torchrun every rank walks the same ranked list (each keeps its own pool / label bookkeeping in step), but
    files and wandb records are written by rank 0 only."""
    try:
        import torch.distributed as td
        return not (td.is_available() and td.is_initialized()) or td.get_rank() == 0
    except Exception:
        return True


class RegionActiveDataset:
    def __init__(self, args, trg_pool_dataset, trg_label_dataset):
        self.args = args
        self.selection_iter = 0
        self.trg_pool_dataset = trg_pool_dataset
        self.trg_label_dataset = trg_label_dataset

    # ------------------------------------------------------------------ selection
    def _region_cost(self):
        if getattr(self.args, "fair_counting", False) and getattr(self.args, "or_labeling", False):
            lab = self.trg_label_dataset
            return lambda spx_path, sid: int(lab.multi_hot_cls[lab.id_to_index[spx_path.split("/")[-1].split(".")[0]], sid].sum())
        return None

    def expand_training_set(self, sample_region, selection_count, selection_method):
        """region_active_dataset.py:16-80.  ``sample_region``: ``[(score, 'img,lbl,spx', id), ...]`` sorted descending."""
        pool, label = self.trg_pool_dataset, self.trg_label_dataset
        cost_of = self._region_cost()
        spent, taken = 0, len(sample_region)
        picks: Dict[str, List[int]] = {}          # spx path -> ids in pick order
        keys: Dict[str, List[str]] = {}           # spx path -> [img, lbl, spx]
        order: List[str] = []                     # spx paths in order of first pick
        for n, (_, joined, sid) in enumerate(sample_region):
            key = joined.split(",")
            spx_path = key[2]
            if spx_path not in picks:
                picks[spx_path] = []
                keys[spx_path] = key
                order.append(spx_path)
            picks[spx_path].append(sid)
            spent += 1 if cost_of is None else cost_of(spx_path, sid)
            if spent > selection_count:           # strict '>', :66
                taken = n + 1
                break
        exceeded = spent > selection_count

        labelled = {tuple(k) for k in label.im_idx}
        emptied = set()
        for spx_path in order:
            key, ids = keys[spx_path], picks[spx_path]
            if tuple(key) not in labelled:
                label.im_idx.append(key)
                labelled.add(tuple(key))
                label.suppix[spx_path] = list(ids)
            else:
                label.suppix[spx_path].extend(ids)
            gone = set(ids)
            if len(gone) != len(ids):
                raise ValueError(f"region picked twice in {spx_path}")     # list.remove would fail too (:42)
            before = pool.suppix[spx_path]
            remaining = [s for s in before if s not in gone]
            if len(before) - len(remaining) != len(ids):
                raise ValueError(f"picked region not in the pool of {spx_path}")
            if remaining:
                pool.suppix[spx_path] = remaining
            else:
                pool.suppix.pop(spx_path)
                emptied.add(tuple(key))
            if hasattr(pool, "isselected"):
                stem = spx_path.split("/")[-1].split(".")[0]
                pool.isselected[label.id_to_index[stem], ids] = 1
        if emptied:
            pool.im_idx[:] = [k for k in pool.im_idx if tuple(k) not in emptied]

        if exceeded and _is_writer():
            fname = f"{selection_method}_selection_{self.selection_iter:02d}.pkl"
            with open(os.path.join(self.args.model_save_dir, fname), "wb") as f:
                pickle.dump(sample_region[:taken], f)
            print(taken)
        wandb = getattr(self.args, "wandb", None) if _is_writer() else None
        if wandb is not None:                      # :75-80
            global_step = int(self.args.finetune_itrs) * (self.selection_iter - 1)
            wandb.log({"num_selected_spx": taken, "num_cls_spx": selection_count / taken, "sampling_iter": self.selection_iter},
                      step=global_step)
        return taken

    # ------------------------------------------------------------------ persistence (identical files)
    def dump_datalist(self):
        if not _is_writer():
            return
        path = os.path.join(self.args.model_save_dir, f"datalist_{self.selection_iter:02d}.pkl")
        with open(path, "wb") as f:
            pickle.dump({"trg_label_im_idx": self.trg_label_dataset.im_idx, "trg_pool_im_idx": self.trg_pool_dataset.im_idx,
                         "trg_label_suppix": self.trg_label_dataset.suppix, "trg_pool_suppix": self.trg_pool_dataset.suppix}, f)

    def load_datalist(self, datalist_path=None):
        if datalist_path is None:
            datalist_path = os.path.join(self.args.model_save_dir, f"datalist_{self.selection_iter:02d}.pkl")
        with open(datalist_path, "rb") as f:
            data = pickle.load(f)
        self.trg_label_dataset.im_idx = data["trg_label_im_idx"]
        self.trg_pool_dataset.im_idx = data["trg_pool_im_idx"]
        self.trg_label_dataset.suppix = data["trg_label_suppix"]
        self.trg_pool_dataset.suppix = data["trg_pool_suppix"]
