"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference classes.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Run in the build container
only (needs ``/root/reference``):  ``python -m oracle.gen_golden``.

The reference classes are imported from the reference checkout through the
import shims of ``oracle/ref_shims.py`` (torch_scatter -> ``oracle.scatter_ref``;
the reference has no tests or golden vectors of its own) and
driven with an identity "network" (the dataset hands over logits as ``images``)
on small seeded synthetic tensors.  Inputs AND outputs are stored, so the
fixtures do not depend on RNG stability.
"""
from __future__ import annotations

import importlib
import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from mulactseg_b200 import synth  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class _IdentityNet:
    def eval(self):
        return self

    def __call__(self, x):
        return x


class _Pool(torch.utils.data.Dataset):
    def __init__(self, logits, spx, im_idx, suppix):
        self.logits, self.spx, self.im_idx, self.suppix = logits, spx, im_idx, suppix

    def __len__(self):
        return len(self.im_idx)

    def __getitem__(self, i):
        return {"images": self.logits[i], "spx": self.spx[i], "labels": self.spx[i]}


SELECTORS = ["my_bvsb", "my_bvsb_banignore", "my_bvsb_predclsbal_pwr", "my_bvsb_predclsbal_pwr_banignore",
             "my_bvsb_clsbal_v2", "my_bvsb_clsbal_v2_banignore"]


def _selector_args(method_name, nseg, num_classes, predignore, temp, coeff, bs):
    return types.SimpleNamespace(val_batch_size=bs, val_num_workers=0, nseg=nseg, active_method=method_name,
                                 num_classes=num_classes, ce_temp=temp, cls_weight_coeff=coeff,
                                 method="active_joint_multi_predignore_lossdecomp" if predignore
                                 else "active_joint_multi_lossdecomp", save_scores=False)


def gen_acquisition():
    cases = {
        # name: n, channels, h, w, nseg, map kind, logits kind, T, coeff, batch, dropped ids
        "city_small": (5, 8, 24, 40, 15, "jitter", "cosine", 0.1, 6.0, 2, 1),
        "voc_small": (4, 6, 19, 23, 6, "grid", "cosine", 0.1, 12.0, 4, 0),
        "adversarial": (3, 20, 16, 48, 64, "random", "normal", 1.0, 6.0, 2, 0),
    }
    out = {}
    for name, (n, c, h, w, nseg, mk, lk, temp, coeff, bs, drop) in cases.items():
        seed = abs(hash(name)) % 1000 if False else {"city_small": 11, "voc_small": 12, "adversarial": 13}[name]
        logits = synth.logits(n, c, h, w, lk, seed=seed)
        spx = synth.superpixel_map(n, h, w, nseg, mk, seed=seed + 100, drop_ids=drop)
        im_idx, suppix = synth.pool_lists(n, nseg)  # every id listed -> the list carries the full (N,S) tensor
        out[f"{name}/logits"] = logits.numpy()
        out[f"{name}/spx"] = spx.numpy().astype(np.int32)
        out[f"{name}/meta"] = np.array([nseg, bs], dtype=np.int64)
        out[f"{name}/temp_coeff"] = np.array([temp, coeff], dtype=np.float64)
        for sel in SELECTORS:
            ban = sel.endswith("banignore")
            # banignore selectors consume all C' = num_classes+1 channels and assert predignore;
            # my_bvsb slices the last channel for predignore nets; the others take C' = num_classes.
            for predignore in ([True] if ban else ([True, False] if sel == "my_bvsb" else [False])):
                num_classes = c - 1 if ban else c
                if sel == "my_bvsb" and predignore:
                    num_classes = c - 1
                args = _selector_args(sel, nseg, num_classes, predignore, temp, coeff, bs)
                mod = importlib.import_module(f"active_selection.{sel}")
                selector = mod.RegionSelector(args)
                trainer = types.SimpleNamespace(net=_IdentityNet(), device=torch.device("cpu"))
                pool = _Pool(logits, spx, im_idx, suppix)
                scores = selector.calculate_scores(trainer, pool)
                arr = np.array([s for s, _, _ in scores], dtype=np.float64).reshape(n, nseg)
                assert [i for _, _, i in scores] == list(range(nseg)) * n
                out[f"{name}/{sel}{'/predignore' if (sel == 'my_bvsb' and predignore) else ''}"] = arr
        # histogram of the arg-max class (what the ban/clsbal variants reduce), via the reference's own op chain
        my_bvsb = importlib.import_module("active_selection.my_bvsb")
        rs = my_bvsb.RegionSelector(_selector_args("my_bvsb", nseg, c, False, temp, coeff, bs))
        from torch_scatter import scatter
        _, top1 = rs.softmax_bvsb(logits)
        oh = torch.nn.functional.one_hot(top1.view(n, -1), num_classes=c)
        out[f"{name}/hist"] = scatter(oh, spx.view(n, -1), dim=1, reduce="sum", dim_size=nseg).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "acquisition.npz"), **out)
    print("acquisition.npz", len(out), "arrays")


def gen_selection():
    from dataloader.region_active_dataset import RegionActiveDataset
    rng = np.random.RandomState(5)
    n, nseg, c = 6, 9, 5
    out = {}
    for fair in (False, True):
        im_idx, suppix = synth.pool_lists(n, nseg)
        # some regions already labelled (round >= 2): image 0 fully pooled, image 1 partly labelled
        label_im_idx = [list(im_idx[1])]
        label_suppix = {im_idx[1][2]: [2, 5]}
        suppix[im_idx[1][2]] = [s for s in suppix[im_idx[1][2]] if s not in (2, 5)]
        multi_hot = (rng.rand(n, nseg, c) < 0.3).astype(np.uint8)
        multi_hot[..., 0] |= (multi_hot.sum(-1) == 0).astype(np.uint8)
        scores = []
        for k, key in enumerate(im_idx):
            for sid in suppix[key[2]]:
                # a few exact score ties exercise the (score, path, id) tuple order
                scores.append((float(np.float32(rng.randint(0, 12) / 12.0)), ",".join(key), sid))
        ranked = sorted(scores, reverse=True)
        tmp = tempfile.mkdtemp()
        args = types.SimpleNamespace(fair_counting=fair, or_labeling=fair, model_save_dir=tmp, finetune_itrs=1,
                                     wandb=types.SimpleNamespace(log=lambda *a, **k: None))
        pool_ds = types.SimpleNamespace(im_idx=[list(k) for k in im_idx], suppix={k: list(v) for k, v in suppix.items()})
        label_ds = types.SimpleNamespace(im_idx=label_im_idx, suppix=label_suppix,
                                         id_to_index={k[2].split("/")[-1].split(".")[0]: i for i, k in enumerate(im_idx)},
                                         multi_hot_cls=multi_hot)
        ads = RegionActiveDataset(args, pool_ds, label_ds)
        ads.selection_iter = 2
        budget = 11
        before = {"pool_im_idx": [list(k) for k in pool_ds.im_idx], "pool_suppix": {k: list(v) for k, v in pool_ds.suppix.items()},
                  "label_im_idx": [list(k) for k in label_ds.im_idx], "label_suppix": {k: list(v) for k, v in label_ds.suppix.items()}}
        ads.expand_training_set(ranked, budget, "unit")
        with open(os.path.join(tmp, "unit_selection_02.pkl"), "rb") as f:
            prefix = pickle.load(f)
        ads.dump_datalist()
        with open(os.path.join(tmp, "datalist_02.pkl"), "rb") as f:
            datalist = pickle.load(f)
        out["fair" if fair else "plain"] = {
            "budget": budget, "scores": scores, "before": before, "multi_hot": multi_hot.tolist(),
            "prefix": [list(t) for t in prefix], "datalist": datalist,
        }
    with open(os.path.join(GOLDEN, "selection.json"), "w") as f:
        json.dump(out, f)
    print("selection.json")


def gen_losses():
    from utils.loss import GroupMultiLabelCE, MultiChoiceCE
    from trainer.active_joint_multi_predignore import GroupMultiLabelCE_, MultiChoiceCE_
    from trainer.active_joint_multi_predignore_mclossablation2 import GroupMultiLabelCE_onlymulti
    from trainer.active_joint_multi_predignore_lossdecomp import OnehotCEMultihotChoice as DecompPredIgnore
    from trainer.active_joint_multi_lossdecomp import OnehotCEMultihotChoice as DecompVoc

    n, c, h, w, nseg = 4, 6, 20, 28, 12
    out = {}
    for case, (temp, rho, seed) in {"t01_rho05": (0.1, 0.5, 21), "t1_rho1": (1.0, 1.0, 22), "t01_rho02": (0.1, 0.2, 23)}.items():
        x = synth.logits(n, c, h, w, "cosine" if temp < 1 else "normal", seed=seed)
        spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=seed + 1), nseg, 2)
        trg = synth.multihot_targets(n, nseg, c, seed=seed + 2, p_extra=0.25)
        mask = synth.region_mask(spx, nseg, rho, seed=seed + 3)
        mask[2] = False  # an image with nothing labelled is skipped by every loss
        out[f"{case}/inputs"] = x.numpy()
        out[f"{case}/spx"] = spx.numpy().astype(np.int32)
        out[f"{case}/targets"] = trg.numpy()
        out[f"{case}/mask"] = mask.numpy()
        out[f"{case}/temp"] = np.array([temp])
        args = types.SimpleNamespace()
        mods = {
            "group_base": GroupMultiLabelCE(args, c, nseg, temperature=temp),
            "group_predignore": GroupMultiLabelCE_(args, c, nseg, temperature=temp),
            "group_onlymulti": GroupMultiLabelCE_onlymulti(args, c, nseg, temperature=temp),
            "mc_base": MultiChoiceCE(c, temperature=temp),
            "mc_predignore": MultiChoiceCE_(c, temperature=temp),
            "decomp_predignore": DecompPredIgnore(c, temperature=temp),
            "decomp_voc": DecompVoc(c, temperature=temp),
        }
        for name, m in mods.items():
            xin = x.clone().requires_grad_(True)
            # the *_base classes slice targets[..., :-1]: their net has one channel fewer than the targets
            xi = xin[:, :-1] if name.endswith("_base") else xin
            res = m(xi, trg, spx, mask)
            if isinstance(res, tuple):
                vals = torch.stack(list(res))
                total = 16.0 * res[0] + 8.0 * res[1]
            else:
                vals = res.reshape(1)
                total = res
            total.backward()
            out[f"{case}/{name}/value"] = vals.detach().numpy()
            out[f"{case}/{name}/grad"] = xin.grad.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "losses.npz"), **out)
    print("losses.npz", len(out), "arrays")


def gen_labeller():
    out = {}
    h, w, nseg, c, ch = 24, 36, 12, 6, 16
    for variant in ("eval_save_cosplbl_prop_includeonehot", "eval_save_cosplbl_prop"):
        mod = importlib.import_module(f"trainer.{variant}")
        for thr in ("median", "min"):
            for seed in (31, 32, 33):
                tr = mod.ActiveTrainer.__new__(mod.ActiveTrainer)
                tr.args = types.SimpleNamespace(nseg=nseg, cosprop_threshold_method=thr)
                tr.kernel = np.ones((3, 3), np.uint8)
                feats = synth.features(1, ch, h, w, seed=seed)
                logits = synth.logits(1, c, h, w, "normal", seed=seed + 1, coherent=4)
                spx = synth.superpixel_map(1, h, w, nseg, "jitter", seed=seed + 2)
                trg = synth.multihot_targets(1, nseg, c, seed=seed + 3, p_extra=0.3)
                mask = synth.region_mask(spx, nseg, 0.4, seed=seed + 4)
                labels = torch.zeros((1, h, w), dtype=torch.long)
                with torch.no_grad():
                    plbl = tr.pseudo_label_generation(labels, feats, logits, trg, mask, spx)
                key = f"{variant}/{thr}/{seed}"
                out[f"{key}/feats"] = feats.numpy()
                out[f"{key}/logits"] = logits.numpy()
                out[f"{key}/spx"] = spx.numpy().astype(np.int32)
                out[f"{key}/targets"] = trg.numpy()
                out[f"{key}/mask"] = mask.numpy()
                out[f"{key}/plbl"] = plbl.numpy().astype(np.int16)
    # candidate arg-max labeller
    mod = importlib.import_module("trainer.eval_within_multihot")
    tr = mod.ActiveTrainer.__new__(mod.ActiveTrainer)
    n = 2
    logits = synth.logits(n, c, h, w, "normal", seed=41)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=42)
    trg = synth.multihot_targets(n, nseg, c, seed=43, p_extra=0.3)
    mask = synth.region_mask(spx, nseg, 0.5, seed=44)
    plbl = tr.top_pseudo_label_generation(torch.zeros((n, h, w), dtype=torch.long), logits, trg, mask, spx)
    out.update({"top/logits": logits.numpy(), "top/spx": spx.numpy().astype(np.int32), "top/targets": trg.numpy(),
                "top/mask": mask.numpy(), "top/plbl": plbl.numpy().astype(np.int16)})
    np.savez_compressed(os.path.join(GOLDEN, "labeller.npz"), **out)
    print("labeller.npz", len(out), "arrays")


def gen_labelgen():
    """dataloader/region_cityscapes_tensor.py:23-84 ``RegionCityscapesTensor.__getitem__`` (offline multi-hot labels),
    run unmodified with its file I/O stubbed out (PIL open / transform / open_spx / encode_target)."""
    mod = importlib.import_module("dataloader.region_cityscapes_tensor")

    class _FakeImage:
        def convert(self, *_):
            return self

    mod.Image = types.SimpleNamespace(open=lambda *_a, **_k: _FakeImage())
    out = {}
    h, w, nseg, c = 40, 64, 24, 7
    for case, (trim, k, seed, kind) in {"notrim": (False, 3, 51, "jitter"), "trim3": (True, 3, 52, "jitter"),
                                        "trim5": (True, 5, 53, "jitter"), "trim2_random": (True, 2, 54, "random")}.items():
        spx = synth.superpixel_map(1, h, w, nseg, kind, seed=seed, drop_ids=1)[0]
        g = torch.Generator().manual_seed(seed)
        coarse = torch.randint(0, c, (1, 1, h // 8, w // 8), generator=g).float()
        target = torch.nn.functional.interpolate(coarse, size=(h, w), mode="nearest")[0, 0].long()
        target[torch.rand((h, w), generator=g) < 0.08] = 255
        target[spx == 3] = 255                                    # an all-ignore superpixel
        ids = sorted(set(torch.unique(spx).tolist()) - {5})       # one id not preserved -> zero row, size -1
        ds = mod.RegionCityscapesTensor.__new__(mod.RegionCityscapesTensor)
        ds.args = types.SimpleNamespace(nseg=nseg, num_classes=c, trim_multihot_boundary=trim, trim_kernel_size=k)
        ds.kernel = np.ones((k, k), np.uint8)
        ds.im_idx = [["img", "lbl", "spx"]]
        ds.suppix = {"spx": ids}
        ds.transform = lambda image, lbls: (image, lbls)
        ds.open_spx = lambda _f, _spx=spx: _spx.clone()
        ds.encode_target = lambda t: t
        mod.Image.open = lambda f, _t=target: _FakeImage() if f == "img" else _t.clone()
        cls, size = ds.__getitem__(0)["superpixel_info"]
        out[f"{case}/spx"] = spx.numpy().astype(np.int32)
        out[f"{case}/target"] = target.numpy().astype(np.uint8)
        out[f"{case}/ids"] = np.asarray(ids, dtype=np.int32)
        out[f"{case}/meta"] = np.asarray([nseg, c, k if trim else 0], dtype=np.int32)
        out[f"{case}/multi_hot"] = cls.numpy()
        out[f"{case}/size"] = size.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "labelgen.npz"), **out)
    print("labelgen.npz", len(out), "arrays")


def gen_metrics():
    """utils/miou.py ``MeanIoU`` (both step hooks, both epoch summaries) and the dominant label assignment of
    ``RegionCityscapesDominantAll.__getitem__`` (dataloader/region_dataset.py:201-240), both run unmodified (the
    dataset's file I/O stubbed out like in ``gen_labelgen``)."""
    miou = importlib.import_module("utils.miou")
    out = {}
    for case, (c, ignore, n, h, w, seed) in {"cs19": (19, 255, 3, 24, 40, 61), "voc22_ignore21": (22, 21, 2, 17, 33, 62),
                                             "tiny": (3, 255, 1, 4, 5, 63)}.items():
        g = torch.Generator().manual_seed(seed)
        targets = torch.randint(0, c, (n, h, w), generator=g)
        outputs = torch.where(torch.rand((n, h, w), generator=g) < 0.6, targets, torch.randint(0, c, (n, h, w), generator=g))
        targets[torch.rand((n, h, w), generator=g) < 0.15] = ignore
        outputs[torch.rand((n, h, w), generator=g) < 0.10] = ignore
        if c > 4:
            targets[targets == 2] = 3                                  # a class that never occurs in the targets
            outputs[outputs == 4] = 5                                  # a class that is never predicted
        out[f"{case}/meta"] = np.asarray([c, ignore], dtype=np.int64)
        out[f"{case}/outputs"] = outputs.numpy()
        out[f"{case}/targets"] = targets.numpy()
        for hook in ("_after_step", "_after_step_within_predregion"):
            helper = miou.MeanIoU(c, ignore)
            helper._before_epoch()
            for i in range(n):                                         # one "batch" per image, like the trainer loop
                getattr(helper, hook)({"outputs": outputs[i:i + 1], "targets": targets[i:i + 1]})
            out[f"{case}/{hook}/counts"] = np.stack([helper.total_seen, helper.total_correct, helper.total_positive]).astype(np.int64)
            out[f"{case}/{hook}/ious"] = np.asarray(helper._after_epoch(), dtype=np.float64)
            out[f"{case}/{hook}/ious_skip"] = np.asarray(helper._after_epoch(ignore_label_list=[0, c - 1]), dtype=np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                out[f"{case}/{hook}/ipr"] = np.asarray(helper._after_epoch_ipr(), dtype=np.float64)

    mod = importlib.import_module("dataloader.region_dataset")

    class _FakeImage:
        def convert(self, *_):
            return self

    h, w, nseg, c = 40, 64, 24, 7
    for case, (seed, kind) in {"dominant_jitter": (71, "jitter"), "dominant_random": (72, "random")}.items():
        spx = synth.superpixel_map(1, h, w, nseg, kind, seed=seed, drop_ids=1)[0]
        g = torch.Generator().manual_seed(seed)
        coarse = torch.randint(0, c, (1, 1, h // 4, w // 4), generator=g).float()
        target = torch.nn.functional.interpolate(coarse, size=(h, w), mode="nearest")[0, 0].long()
        target[torch.rand((h, w), generator=g) < 0.08] = 255
        target[spx == 3] = 255                                         # an all-ignore superpixel
        target[spx == 7] = torch.where(torch.arange(h * w).view(h, w)[spx == 7] % 2 == 0, 4, 1)   # a tie: the smaller id wins
        ids = sorted(set(torch.unique(spx).tolist()) - {5})            # one id outside the region dict: left as it is
        ds = mod.RegionCityscapesDominantAll.__new__(mod.RegionCityscapesDominantAll)
        ds.im_idx = [["img", "lbl", "spx"]]
        ds.suppix = {"spx": ids}
        ds.mask_region, ds.return_spx = True, False
        ds.transform = lambda image, lbls: (image, lbls)
        ds.open_spx = lambda _f, _spx=spx: _spx.clone()
        ds.encode_target = lambda t: t
        saved = mod.Image
        mod.Image = types.SimpleNamespace(open=lambda f, _t=target: _FakeImage() if f == "img" else _t.numpy().astype(np.uint8).copy())
        try:
            got = ds.__getitem__(0)["labels"]
        finally:
            mod.Image = saved
        out[f"{case}/spx"] = spx.numpy().astype(np.int32)
        out[f"{case}/target"] = target.numpy().astype(np.uint8)
        out[f"{case}/ids"] = np.asarray(ids, dtype=np.int32)
        out[f"{case}/meta"] = np.asarray([nseg, c], dtype=np.int32)
        out[f"{case}/labels"] = np.asarray(got).astype(np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "metrics.npz"), **out)
    print("metrics.npz", len(out), "arrays")


def main():
    ref_shims.install()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    gen_acquisition()
    gen_selection()
    gen_losses()
    gen_labeller()
    gen_labelgen()
    gen_metrics()


if __name__ == "__main__":
    main()
