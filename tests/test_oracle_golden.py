"""Pin the CPU oracle against vectors produced by the unmodified reference (oracle/gen_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import acquisition as oa
from oracle import labeller as ol
from oracle import losses as olo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ACQ = np.load(os.path.join(GOLDEN, "acquisition.npz"))
LOS = np.load(os.path.join(GOLDEN, "losses.npz"))
LAB = np.load(os.path.join(GOLDEN, "labeller.npz"))
ACQ_CASES = ["city_small", "voc_small", "adversarial"]


def _pool(case):
    logits = torch.from_numpy(ACQ[f"{case}/logits"])
    spx = torch.from_numpy(ACQ[f"{case}/spx"]).long()
    nseg, bs = (int(v) for v in ACQ[f"{case}/meta"])
    temp, coeff = (float(v) for v in ACQ[f"{case}/temp_coeff"])
    pool = [(logits[i:i + bs], spx[i:i + bs]) for i in range(0, logits.shape[0], bs)]
    return pool, nseg, temp, coeff


@pytest.mark.parametrize("case", ACQ_CASES)
def test_acquisition_selectors_match_reference(case):
    pool, nseg, temp, coeff = _pool(case)
    got = {
        "my_bvsb": oa.scores_my_bvsb(pool, nseg, temp, predignore=False),
        "my_bvsb/predignore": oa.scores_my_bvsb(pool, nseg, temp, predignore=True),
        "my_bvsb_banignore": oa.scores_my_bvsb_banignore(pool, nseg, temp),
        "my_bvsb_predclsbal_pwr": oa.scores_predclsbal_pwr(pool, nseg, temp, coeff, ban_ignore=False),
        "my_bvsb_predclsbal_pwr_banignore": oa.scores_predclsbal_pwr(pool, nseg, temp, coeff, ban_ignore=True),
        "my_bvsb_clsbal_v2": oa.scores_clsbal_v2(pool, nseg, temp, ban_ignore=False),
        "my_bvsb_clsbal_v2_banignore": oa.scores_clsbal_v2(pool, nseg, temp, ban_ignore=True),
    }
    for name, val in got.items():
        ref = ACQ[f"{case}/{name}"]
        # same op chain on the same machine -> identical floats
        np.testing.assert_array_equal(val.double().numpy(), ref, err_msg=name)
    hist = oa.region_histograms(pool, nseg, temp)
    np.testing.assert_array_equal(hist.numpy(), ACQ[f"{case}/hist"])


@pytest.mark.parametrize("mode", ["plain", "fair"])
def test_region_selection_matches_reference(mode):
    with open(os.path.join(GOLDEN, "selection.json")) as f:
        g = json.load(f)[mode]
    scores = [tuple(s) for s in g["scores"]]
    ranked = oa.rank_regions(scores)
    b = g["before"]
    multi_hot = np.array(g["multi_hot"])
    index_of = {k[2]: i for i, k in enumerate(sorted(set(tuple(k) for k in b["pool_im_idx"] + b["label_im_idx"])))}
    cost = (lambda p, s: multi_hot[index_of[p], s].sum()) if mode == "fair" else None
    n = oa.expand_training_set(ranked, g["budget"], b["label_im_idx"], b["label_suppix"],
                               b["pool_im_idx"], b["pool_suppix"], cost)
    assert [list(t) for t in ranked[:n]] == g["prefix"]
    d = g["datalist"]
    assert b["label_im_idx"] == d["trg_label_im_idx"]
    assert b["pool_im_idx"] == d["trg_pool_im_idx"]
    assert b["label_suppix"] == d["trg_label_suppix"]
    assert b["pool_suppix"] == d["trg_pool_suppix"]


LOSS_CASES = ["t01_rho05", "t1_rho1", "t01_rho02"]


@pytest.mark.parametrize("case", LOSS_CASES)
def test_losses_match_reference(case):
    x = torch.from_numpy(LOS[f"{case}/inputs"])
    spx = torch.from_numpy(LOS[f"{case}/spx"]).long()
    trg = torch.from_numpy(LOS[f"{case}/targets"])
    mask = torch.from_numpy(LOS[f"{case}/mask"])
    temp = float(LOS[f"{case}/temp"][0])
    nseg = trg.shape[1]
    fns = {
        "group_base": lambda xi: olo.group_multilabel_ce(xi[:, :-1], trg, spx, mask, nseg, temp, "base"),
        "group_predignore": lambda xi: olo.group_multilabel_ce(xi, trg, spx, mask, nseg, temp, "predignore"),
        "group_onlymulti": lambda xi: olo.group_multilabel_ce(xi, trg, spx, mask, nseg, temp, "onlymulti"),
        "mc_base": lambda xi: olo.multi_choice_ce(xi[:, :-1], trg, spx, mask, temp, "base"),
        "mc_predignore": lambda xi: olo.multi_choice_ce(xi, trg, spx, mask, temp, "predignore"),
        "decomp_predignore": lambda xi: olo.onehot_ce_multihot_choice(xi, trg, spx, mask, temp, False),
        "decomp_voc": lambda xi: olo.onehot_ce_multihot_choice(xi, trg, spx, mask, temp, True),
    }
    for name, fn in fns.items():
        xin = x.clone().requires_grad_(True)
        res = fn(xin)
        if isinstance(res, tuple):
            vals = torch.stack(list(res))
            total = 16.0 * res[0] + 8.0 * res[1]
        else:
            vals, total = res.reshape(1), res
        total.backward()
        np.testing.assert_allclose(vals.detach().numpy(), LOS[f"{case}/{name}/value"], rtol=1e-6, atol=0, err_msg=name)
        ref_grad = LOS[f"{case}/{name}/grad"]
        np.testing.assert_allclose(xin.grad.numpy(), ref_grad, rtol=1e-5, atol=1e-6 * np.abs(ref_grad).max(), err_msg=name)


LAB_KEYS = sorted({k.rsplit("/", 1)[0] for k in LAB.files if k.startswith("eval_save")})


@pytest.mark.parametrize("key", LAB_KEYS)
def test_labeller_matches_reference(key):
    variant, thr, _ = key.split("/")
    feats = torch.from_numpy(LAB[f"{key}/feats"])
    logits = torch.from_numpy(LAB[f"{key}/logits"])
    spx = torch.from_numpy(LAB[f"{key}/spx"]).long()
    trg = torch.from_numpy(LAB[f"{key}/targets"])
    mask = torch.from_numpy(LAB[f"{key}/mask"])
    got = ol.pseudo_label_generation(feats, logits, trg, mask, spx,
                                     only_multihot=(variant == "eval_save_cosplbl_prop"), threshold=thr)
    np.testing.assert_array_equal(got.numpy(), LAB[f"{key}/plbl"].astype(np.int64))


def test_top_labeller_matches_reference():
    got = ol.top_pseudo_label_generation(torch.from_numpy(LAB["top/logits"]), torch.from_numpy(LAB["top/targets"]),
                                         torch.from_numpy(LAB["top/mask"]), torch.from_numpy(LAB["top/spx"]).long())
    np.testing.assert_array_equal(got.numpy(), LAB["top/plbl"].astype(np.int64))
