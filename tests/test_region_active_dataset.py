"""CPU tier: the drop-in ``RegionActiveDataset`` writes the same containers and pickles as the reference
(golden: produced by the unmodified reference class, tests/golden/selection.json) and as the oracle walk."""
import copy
import json
import os
import pickle
import types

import numpy as np
import pytest

from mulactseg_b200 import synth
from mulactseg_b200.region_active_dataset import RegionActiveDataset
from oracle import acquisition as oa

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _dataset(tmp, fair, pool_im_idx, pool_suppix, label_im_idx, label_suppix, id_to_index, multi_hot):
    args = types.SimpleNamespace(fair_counting=fair, or_labeling=fair, model_save_dir=str(tmp), finetune_itrs=1,
                                 wandb=types.SimpleNamespace(log=lambda *a, **k: None))
    pool = types.SimpleNamespace(im_idx=pool_im_idx, suppix=pool_suppix)
    label = types.SimpleNamespace(im_idx=label_im_idx, suppix=label_suppix, id_to_index=id_to_index, multi_hot_cls=multi_hot)
    ads = RegionActiveDataset(args, pool, label)
    ads.selection_iter = 2
    return ads


@pytest.mark.parametrize("mode", ["plain", "fair"])
def test_matches_reference_golden(mode, tmp_path):
    with open(os.path.join(GOLDEN, "selection.json")) as f:
        g = json.load(f)[mode]
    b = g["before"]
    ranked = sorted([tuple(s) for s in g["scores"]], reverse=True)
    multi_hot = np.array(g["multi_hot"], dtype=np.uint8)
    im_all = sorted(set(tuple(k) for k in b["pool_im_idx"] + b["label_im_idx"]))
    id_to_index = {k[2].split("/")[-1].split(".")[0]: i for i, k in enumerate(im_all)}
    ads = _dataset(tmp_path, mode == "fair", b["pool_im_idx"], b["pool_suppix"], b["label_im_idx"], b["label_suppix"],
                   id_to_index, multi_hot)
    taken = ads.expand_training_set(ranked, g["budget"], "unit")
    with open(tmp_path / "unit_selection_02.pkl", "rb") as f:
        prefix = pickle.load(f)
    assert [list(t) for t in prefix] == g["prefix"] and taken == len(g["prefix"])
    ads.dump_datalist()
    with open(tmp_path / "datalist_02.pkl", "rb") as f:
        assert pickle.load(f) == g["datalist"]
    # round trip through load_datalist
    other = _dataset(tmp_path, False, [], {}, [], {}, {}, None)
    other.load_datalist()
    assert other.trg_pool_dataset.suppix == g["datalist"]["trg_pool_suppix"]


@pytest.mark.parametrize("seed,fair,budget", [(0, False, 40), (1, True, 55), (2, True, 10 ** 6), (3, False, 0)])
def test_matches_oracle_walk_on_random_pools(seed, fair, budget, tmp_path):
    rng = np.random.RandomState(seed)
    n, nseg, c = 12, 16, 6
    im_idx, suppix = synth.pool_lists(n, nseg, labelled_frac=0.3, seed=seed)
    label_im_idx = [list(im_idx[i]) for i in (3, 7)]
    label_suppix = {im_idx[3][2]: [99], im_idx[7][2]: [98, 97]}
    suppix[im_idx[5][2]] = suppix[im_idx[5][2]][:1] or [0]          # an image that runs empty after one pick
    multi_hot = (rng.rand(n, nseg, c) < 0.3).astype(np.uint8)
    multi_hot[..., 0] |= (multi_hot.sum(-1) == 0).astype(np.uint8)
    id_to_index = {k[2].split("/")[-1].split(".")[0]: i for i, k in enumerate(im_idx)}
    scores = [(float(np.float32(rng.randint(0, 9) / 9.0)), ",".join(k), s) for k in im_idx for s in suppix[k[2]]]
    ranked = sorted(scores, reverse=True)

    ref = copy.deepcopy(([list(k) for k in label_im_idx], label_suppix, [list(k) for k in im_idx], suppix))
    cost = (lambda p, s: multi_hot[id_to_index[p.split("/")[-1].split(".")[0]], s].sum()) if fair else None
    n_ref = oa.expand_training_set(ranked, budget, ref[0], ref[1], ref[2], ref[3], cost)

    ads = _dataset(tmp_path, fair, [list(k) for k in im_idx], copy.deepcopy(suppix), [list(k) for k in label_im_idx],
                   copy.deepcopy(label_suppix), id_to_index, multi_hot)
    ads.trg_pool_dataset.isselected = np.zeros((n, nseg), dtype=np.uint8)
    taken = ads.expand_training_set(ranked, budget, "unit")
    assert taken == n_ref
    assert ads.trg_label_dataset.im_idx == ref[0] and ads.trg_label_dataset.suppix == ref[1]
    assert ads.trg_pool_dataset.im_idx == ref[2] and ads.trg_pool_dataset.suppix == ref[3]
    assert int(ads.trg_pool_dataset.isselected.sum()) == taken
    assert os.path.exists(tmp_path / "unit_selection_02.pkl") == (n_ref < len(ranked) or
                                                                 (cost is None and n_ref > budget) or
                                                                 (cost is not None and sum(cost(r[1].split(",")[2], r[2]) for r in ranked[:n_ref]) > budget))
