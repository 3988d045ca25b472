"""Offline multi-hot label generation: oracle vs the vectors produced by the unmodified reference class (CPU tier)
and the CUDA kernels vs both (GPU tier).  Integer outputs: bit-exact."""
import os

import numpy as np
import pytest
import torch

from mulactseg_b200 import synth
from oracle import labelgen as olg

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
LG = np.load(os.path.join(GOLDEN, "labelgen.npz"))
CASES = sorted({k.split("/")[0] for k in LG.files})


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(case):
    nseg, c, k = (int(v) for v in LG[f"{case}/meta"])
    cls, size = olg.multi_hot_labels(LG[f"{case}/target"], LG[f"{case}/spx"].astype(np.int64), LG[f"{case}/ids"].tolist(), nseg, c, k)
    np.testing.assert_array_equal(cls, LG[f"{case}/multi_hot"])
    np.testing.assert_array_equal(size, LG[f"{case}/size"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("id_dtype", [torch.int64, torch.int32])
def test_kernels_match_reference_golden(case, id_dtype):
    from mulactseg_b200 import label_assignment
    nseg, c, k = (int(v) for v in LG[f"{case}/meta"])
    cls, size = label_assignment.superpixel_info(torch.from_numpy(LG[f"{case}/target"]), torch.from_numpy(LG[f"{case}/spx"]).to(id_dtype),
                                                 LG[f"{case}/ids"].tolist(), nseg, c, k > 0, max(k, 1))
    np.testing.assert_array_equal(cls.numpy(), LG[f"{case}/multi_hot"])
    np.testing.assert_array_equal(size.numpy(), LG[f"{case}/size"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(97, 131, 60, 19, 3, "jitter"), (64, 64, 30, 21, 4, "grid"), (50, 70, 2048, 19, 5, "random"),
                                   (128, 256, 128, 19, 0, "jitter")])
def test_kernels_match_oracle(shape):
    from mulactseg_b200 import label_assignment
    h, w, nseg, c, k, kind = shape
    spx = synth.superpixel_map(1, h, w, nseg, kind, seed=h, drop_ids=2)[0]
    g = torch.Generator().manual_seed(w)
    target = torch.randint(0, c, (h, w), generator=g)
    target[torch.rand((h, w), generator=g) < 0.1] = 255
    ids = torch.unique(spx).tolist()[::2] + [nseg - 1]
    ref_cls, ref_size = olg.multi_hot_labels(target.numpy(), spx.numpy(), sorted(set(ids)), nseg, c, k)
    cls, size = label_assignment.superpixel_info(target, spx, sorted(set(ids)), nseg, c, k > 0, max(k, 1))
    np.testing.assert_array_equal(cls.numpy(), ref_cls)
    np.testing.assert_array_equal(size.numpy(), ref_size)


@pytest.mark.gpu
def test_full_size_counts_are_exact():
    from mulactseg_b200 import label_assignment
    h, w, nseg, c = 1024, 2048, 2048, 19
    spx = synth.superpixel_map(1, h, w, nseg, "jitter", seed=1, device="cuda")[0]
    target = torch.randint(0, c, (h, w), device="cuda")
    cls, size = label_assignment.superpixel_info(target, spx, list(range(nseg)), nseg, c)
    assert torch.equal(size.long(), torch.bincount(spx.reshape(-1).cpu(), minlength=nseg))
    key = (spx.reshape(-1) * c + target.reshape(-1)).cpu()
    present = torch.bincount(key, minlength=nseg * c).view(nseg, c) > 0
    assert torch.equal(cls[:, :c].bool(), present) and int(cls[:, c].sum()) == 0
