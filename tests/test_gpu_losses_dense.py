"""GPU parity of the DENSE (TMA-staged strip walk) loss kernels of csrc/losses_dense.cu: forced on
(``MAS_LOSS_DENSE=1``) they must give the oracle's losses and gradients on every shape the TMA path accepts
(W % 16 == 0), and the same answers as the tile / list walk they replace for densely selected batches."""
import types

import numpy as np
import pytest
import torch

from mulactseg_b200 import synth
from oracle import losses as olo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _step(x, trg, spx, mask, nseg, temp=0.1):
    from mulactseg_b200 import losses as L
    c = x.shape[1]
    group, multi = L.stage1_criterion(types.SimpleNamespace(nseg=nseg, group_ce_temp=temp, multi_ce_temp=temp), c - 1)
    xd = x.to(DEV).requires_grad_(True)
    td, sd, md = trg.to(DEV), spx.to(DEV), mask.to(DEV)
    g = group(xd, td, sd, md)
    ce, mc = multi(xd, td, sd, md)
    (16.0 * ce + 8.0 * mc + g).backward()
    torch.cuda.synchronize()
    return np.array([ce.item(), mc.item(), g.item()]), xd.grad.cpu().numpy()


# n, c, h, w, nseg, rho, p_extra, id dtype: ragged right edges (w % 128 != 0), one strip, many strips, generic channel
# padding (6, 12, 31) and exact instantiations (19-22), more than three candidate classes per superpixel (p_extra 0.5)
CASES = [(3, 20, 40, 64, 24, 0.5, 0.1, torch.int64), (2, 21, 33, 48, 30, 1.0, 0.06, torch.int32), (2, 6, 17, 128, 9, 0.3, 0.2, torch.int64),
         (1, 31, 12, 32, 5, 1.0, 0.5, torch.int32), (4, 19, 24, 272, 40, 0.15, 0.1, torch.int64), (2, 20, 21, 144, 12, 0.8, 0.5, torch.int64),
         (2, 22, 9, 400, 16, 1.0, 0.3, torch.int32), (1, 12, 64, 16, 6, 0.6, 0.3, torch.int64), (3, 3, 5, 256, 4, 1.0, 0.0, torch.int32)]


@pytest.mark.parametrize("case", CASES)
def test_dense_kernels_match_the_oracle(case, monkeypatch):
    n, c, h, w, nseg, rho, p_extra, id_dtype = case
    x = synth.logits(n, c, h, w, "cosine", seed=h + w)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=3), nseg, 1 if min(h, w) > 8 else 0)
    trg = synth.multihot_targets(n, nseg, c, seed=5, p_extra=p_extra, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, rho, seed=6)
    if n > 2:
        mask[1] = False
    xr = x.clone().requires_grad_(True)
    total_ref, parts_ref = olo.stage1_total(xr, trg, spx, mask, nseg, 0.1, 0.1)
    total_ref.backward()
    ref_grad = xr.grad.numpy()
    monkeypatch.setenv("MAS_LOSS_DENSE", "1")
    vals, grad = _step(x, trg, spx.to(id_dtype), mask, nseg)
    msg = f"{case}"
    np.testing.assert_allclose(vals, [float(v.detach()) if torch.is_tensor(v) else float(v) for v in parts_ref], rtol=1e-5, atol=1.2e-7, err_msg=msg)
    np.testing.assert_allclose(grad, ref_grad, rtol=1e-4, atol=1e-5 * float(np.abs(ref_grad).max()) + 1e-6, err_msg=msg)


@pytest.mark.parametrize("rho", [0.08, 0.5, 1.0])
def test_dense_kernels_equal_the_list_walk_at_crop_size(rho, monkeypatch):
    """Same batch through the tile / list walk (MAS_LOSS_DENSE=0), the dense kernels (=1) and the device's own pick
    (unset): equal losses (fp32 sums in a different order: 1e-6) and gradients (the same per-pixel arithmetic)."""
    n, c, h, w, nseg = 3, 20, 384, 768, 1024
    x = synth.logits(n, c, h, w, "cosine", seed=41, coherent=4)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=42), nseg, 8)
    trg = synth.multihot_targets(n, nseg, c, seed=43, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, rho, seed=44)
    monkeypatch.setenv("MAS_LOSS_DENSE", "0")
    vals0, grad0 = _step(x, trg, spx, mask, nseg)
    monkeypatch.setenv("MAS_LOSS_DENSE", "1")
    vals1, grad1 = _step(x, trg, spx, mask, nseg)
    monkeypatch.delenv("MAS_LOSS_DENSE")
    vals2, grad2 = _step(x, trg, spx, mask, nseg)
    for vals, grad, what in ((vals1, grad1, "dense"), (vals2, grad2, "auto")):
        np.testing.assert_allclose(vals, vals0, rtol=2e-6, err_msg=what)
        np.testing.assert_allclose(grad, grad0, rtol=1e-5, atol=1e-6 * float(np.abs(grad0).max()), err_msg=what)
        assert np.array_equal(grad == 0.0, grad0 == 0.0), what      # nothing leaks outside the selected pixels


def test_device_side_regime_switch_runs_exactly_one_kernel_set():
    """Both kernel sets are launched and one returns at once: a fully selected batch and a 2 % batch must each give the
    gradient exactly once (a double write would be invisible, a double accumulation of the bucket sums would not)."""
    from mulactseg_b200 import _lib, ops
    n, c, h, w, nseg = 2, 20, 128, 256, 64
    x = synth.logits(n, c, h, w, "cosine", seed=1).to(DEV)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=2).to(DEV)
    trg = synth.multihot_targets(n, nseg, c, seed=3, p_ignore=0.0).to(DEV)
    flags = _lib.MAS_LOSS_CHOICE | _lib.MAS_LOSS_GROUP
    for rho in (0.02, 1.0):
        mask = synth.region_mask(spx, nseg, rho, seed=4)
        info = ops.multihot_info(trg, c, _lib.MAS_GROUP_ONLYMULTI)
        acc_list, _ = ops.multihot_loss_forward(x, spx, mask, info, nseg, 0.1, flags, None)               # no list: tile walk
        acc_auto, _ = ops.multihot_loss_forward(x, spx, mask, info, nseg, 0.1, flags, ops.multihot_tiles(mask))
        torch.cuda.synchronize()
        a, b = acc_list.cpu().numpy(), acc_auto.cpu().numpy()
        assert np.array_equal(a[1::2], b[1::2]), rho                                                      # pixel counts: exact
        np.testing.assert_allclose(b[0::2], a[0::2], rtol=2e-6)
