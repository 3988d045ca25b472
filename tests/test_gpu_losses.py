"""GPU parity of the fused stage-1 losses (forward values and dense gradients) against the vectors produced by
the unmodified reference classes (tests/golden/losses.npz) and against the CPU oracle on further shapes."""
import os
import types

import numpy as np
import pytest
import torch

from mulactseg_b200 import synth
from oracle import losses as olo

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
LOS = np.load(os.path.join(GOLDEN, "losses.npz"))
DEV = "cuda:0"
RTOL = 1e-5   # north_star: losses within 1e-5 relative (fp32)


def _modules(c, nseg, temp):
    from mulactseg_b200 import losses as L
    args = types.SimpleNamespace()
    return {
        "group_base": L.GroupMultiLabelCE(args, c, nseg, temperature=temp),
        "group_predignore": L.GroupMultiLabelCE_(args, c, nseg, temperature=temp),
        "group_onlymulti": L.GroupMultiLabelCE_onlymulti(args, c, nseg, temperature=temp),
        "mc_base": L.MultiChoiceCE(c, temperature=temp),
        "mc_predignore": L.MultiChoiceCE_(c, temperature=temp),
        "decomp_predignore": L.OnehotCEMultihotChoice(c, temperature=temp),
        "decomp_voc": L.OnehotCEMultihotChoiceVOC(c, temperature=temp),
    }


def _run(module, x, trg, spx, mask, base):
    """-> (loss values, dense gradient, rounding scale).  scale = largest |d total / d bucket sum| / T: every gradient
    element is scale * (a difference of O(1) softmax terms), so it carries an absolute rounding error of a few
    float32 ulps of 1.0 times this, whatever the evaluation order."""
    from mulactseg_b200 import _lib, losses as L
    xin = x.clone().requires_grad_(True)
    xi = xin[:, :-1] if base else xin
    res = module(xi, trg, spx, mask)
    _, counts = L.segmented_loss_sums(xi.detach(), trg, spx, mask, module.temp,
                                      module.group_mode if module.group_mode is not None else _lib.MAS_GROUP_ALL, True)
    n = counts.cpu().numpy()
    if isinstance(res, tuple):
        vals = torch.stack(list(res))
        total = 16.0 * res[0] + 8.0 * res[1]
        scale = max(16.0 / (1 + n[0]), 8.0 / (1 + n[1] + (0 if module.strict_multihot else n[2])))
    else:
        vals, total = res.reshape(1), res
        scale = 1.0 / (1 + n[3]) if module.group_mode is not None else 1.0 / (1 + n[0] + n[1])
    total.backward()
    torch.cuda.synchronize()
    return vals.detach().cpu().numpy(), xin.grad.cpu().numpy(), scale / float(module.temp)


def _check(got, ref_vals, ref_grad, msg):
    """Loss values: 1e-5 relative (north_star).  Gradients: 1e-4 relative, plus 1e-5 of the largest gradient, plus
    4 ulps of the rounding scale (see ``_run``) absolute -- softmax backward subtracts nearly equal terms where P -> 1."""
    vals, grad, scale = got
    # atol: a probability near 1 is only known to half an fp32 ulp (6e-8), and so is -log of it -- the floor of ANY fp32
    # softmax, the reference's included; it matters only for losses below ~1e-2
    np.testing.assert_allclose(vals, ref_vals, rtol=RTOL, atol=1.2e-7, err_msg=msg)
    atol = 1e-5 * np.abs(ref_grad).max() + 4 * 1.2e-7 * scale
    np.testing.assert_allclose(grad, ref_grad, rtol=1e-4, atol=atol, err_msg=msg)


@pytest.mark.parametrize("case", ["t01_rho05", "t1_rho1", "t01_rho02"])
@pytest.mark.parametrize("id_dtype", [torch.int64, torch.int32])
def test_losses_match_reference_golden(case, id_dtype):
    x = torch.from_numpy(LOS[f"{case}/inputs"]).to(DEV)
    spx = torch.from_numpy(LOS[f"{case}/spx"]).to(DEV, id_dtype)
    trg = torch.from_numpy(LOS[f"{case}/targets"]).to(DEV)
    mask = torch.from_numpy(LOS[f"{case}/mask"]).to(DEV)
    temp = float(LOS[f"{case}/temp"][0])
    for name, module in _modules(x.shape[1], trg.shape[1], temp).items():
        got = _run(module, x, trg, spx, mask, name.endswith("_base"))
        _check(got, LOS[f"{case}/{name}/value"], LOS[f"{case}/{name}/grad"], f"{case}/{name}")


def _oracle(name, x, trg, spx, mask, nseg, temp):
    xin = x.clone().requires_grad_(True)
    if name == "group_onlymulti":
        res = olo.group_multilabel_ce(xin, trg, spx, mask, nseg, temp, "onlymulti")
    elif name == "group_predignore":
        res = olo.group_multilabel_ce(xin, trg, spx, mask, nseg, temp, "predignore")
    elif name == "mc_predignore":
        res = olo.multi_choice_ce(xin, trg, spx, mask, temp, "predignore")
    elif name == "decomp_voc":
        res = olo.onehot_ce_multihot_choice(xin, trg, spx, mask, temp, True)
    else:
        raise KeyError(name)
    if isinstance(res, tuple):
        res = tuple(torch.as_tensor(r, dtype=torch.float32) for r in res)   # a bucket without pixels is a python 0.0
        vals = torch.stack(list(res))
        total = 16.0 * res[0] + 8.0 * res[1]
    else:
        vals, total = res.reshape(1), res
    total.backward()
    return vals.detach().numpy(), xin.grad.numpy()


@pytest.mark.parametrize("shape", [(3, 20, 40, 64, 24, 0.5, 0.1), (2, 21, 33, 45, 30, 1.0, 1.0), (2, 6, 17, 128, 9, 0.3, 0.1),
                                   (1, 31, 12, 36, 5, 1.0, 0.5), (4, 19, 24, 260, 40, 0.15, 0.1)])
def test_losses_match_oracle(shape):
    n, c, h, w, nseg, rho, temp = shape
    x = synth.logits(n, c, h, w, "cosine" if temp < 1 else "normal", seed=h + w)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=3), nseg, 2)
    trg = synth.multihot_targets(n, nseg, c, seed=5, p_extra=0.06)
    mask = synth.region_mask(spx, nseg, rho, seed=6)
    if n > 1:
        mask[1] = False
    mods = _modules(c, nseg, temp)
    for name in ("group_onlymulti", "group_predignore", "mc_predignore", "decomp_voc"):
        ref_vals, ref_grad = _oracle(name, x, trg, spx, mask, nseg, temp)
        got = _run(mods[name], x.to(DEV), trg.to(DEV), spx.to(DEV), mask.to(DEV), False)
        _check(got, ref_vals, ref_grad, f"{shape}/{name}")


def test_shared_pass_equals_separate_passes_and_trainer_total():
    """The trainer's step (..._lossdecomp.py:101-104): group + decomposed loss on the same tensors."""
    from mulactseg_b200 import _lib, losses as L
    n, c, h, w, nseg = 3, 20, 48, 64, 32
    x = synth.logits(n, c, h, w, "cosine", seed=1)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=2), nseg, 3)
    trg = synth.multihot_targets(n, nseg, c, seed=3, p_extra=0.2, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, 0.6, seed=4)
    total_ref, (ce_ref, mc_ref, group_ref) = olo.stage1_total(x.clone().requires_grad_(True), trg, spx, mask, nseg, 0.1, 0.1)
    xr = x.clone().requires_grad_(True)
    olo.stage1_total(xr, trg, spx, mask, nseg, 0.1, 0.1)[0].backward()

    args = types.SimpleNamespace(nseg=nseg, group_ce_temp=0.1, multi_ce_temp=0.1)
    group, multi = L.stage1_criterion(args, c - 1)
    xd = x.to(DEV).requires_grad_(True)
    td, sd, md = trg.to(DEV), spx.to(DEV), mask.to(DEV)
    before = _lib.load().mas_kernel_launches()
    g = group(xd, td, sd, md)
    ce, mc = multi(xd, td, sd, md)
    loss = 16.0 * ce + 8.0 * mc + 1.0 * g
    loss.backward()
    torch.cuda.synchronize()
    launches = _lib.load().mas_kernel_launches() - before
    # candidate words + active-tile scan + fused forward (dense and list kernels: the device picks, one returns at once) +
    # group reduce + finish; coefficients + zero sweep + fused backward (dense and list)
    assert launches == 10
    np.testing.assert_allclose([ce.item(), mc.item(), g.item()], [ce_ref.item(), mc_ref.item(), group_ref.item()], rtol=RTOL)
    np.testing.assert_allclose(loss.item(), total_ref.item(), rtol=RTOL)
    ref_grad = xr.grad.numpy()
    np.testing.assert_allclose(xd.grad.cpu().numpy(), ref_grad, rtol=1e-4, atol=1e-5 * np.abs(ref_grad).max())

    # different temperatures: two passes, same answers as stand-alone modules
    args2 = types.SimpleNamespace(nseg=nseg, group_ce_temp=0.5, multi_ce_temp=0.1)
    group2, multi2 = L.stage1_criterion(args2, c - 1)
    g2 = group2(xd, td, sd, md)
    ce2, mc2 = multi2(xd, td, sd, md)
    alone = L.GroupMultiLabelCE_onlymulti(args2, c - 1, nseg, temperature=0.5)(xd, td, sd, md)
    assert g2.item() == alone.item() and ce2.item() == ce.item() and mc2.item() == mc.item()


def test_reduction_none_returns_sum_and_count_like_the_reference():
    """utils/loss.py:137-139 and :583-586: reduction='none' hands back (loss sum, num_valid) for the caller to combine."""
    from mulactseg_b200 import losses as L
    n, c, h, w, nseg = 2, 9, 24, 40, 12
    x = synth.logits(n, c, h, w, "cosine", seed=1)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=2), nseg, 2)
    trg = synth.multihot_targets(n, nseg, c + 1, seed=3, p_extra=0.2)
    mask = synth.region_mask(spx, nseg, 0.7, seed=4)
    args = types.SimpleNamespace()
    xd, td, sd, md = x.to(DEV), trg.to(DEV), spx.to(DEV), mask.to(DEV)
    for make, oracle in ((lambda red: L.MultiChoiceCE(c, temperature=0.1, reduction=red),
                          lambda: olo.multi_choice_ce(x, trg, spx, mask, 0.1, "base")),
                         (lambda red: L.GroupMultiLabelCE(args, c, nseg, temperature=0.1, reduction=red),
                          lambda: olo.group_multilabel_ce(x, trg, spx, mask, nseg, 0.1, "base"))):
        total, count = make("none")(xd, td, sd, md)
        mean = make("mean")(xd, td, sd, md)
        np.testing.assert_allclose(float(total) / float(count), float(mean), rtol=1e-6)
        np.testing.assert_allclose(float(mean), float(oracle()), rtol=RTOL)


def test_bf16_logits_are_accepted_and_differentiable():
    """Autocast heads hand over bf16 logits: widened once to fp32 (the kernels compute in fp32), gradient returned in bf16."""
    from mulactseg_b200 import losses as L
    n, c, h, w, nseg = 2, 20, 32, 64, 16
    x = synth.logits(n, c, h, w, "cosine", seed=1).to(torch.bfloat16)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=2).to(DEV)
    trg = synth.multihot_targets(n, nseg, c, seed=3, p_ignore=0.0).to(DEV)
    mask = synth.region_mask(spx, nseg, 0.7, seed=4)
    group, multi = L.stage1_criterion(types.SimpleNamespace(nseg=nseg, group_ce_temp=0.1, multi_ce_temp=0.1), c - 1)
    xb = x.to(DEV).requires_grad_(True)
    xf = x.float().to(DEV).requires_grad_(True)
    for xin in (xb, xf):
        ce, mc = multi(xin, trg, spx, mask)
        (16.0 * ce + 8.0 * mc + group(xin, trg, spx, mask)).backward()
    assert xb.grad.dtype == torch.bfloat16
    np.testing.assert_allclose(xb.grad.float().cpu().numpy(), xf.grad.cpu().numpy(), rtol=1e-2, atol=1e-2 * float(xf.grad.abs().max()))


def test_shared_pass_never_serves_stale_results():
    from mulactseg_b200 import losses as L
    n, c, h, w, nseg = 2, 8, 16, 32, 8
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=2).to(DEV)
    trg = synth.multihot_targets(n, nseg, c, seed=3, p_ignore=0.0).to(DEV)
    mask = synth.region_mask(spx, nseg, 0.7, seed=4)
    args = types.SimpleNamespace(nseg=nseg, group_ce_temp=0.1, multi_ce_temp=0.1)
    group, multi = L.stage1_criterion(args, c - 1)
    seen = []
    for step in range(4):                       # fresh logits every step, like net(images) in the trainer
        x = synth.logits(n, c, h, w, "cosine", seed=10 + step).to(DEV).requires_grad_(True)
        with torch.no_grad():
            g0 = group(x, trg, spx, mask)       # a no-grad evaluation must not poison the cache
        g = group(x, trg, spx, mask)
        ce, mc = multi(x, trg, spx, mask)
        (ce + mc + g).backward()
        assert x.grad is not None and g0.item() == g.item()
        alone = L.GroupMultiLabelCE_onlymulti(args, c - 1, nseg, temperature=0.1)(x, trg, spx, mask)
        assert alone.item() == g.item()
        seen.append(g.item())
        del x
    assert len(set(seen)) == 4
    x = synth.logits(n, c, h, w, "cosine", seed=99).to(DEV)
    a = group(x, trg, spx, mask).item()
    x.mul_(0.5)                                 # in-place change bumps the version: recomputed
    assert group(x, trg, spx, mask).item() != a


def test_empty_candidate_row_raises_like_the_reference():
    from mulactseg_b200 import losses as L
    n, c, h, w, nseg = 1, 6, 8, 16, 4
    x = synth.logits(n, c, h, w, "normal", seed=1).to(DEV)
    spx = synth.superpixel_map(n, h, w, nseg, "grid", seed=2).to(DEV)
    trg = synth.multihot_targets(n, nseg, c, seed=3).to(DEV)
    trg[0, 1] = 0                                   # a selected superpixel without any candidate class
    mask = torch.ones((n, h, w), dtype=torch.bool, device=DEV)
    with pytest.raises(AssertionError):
        L.OnehotCEMultihotChoice(c, temperature=1.0)(x, trg, spx, mask)
    one, multi = L.OnehotCEMultihotChoice(c, temperature=1.0, assert_partition=False)(x, trg, spx, mask)
    assert torch.isfinite(one) and torch.isfinite(multi)
    # deferred check (what stage1_criterion installs): same assertion, raised by the NEXT call instead of stalling this one
    lazy = L.OnehotCEMultihotChoice(c, temperature=1.0, assert_partition="deferred")
    lazy(x, trg, spx, mask)
    with pytest.raises(AssertionError):
        lazy(x, trg, spx, mask)
    lazy(x, trg, spx, mask)
    with pytest.raises(AssertionError):
        lazy.check_partition()
    with pytest.raises(RuntimeError):
        L.MultiChoiceCE_(c)(x.cpu(), trg.cpu(), spx.cpu(), mask.cpu())      # no CPU path


def test_full_size_properties():
    """BASELINE config 4 shape (per image): size-independent invariants instead of the slow oracle."""
    from mulactseg_b200 import _lib, losses as L
    n, c, h, w, nseg = 4, 20, 768, 768, 2048
    x = synth.logits(n, c, h, w, "cosine", seed=1, device=DEV, coherent=4)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=2, device=DEV), nseg, 16)
    trg = synth.multihot_targets(n, nseg, c, seed=3, device=DEV, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, 0.2, seed=4)
    xin = x.clone().requires_grad_(True)
    sums, counts = L.segmented_loss_sums(xin, trg, spx, mask, 0.1, _lib.MAS_GROUP_ONLYMULTI, True)
    (16.0 * sums[0] + 8.0 * sums[1] + sums[3]).backward()
    torch.cuda.synchronize()
    ncand = trg.sum(dim=2)
    pix_n = torch.gather(ncand, 1, spx.reshape(n, -1).clamp(max=nseg - 1)).view(n, h, w)
    # bucket counts are exact pixel counts
    assert int(counts[0]) == int((mask & (pix_n == 1)).sum())
    assert int(counts[1]) == int((mask & (pix_n > 1)).sum())
    assert int(counts[2]) == 0
    # group count = labelled classes of multi-hot superpixels that own at least one selected pixel
    chosen = torch.zeros((n, nseg + 1), dtype=torch.bool, device=DEV)
    chosen.scatter_(1, spx.reshape(n, -1), mask.reshape(n, -1))   # pads land in column nseg; masks are per region
    multi = chosen[:, :nseg] & (ncand > 1)
    assert int(counts[3]) == int(trg[multi].sum())
    g = xin.grad
    # softmax gradients sum to zero over the classes; nothing leaks outside the mask
    assert float(g.sum(dim=1).abs().max()) < 1e-4 * float(g.abs().max())
    assert float(g.permute(0, 2, 3, 1)[~mask].abs().max()) == 0.0
    # the sums agree with a straightforward torch evaluation on the device (fp32, 1e-5)
    p = torch.softmax(x / 0.1, dim=1)
    rows = trg[torch.arange(n, device=DEV).view(n, 1, 1), spx.clamp(max=nseg - 1)]          # (n,h,w,c)
    pos = (p.permute(0, 2, 3, 1) * rows).sum(dim=3)
    l = -torch.log(pos + 1e-8)
    np.testing.assert_allclose(float(sums[0].detach()), float(l[mask & (pix_n == 1)].double().sum()), rtol=RTOL)
    np.testing.assert_allclose(float(sums[1].detach()), float(l[mask & (pix_n > 1)].double().sum()), rtol=RTOL)
