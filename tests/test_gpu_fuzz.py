"""Seeded random-shape sweeps (GPU tier): every kernel family against the CPU oracle on small, awkward shapes -- single rows and
columns, widths that are not multiples of 4 / 32, a single superpixel, two classes, masks that are all set or all clear,
pad ids, batches shorter than the reference batch.  Integer outputs bit-exact, fp32 within the north_star tolerance."""
import types

import numpy as np
import pytest
import torch

from helpers import assert_scores_close, batches
from mulactseg_b200 import synth
from oracle import acquisition as oa, labeller as ol, losses as olo, metrics as om

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rng_shape(rng, max_h=70, max_w=150):
    h = int(rng.choice([1, 2, 3, 15, 16, 17, 31, 33, int(rng.randint(1, max_h))]))
    w = int(rng.choice([1, 3, 4, 31, 32, 33, 63, 65, 127, 128, 129, int(rng.randint(1, max_w))]))
    return h, w


@pytest.mark.parametrize("seed", range(16))
def test_acquisition_random_shapes(seed):
    from mulactseg_b200 import acquisition as acq
    rng = np.random.RandomState(1000 + seed)
    h, w = _rng_shape(rng)
    n = int(rng.randint(1, 6))
    c = int(rng.choice([2, 3, 7, 19, 20, 21, 22, 32]))
    nseg = int(rng.choice([1, 2, 5, 40, 150, 2048]))
    bs = int(rng.randint(1, 4))
    kind = str(rng.choice(["jitter", "grid", "random"]))
    temp = float(rng.choice([0.1, 1.0]))
    logits = synth.logits(n, c, h, w, "cosine" if temp < 1 else "normal", seed=seed)
    spx = synth.superpixel_map(n, h, w, nseg, kind, seed=seed + 1)
    pool = batches(logits, spx, bs)
    for method, ref in (("my_bvsb_predclsbal_pwr_banignore", oa.scores_predclsbal_pwr(pool, nseg, temp, 6.0, ban_ignore=True)),
                        ("my_bvsb_banignore", oa.scores_my_bvsb_banignore(pool, nseg, temp)),
                        ("my_bvsb_clsbal_v2_banignore", oa.scores_clsbal_v2(pool, nseg, temp, ban_ignore=True))):
        spec = acq.SELECTORS[method]
        stats = acq.RegionStats(n, nseg, c, DEV, need_prob=spec.weighting == "predclsbal")
        for i in range(0, n, bs):
            stats.add_batch(i, logits[i:i + bs].to(DEV), spx[i:i + bs].to(DEV, torch.int32), temp)
        score, _ = acq.finalize(stats, spec, 6.0, bs)
        msg = f"seed {seed}: n={n} c={c} {h}x{w} nseg={nseg} bs={bs} {kind} T={temp} {method}"
        if spec.normalise and not np.isfinite(ref.numpy()).all():
            continue            # a pool whose scores are all equal: the reference divides 0 by 0
        if spec.normalise:
            # (u - min) / (max - min) turns a relative error e on the region means into e * max / (max - min) absolute:
            # tiny pools of near-identical regions (one superpixel per image ...) are ill-conditioned in the reference too
            raw = (stats.cls_sum.double().sum(-1) / stats.cls_cnt.sum(-1).clamp(min=1)).cpu().numpy().reshape(-1)
            nz = raw[raw != 0]
            amp = float(nz.max() / max(nz.max() - nz.min(), 1e-30)) if nz.size else 1.0
            np.testing.assert_allclose(score.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5 * max(1.0, amp), err_msg=msg)
        else:
            assert_scores_close(score.cpu().numpy(), ref.numpy(), False, msg)
        np.testing.assert_array_equal(stats.cls_cnt.cpu().numpy().astype(np.int64), oa.region_histograms(pool, nseg, temp).numpy(), err_msg=msg)


@pytest.mark.parametrize("seed", range(16))
def test_losses_random_shapes(seed):
    from mulactseg_b200 import losses as L
    rng = np.random.RandomState(2000 + seed)
    h, w = _rng_shape(rng, 50, 100)
    n = int(rng.randint(1, 4))
    c = int(rng.choice([2, 3, 8, 20, 21, 31]))
    nseg = int(rng.choice([1, 3, 12, 64]))
    rho = float(rng.choice([0.0, 0.05, 0.5, 1.0]))
    temp = float(rng.choice([0.1, 0.5, 1.0]))
    x = synth.logits(n, c, h, w, "cosine" if temp < 1 else "normal", seed=seed)
    spx = synth.superpixel_map(n, h, w, nseg, str(rng.choice(["jitter", "grid", "random"])), seed=seed + 1)
    if min(h, w) > 4:
        spx = synth.pad_border(spx, nseg, 1)                       # crop-padding id = nseg on the border
    trg = synth.multihot_targets(n, nseg, c, seed=seed + 2, p_extra=float(rng.choice([0.0, 0.1, 0.5])))
    mask = synth.region_mask(spx, nseg, rho, seed=seed + 3) if rho > 0 else torch.zeros((n, h, w), dtype=torch.bool)
    mask = mask & (spx < nseg)                                      # spmask always excludes the pad id (utils/loss.py:99,107)
    args = types.SimpleNamespace(nseg=nseg, group_ce_temp=temp, multi_ce_temp=temp)
    group, multi = L.stage1_criterion(args, c - 1, voc=True)
    xd = x.to(DEV).requires_grad_(True)
    g = group(xd, trg.to(DEV), spx.to(DEV), mask.to(DEV))
    ce, mc = multi(xd, trg.to(DEV), spx.to(DEV), mask.to(DEV))
    (16.0 * ce + 8.0 * mc + g).backward()
    xr = x.clone().requires_grad_(True)
    g_ref = olo.group_multilabel_ce(xr, trg, spx, mask, nseg, temp, "onlymulti")
    ce_ref, mc_ref = (torch.as_tensor(v, dtype=torch.float32) for v in olo.onehot_ce_multihot_choice(xr, trg, spx, mask, temp, True))
    total_ref = 16.0 * ce_ref + 8.0 * mc_ref + g_ref
    msg = f"seed {seed}: n={n} c={c} {h}x{w} nseg={nseg} rho={rho} T={temp}"
    np.testing.assert_allclose([ce.item(), mc.item(), g.item()], [float(torch.as_tensor(v).detach()) for v in (ce_ref, mc_ref, g_ref)], rtol=1e-5, atol=1.2e-7, err_msg=msg)
    if total_ref.requires_grad:
        total_ref.backward()
        ref_grad = xr.grad.numpy()
        atol = 1e-5 * np.abs(ref_grad).max() + 4 * 1.2e-7 * 16.0 / temp
        np.testing.assert_allclose(xd.grad.cpu().numpy(), ref_grad, rtol=1e-4, atol=atol, err_msg=msg)
    else:
        assert float(xd.grad.abs().max()) == 0.0


@pytest.mark.parametrize("seed", range(12))
def test_labeller_random_shapes(seed):
    from mulactseg_b200 import labeller
    rng = np.random.RandomState(3000 + seed)
    h, w = int(rng.randint(2, 40)), int(rng.randint(2, 70))
    c = int(rng.choice([2, 5, 12, 21]))
    ch = int(rng.choice([1, 7, 16, 33]))
    nseg = int(rng.choice([1, 2, 9, 40]))
    rho = float(rng.choice([0.2, 0.6, 1.0]))
    feats = synth.features(1, ch, h, w, seed=seed) if min(h, w) >= 4 else torch.nn.functional.normalize(torch.randn(1, ch, h, w, generator=torch.Generator().manual_seed(seed)), dim=1)
    logits = synth.logits(1, c, h, w, "normal", seed=seed + 1)
    spx = synth.superpixel_map(1, h, w, nseg, str(rng.choice(["jitter", "grid", "random"])), seed=seed + 2)
    trg = synth.multihot_targets(1, nseg, c, seed=seed + 3, p_extra=float(rng.choice([0.0, 0.3, 0.9])))
    mask = synth.region_mask(spx, nseg, rho, seed=seed + 4)
    only_multi, thr = bool(rng.randint(0, 2)), str(rng.choice(["median", "min"]))
    ref = ol.pseudo_label_generation(feats, logits, trg, mask, spx, only_multihot=only_multi, threshold=thr)
    got = labeller.pseudo_label_generation(None, feats.to(DEV), logits.to(DEV), trg.to(DEV), mask.to(DEV), spx.to(DEV), only_multi, thr).cpu()
    assert int((got != ref).sum()) == 0, f"seed {seed}: {h}x{w} c={c} F={ch} nseg={nseg} rho={rho} {only_multi} {thr}"
    ref_top = ol.top_pseudo_label_generation(logits, trg, mask, spx)
    got_top = labeller.top_pseudo_label_generation(None, logits.to(DEV), trg.to(DEV), mask.to(DEV), spx.to(DEV)).cpu()
    np.testing.assert_array_equal(got_top.numpy(), ref_top.numpy())


@pytest.mark.parametrize("seed", range(8))
def test_selection_and_metrics_random(seed):
    from mulactseg_b200 import label_assignment, ops, selection
    from mulactseg_b200.miou import MeanIoU
    rng = np.random.RandomState(4000 + seed)
    # ranked + cut list == numpy on the same keys
    n_img, nseg = int(rng.randint(1, 9)), int(rng.choice([1, 7, 150]))
    score = torch.from_numpy(rng.rand(n_img, nseg).astype(np.float32) * (rng.rand(n_img, nseg) < 0.8))
    in_pool = torch.from_numpy((rng.rand(n_img, nseg) < 0.7).astype(np.uint8))
    rank = torch.from_numpy(rng.permutation(n_img).astype(np.int32))
    cost = rng.randint(0, 4, size=n_img * nseg).astype(np.uint8)
    k, budget = int(rng.randint(1, n_img * nseg + 3)), int(rng.randint(0, 2 * n_img * nseg + 1))
    keys = selection.top_regions(score.to(DEV), in_pool.to(DEV), rank.to(DEV), k, None, torch.from_numpy(cost).to(DEV), budget)
    all_keys = ops.region_keys(score.to(DEV), in_pool.to(DEV), rank.to(DEV)).cpu().numpy().view(np.uint64)
    want = np.sort(all_keys[all_keys != 0])[::-1][:k]
    want = want[: selection.cumulative_cut(cost[(want & np.uint64(0xFFFFFFFF)).astype(np.int64)], budget)]
    np.testing.assert_array_equal(keys, want)
    # mIoU counters and dominant labels on odd sizes
    c, h, w = int(rng.choice([1, 2, 19, 200])), int(rng.randint(1, 40)), int(rng.randint(1, 70))
    t = torch.from_numpy(rng.randint(0, c + 2, size=(2, h, w))).long()
    o = torch.from_numpy(rng.randint(0, c + 2, size=(2, h, w))).long()
    helper = MeanIoU(c, c + 1)
    helper._before_epoch()
    helper._after_step({"outputs": o.to(DEV), "targets": t.to(DEV)})
    np.testing.assert_array_equal(np.stack([helper.total_seen, helper.total_correct, helper.total_positive]).astype(np.int64),
                                  om.miou_counts(o.numpy(), t.numpy(), c, c + 1))
    nseg2, c2 = int(rng.choice([1, 5, 64])), int(rng.choice([1, 7, 19]))
    spx = synth.superpixel_map(1, h, w, nseg2, "random" if seed % 2 else "jitter", seed=seed)[0]
    target = torch.from_numpy(rng.randint(0, c2, size=(h, w))).to(torch.uint8)
    target[torch.from_numpy(rng.rand(h, w) < 0.2)] = 255
    ids = sorted(set(rng.randint(0, nseg2, size=max(1, nseg2 // 2)).tolist()))
    np.testing.assert_array_equal(label_assignment.dominant_target(target, spx, ids, nseg2, c2).numpy(),
                                  om.dominant_target(target.numpy(), spx.numpy(), ids))


@pytest.mark.parametrize("seed", range(10))
def test_lowres_scorer_random_shapes(seed):
    """mas_bvsb_segment_stats_lowres_dev on random source / target sizes (integer and fractional ratios, identity, one low-
    resolution row or column, odd widths -> scalar id loads) against our own full-resolution kernels fed with
    F.interpolate(..., 'bilinear', align_corners=False) computed on the GPU; regions with a near-tied pixel are left out."""
    from mulactseg_b200 import acquisition as acq
    rng = np.random.RandomState(3000 + seed)
    n = int(rng.randint(1, 4))
    c = int(rng.choice([2, 5, 19, 20, 22]))
    h_in, w_in = int(rng.choice([1, 2, 7, 16, 33])), int(rng.choice([1, 3, 8, 31, 32]))
    fy, fx = float(rng.choice([1.0, 2.0, 3.977, 4.0])), float(rng.choice([1.0, 2.5, 4.0, 4.0]))
    h, w = max(h_in, int(round(h_in * fy))), max(w_in, int(round(w_in * fx)))
    nseg = int(rng.choice([1, 6, 40]))
    dtype = torch.bfloat16 if seed % 3 == 2 else torch.float32
    low = synth.logits(n, c, h_in, w_in, "cosine", seed=seed).to(dtype).to(DEV)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=seed + 1).to(DEV, torch.int32)
    full = torch.nn.functional.interpolate(low.float(), size=(h, w), mode="bilinear", align_corners=False)
    a = acq.RegionStats(n, nseg, c, DEV, need_prob=True)
    b = acq.RegionStats(n, nseg, c, DEV, need_prob=True)
    a.add_batch_lowres(0, low, spx, 0.1)
    b.add_batch(0, full, spx, 0.1)
    top2 = full.topk(2, dim=1).values
    unsafe_px = (top2[:, 0] - top2[:, 1]) <= 1e-5
    unsafe = torch.zeros((n, nseg + 1), dtype=torch.bool, device=DEV)
    for i in range(n):
        unsafe[i, spx[i][unsafe_px[i]].long()] = True
    safe = ~unsafe[:, :nseg]
    msg = f"seed {seed}: n={n} c={c} {h_in}x{w_in} -> {h}x{w} nseg={nseg} {dtype}"
    assert int(a.cls_cnt.sum()) == n * h * w, msg
    assert torch.equal(a.cls_cnt[safe], b.cls_cnt[safe]), msg
    np.testing.assert_allclose(a.cls_sum[safe].cpu().numpy(), b.cls_sum[safe].cpu().numpy(), rtol=2e-5, atol=1e-6, err_msg=msg)
    np.testing.assert_allclose(a.prob_sum.cpu().numpy(), b.prob_sum.cpu().numpy(), rtol=1e-5, atol=1e-6, err_msg=msg)


@pytest.mark.parametrize("seed", range(8))
def test_loss_step_random_masks_and_shapes(seed):
    """The whole-step entry points (active-tile list, zero sweep, one call per direction) on awkward shapes and masks --
    empty, full, a single pixel, odd widths (no 128-bit zero stores), every density around the sparse / dense switch --
    against the oracle: three losses and the dense gradient."""
    from mulactseg_b200 import losses as L
    rng = np.random.RandomState(4000 + seed)
    h, w = _rng_shape(rng, 60, 140)
    n = int(rng.randint(1, 4))
    c = int(rng.choice([3, 8, 20, 21]))
    nseg = int(rng.choice([1, 4, 30]))
    rho = float(rng.choice([0.0, 0.05, 0.3, 0.6, 1.0]))
    x = synth.logits(n, c, h, w, "cosine", seed=seed)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=seed + 1), nseg, int(rng.randint(0, 3)))
    trg = synth.multihot_targets(n, nseg, c, seed=seed + 2, p_extra=0.2, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, rho, seed=seed + 3)
    if seed == 0:
        mask[:] = False
        mask[0, h // 2, w // 2] = bool(spx[0, h // 2, w // 2] < nseg)
    xr = x.clone().requires_grad_(True)
    total_ref, parts_ref = olo.stage1_total(xr, trg, spx, mask, nseg, 0.1, 0.1)
    if torch.is_tensor(total_ref) and total_ref.requires_grad:
        total_ref.backward()
    ref_grad = xr.grad.numpy() if xr.grad is not None else np.zeros(x.shape, dtype=np.float32)
    group, multi = L.stage1_criterion(types.SimpleNamespace(nseg=nseg, group_ce_temp=0.1, multi_ce_temp=0.1), c - 1)
    xd = x.to(DEV).requires_grad_(True)
    td, sd, md = trg.to(DEV), spx.to(DEV), mask.to(DEV)
    g = group(xd, td, sd, md)
    ce, mc = multi(xd, td, sd, md)
    (16.0 * ce + 8.0 * mc + g).backward()
    msg = f"seed {seed}: n={n} c={c} {h}x{w} nseg={nseg} rho={rho}"
    as_float = lambda v: float(v.detach()) if torch.is_tensor(v) else float(v)      # noqa: E731
    np.testing.assert_allclose([ce.item(), mc.item(), g.item()], [as_float(v) for v in parts_ref], rtol=1e-5, atol=1.2e-7, err_msg=msg)
    np.testing.assert_allclose(xd.grad.cpu().numpy(), ref_grad, rtol=1e-4, atol=1e-5 * max(float(np.abs(ref_grad).max()), 1e-30) + 1e-6,
                               err_msg=msg)
