"""Parity against the CPU oracle AT the BASELINE.json shapes (not just size-independent properties):

  configs[1]  acquisition, 1024 x 2048 x 20 logits, 2048 superpixels -- every selector family, per-region scores 1e-5,
              arg-max histograms bit-exact;
  configs[2]  VOC-shaped: 513 x 513 x 22 (the reference's own crop, rows not 16-byte aligned) and 375 x 500 x 21;
  configs[3]  stage-1 losses, 4 x 20 x 768 x 768, pad id, values 1e-5 + the DENSE gradient;
  configs[4]  prototype labeller, one 1024 x 2048 image with 256-d features (and the VOC shape): labels bit-exact.

The oracle costs seconds per case at these sizes, so these run in the default ``-m gpu`` tier.
"""
import types

import numpy as np
import pytest
import torch

from helpers import assert_scores_close, batches, tie_free
from mulactseg_b200 import synth
from oracle import acquisition as oa, labeller as ol, losses as olo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _engine(method, logits, spx, nseg, temp, coeff, bs, predignore=False):
    from mulactseg_b200 import acquisition as acq
    spec = acq.SELECTORS[method]
    x = logits.to(DEV)
    ids = spx.to(DEV, torch.int32)
    if spec.slice_ignore and predignore:
        x = x[:, :-1]
    stats = acq.RegionStats(x.shape[0], nseg, x.shape[1], DEV, need_prob=spec.weighting == "predclsbal")
    for i in range(0, x.shape[0], bs):
        stats.add_batch(i, x[i:i + bs], ids[i:i + bs], temp)
    score, dom = acq.finalize(stats, spec, coeff, bs)
    torch.cuda.synchronize()
    return score.cpu().numpy(), stats


@pytest.mark.parametrize("shape", [("cityscapes", 2, 20, 1024, 2048, 2048, 1, 6.0),
                                   ("voc_crop", 5, 22, 513, 513, 150, 4, 12.0),
                                   ("voc_native", 3, 22, 375, 500, 150, 2, 12.0)])
def test_every_selector_family_matches_the_oracle_at_baseline_shapes(shape):
    name, n, c, h, w, nseg, bs, coeff = shape
    temp = 0.1
    logits = tie_free(synth.logits(n, c, h, w, "cosine", seed=21), temp, drop_last_too=True)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=22, drop_ids=2)
    pool = batches(logits, spx, bs)          # reference batching: the mean-of-batch-means has a short last batch for VOC
    hist_ref = oa.region_histograms(pool, nseg, temp).numpy()
    # Normalised selectors compute (u - min) / (max - min): a relative error e on the region means u (1e-5 allowed; the
    # oracle's own sequential fp32 sums of ~1000 pixels carry ~2e-6) becomes e * u / (max - min) ABSOLUTE on the [0, 1]
    # scale.  The region means of i.i.d. synthetic logits sit in a narrow band, so that factor is ~5-10 here; the
    # tolerance of the normalised comparisons is scaled by it (still 1e-5 relative on what the kernels compute).
    raw = torch.cat([oa._segment_mean(oa.softmax_bvsb(p, temp)[0], s, nseg) for p, s in pool]).view(-1)
    raw = raw[raw != 0]
    amp = float(raw.max() / (raw.max() - raw.min()))

    def close_normalised(got, ref):
        np.testing.assert_allclose(got.astype(np.float64), ref.astype(np.float64), rtol=1e-5, atol=1e-5 * max(amp, 1.0), err_msg=name)

    got, stats = _engine("my_bvsb_predclsbal_pwr_banignore", logits, spx, nseg, temp, coeff, bs)
    np.testing.assert_array_equal(stats.cls_cnt.cpu().numpy().astype(np.int64), hist_ref)           # integer: bit-exact
    assert_scores_close(got, oa.scores_predclsbal_pwr(pool, nseg, temp, coeff, ban_ignore=True).numpy(), False, name)

    got, _ = _engine("my_bvsb_predclsbal_pwr", logits, spx, nseg, temp, coeff, bs)
    assert_scores_close(got, oa.scores_predclsbal_pwr(pool, nseg, temp, coeff, ban_ignore=False).numpy(), False, name)

    got, _ = _engine("my_bvsb_banignore", logits, spx, nseg, temp, coeff, bs)
    close_normalised(got, oa.scores_my_bvsb_banignore(pool, nseg, temp).numpy())

    got, _ = _engine("my_bvsb_clsbal_v2_banignore", logits, spx, nseg, temp, coeff, bs)
    close_normalised(got, oa.scores_clsbal_v2(pool, nseg, temp, ban_ignore=True).numpy())

    # plain my_bvsb on a predignore net: the ignore channel is sliced off and read in place through the image stride
    got, _ = _engine("my_bvsb", logits, spx, nseg, temp, coeff, bs, predignore=True)
    raw = torch.cat([oa._segment_mean(oa.softmax_bvsb(p[:, :-1], temp)[0], s, nseg) for p, s in pool]).view(-1)
    raw = raw[raw != 0]
    amp = float(raw.max() / (raw.max() - raw.min()))
    close_normalised(got, oa.scores_my_bvsb(pool, nseg, temp, predignore=True).numpy())


def test_bf16_logits_weighted_selector_matches_the_oracle_on_tie_free_pixels():
    """bf16 logits (north_star: "stream bf16/fp32 logits"): the kernel must equal the oracle evaluated on the SAME rounded
    values.  bf16 rounding creates exact top-2 ties (the oracle's topk order is then arbitrary), so tied pixels are
    nudged apart in bf16 before both sides see them; then the weighted selector is held to 1e-5 like fp32."""
    n, c, h, w, nseg, bs = 2, 19, 512, 1024, 2048, 1
    xf = tie_free(synth.logits(n, c, h, w, "cosine", seed=5).to(torch.bfloat16), 0.1, bump=0.02, dtype=torch.bfloat16)
    x = xf.to(torch.bfloat16)
    assert torch.equal(x.float(), xf)         # the nudged values are exactly representable in bf16
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=6)
    pool = batches(xf, spx, bs)
    got, stats = _engine("my_bvsb_predclsbal_pwr", x, spx, nseg, 0.1, 6.0, bs)
    np.testing.assert_array_equal(stats.cls_cnt.cpu().numpy().astype(np.int64), oa.region_histograms(pool, nseg, 0.1).numpy())
    assert_scores_close(got, oa.scores_predclsbal_pwr(pool, nseg, 0.1, 6.0, ban_ignore=False).numpy(), False, "bf16 pwr")


@pytest.mark.parametrize("rho", [0.2, 1.0])
def test_stage1_losses_match_the_oracle_at_train_crop_size(rho):
    """4 x 20 x 768 x 768 (BASELINE configs[3] per-GPU micro-batch; the oracle's autograd over N=16 costs minutes), pad
    id in the crop border, one image with nothing selected: the three losses and the dense gradient."""
    from mulactseg_b200 import losses as L
    n, c, h, w, nseg = 4, 20, 768, 768, 2048
    x = synth.logits(n, c, h, w, "cosine", seed=31, coherent=4)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=32), nseg, 16)
    trg = synth.multihot_targets(n, nseg, c, seed=33, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, rho, seed=34)
    mask[2] = False
    xr = x.clone().requires_grad_(True)
    total_ref, (ce_ref, mc_ref, group_ref) = olo.stage1_total(xr, trg, spx, mask, nseg, 0.1, 0.1)
    total_ref.backward()
    ref_grad = xr.grad.numpy()

    group, multi = L.stage1_criterion(types.SimpleNamespace(nseg=nseg, group_ce_temp=0.1, multi_ce_temp=0.1), c - 1)
    xd = x.to(DEV).requires_grad_(True)
    td, sd, md = trg.to(DEV), spx.to(DEV), mask.to(DEV)
    g = group(xd, td, sd, md)
    ce, mc = multi(xd, td, sd, md)
    total = 16.0 * ce + 8.0 * mc + g
    total.backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose([ce.item(), mc.item(), g.item()], [ce_ref.item(), mc_ref.item(), group_ref.item()], rtol=1e-5)
    np.testing.assert_allclose(total.item(), total_ref.item(), rtol=1e-5)
    got = xd.grad.cpu().numpy()
    # 1e-4 relative + 1e-5 of the largest gradient (softmax backward subtracts nearly equal terms where P -> 1)
    np.testing.assert_allclose(got, ref_grad, rtol=1e-4, atol=1e-5 * np.abs(ref_grad).max())
    assert float(np.abs(got[2]).max()) == 0.0


@pytest.mark.parametrize("shape", [("cityscapes", 1024, 2048, 2048, 20, 256, 0.08, "median", False, ""),
                                   ("cityscapes_tile_walk", 1024, 2048, 2048, 20, 256, 0.08, "median", False, "1"),
                                   ("cityscapes_multihot_min", 1024, 2048, 2048, 20, 256, 0.3, "min", True, ""),
                                   ("voc", 375, 500, 150, 21, 256, 0.3, "median", False, ""),
                                   ("voc_superpixel_walk", 375, 500, 150, 21, 256, 0.3, "median", False, "0")])
def test_proto_labeller_matches_the_oracle_at_baseline_shapes(shape, monkeypatch):
    from mulactseg_b200 import labeller
    name, h, w, nseg, c, ch, rho, thr, only_multihot, walk = shape
    if walk:
        monkeypatch.setenv("MAS_LABELLER_TILE", walk)
    feats = synth.features(1, ch, h, w, seed=41)
    logits = synth.logits(1, c, h, w, "normal", seed=42, coherent=4)
    spx = synth.superpixel_map(1, h, w, nseg, "jitter", seed=43)
    trg = synth.multihot_targets(1, nseg, c, seed=44, p_ignore=0.0, p_extra=0.12 if only_multihot else 0.08)
    mask = synth.region_mask(spx, nseg, rho, seed=45)
    ref = ol.pseudo_label_generation(feats, logits, trg, mask, spx, only_multihot=only_multihot, threshold=thr)
    got = labeller.pseudo_label_generation(None, feats.to(DEV), logits.to(DEV), trg.to(DEV), mask.to(DEV), spx.to(DEV),
                                           only_multihot, thr).cpu()
    assert float((ref != 255).float().mean()) > 0.05
    differ = got != ref
    # labels are integers: bit-exact.  The only licence: a pixel whose two best similarities, or a
    # similarity and its threshold, agree to fp32 rounding may flip, because the inner products are summed in another
    # order than the oracle's mm.  At 2 M pixels x 256 channels that can hit a handful of pixels; none is expected.
    assert int(differ.sum()) <= 2, f"{name}: {int(differ.sum())} of {h * w} pixels differ"


def test_top_labeller_matches_the_oracle_at_cityscapes_size():
    from mulactseg_b200 import labeller
    n, c, h, w, nseg = 1, 20, 1024, 2048, 2048
    logits = synth.logits(n, c, h, w, "normal", seed=51)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=52)
    trg = synth.multihot_targets(n, nseg, c, seed=53)
    mask = synth.region_mask(spx, nseg, 0.3, seed=54)
    ref = ol.top_pseudo_label_generation(logits, trg, mask, spx)
    got = labeller.top_pseudo_label_generation(None, logits.to(DEV), trg.to(DEV), mask.to(DEV), spx.to(DEV)).cpu()
    np.testing.assert_array_equal(got.numpy(), ref.numpy())
