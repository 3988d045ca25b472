"""GPU parity of the stage-2 pseudo-labellers against the labels produced by the unmodified reference
(tests/golden/labeller.npz) and against the CPU oracle on further shapes.  Integer outputs: bit-exact."""
import os
import types

import numpy as np
import pytest
import torch

from mulactseg_b200 import synth
from oracle import labeller as ol

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
LAB = np.load(os.path.join(GOLDEN, "labeller.npz"))
DEV = "cuda:0"
LAB_KEYS = sorted({k.rsplit("/", 1)[0] for k in LAB.files if k.startswith("eval_save")})


@pytest.mark.parametrize("key", LAB_KEYS)
@pytest.mark.parametrize("id_dtype,walk", [(torch.int64, "1"), (torch.int32, "0")])
def test_proto_labeller_matches_reference_golden(key, id_dtype, walk, monkeypatch):
    """Both traversals of the propagate step (spatial tiles / one CTA per superpixel slice) against the reference's labels."""
    import importlib
    monkeypatch.setenv("MAS_LABELLER_TILE", walk)
    variant, thr, _ = key.split("/")
    mixin = importlib.import_module(f"mulactseg_b200.trainer.{variant}").LabellerMixin

    class Trainer(mixin):
        args = types.SimpleNamespace(nseg=int(LAB[f"{key}/targets"].shape[1]), cosprop_threshold_method=thr)

    feats = torch.from_numpy(LAB[f"{key}/feats"]).to(DEV)
    logits = torch.from_numpy(LAB[f"{key}/logits"]).to(DEV)
    spx = torch.from_numpy(LAB[f"{key}/spx"]).to(DEV, id_dtype)
    trg = torch.from_numpy(LAB[f"{key}/targets"]).to(DEV)
    mask = torch.from_numpy(LAB[f"{key}/mask"]).to(DEV)
    got = Trainer().pseudo_label_generation(torch.zeros_like(spx), feats, logits, trg, mask, spx)
    assert got.dtype == torch.int64
    np.testing.assert_array_equal(got.cpu().numpy(), LAB[f"{key}/plbl"].astype(np.int64))


def test_top_labeller_matches_reference_golden():
    from mulactseg_b200 import labeller
    got = labeller.top_pseudo_label_generation(None, torch.from_numpy(LAB["top/logits"]).to(DEV),
                                               torch.from_numpy(LAB["top/targets"]).to(DEV),
                                               torch.from_numpy(LAB["top/mask"]).to(DEV),
                                               torch.from_numpy(LAB["top/spx"]).to(DEV).long())
    np.testing.assert_array_equal(got.cpu().numpy(), LAB["top/plbl"].astype(np.int64))


@pytest.mark.parametrize("shape", [(40, 56, 20, 7, 32, 0.4, "jitter"), (33, 47, 12, 21, 64, 0.7, "jitter"),
                                   (24, 36, 16, 6, 16, 1.0, "grid"), (20, 24, 30, 5, 8, 0.5, "random"),
                                   (64, 96, 40, 12, 256, 0.15, "jitter")])
@pytest.mark.parametrize("only_multihot,thr,walk", [(False, "median", "0"), (True, "median", "1"), (False, "min", "1"), (False, "median", "1")])
def test_proto_labeller_matches_oracle(shape, only_multihot, thr, walk, monkeypatch):
    from mulactseg_b200 import labeller
    monkeypatch.setenv("MAS_LABELLER_TILE", walk)
    h, w, nseg, c, ch, rho, kind = shape
    feats = synth.features(1, ch, h, w, seed=h)
    logits = synth.logits(1, c, h, w, "normal", seed=w, coherent=4)
    spx = synth.superpixel_map(1, h, w, nseg, kind, seed=nseg)
    trg = synth.multihot_targets(1, nseg, c, seed=c, p_extra=0.3)
    mask = synth.region_mask(spx, nseg, rho, seed=ch)
    ref = ol.pseudo_label_generation(feats, logits, trg, mask, spx, only_multihot=only_multihot, threshold=thr)
    got = labeller.pseudo_label_generation(None, feats.to(DEV), logits.to(DEV), trg.to(DEV), mask.to(DEV), spx.to(DEV),
                                           only_multihot, thr)
    got = got.cpu()
    mismatch = int((got != ref).sum())
    # fp32 inner products are summed in a different order than the oracle's mm: a pixel may flip only where two
    # similarities (or a similarity and its threshold) agree to rounding -- there are none in these seeded cases
    assert mismatch == 0, f"{mismatch} of {h * w} pixels differ"


@pytest.mark.parametrize("shape", [(64, 96, 40, 12, 64, 0.3, 4), (40, 56, 20, 7, 32, 0.5, 4), (129, 129, 30, 21, 48, 0.4, 0)])
@pytest.mark.parametrize("source", ["lowres_f32", "lowres_bf16", "full_bf16"])
def test_feature_sources_match_oracle(shape, source):
    """mas_proto_labeller_src_dev: bf16 features, and the head's LOW-RESOLUTION features interpolated inside the kernels
     -- labels must equal the oracle run on what the reference would have been handed: the
    F.interpolate(..., mode='bilinear', align_corners=False) of the same map, in fp32 (models/segmentation/utils.py:28-34)."""
    from mulactseg_b200 import labeller
    h, w, nseg, c, ch, rho, stride = shape
    g = torch.Generator().manual_seed(h * w + ch)
    hl, wl = (h // stride, w // stride) if stride else (33, 33)          # 33 -> 129: the VOC ratio (513 = 129 * 3.977)
    low = torch.nn.functional.normalize(torch.randn((1, ch, hl, wl), generator=g), dim=1)
    if source == "lowres_bf16":
        low = low.to(torch.bfloat16)
    full = torch.nn.functional.interpolate(low.float(), size=(h, w), mode="bilinear", align_corners=False)
    if source == "full_bf16":
        full = full.to(torch.bfloat16)
    logits = synth.logits(1, c, h, w, "normal", seed=w, coherent=4)
    spx = synth.superpixel_map(1, h, w, nseg, "jitter", seed=nseg)
    trg = synth.multihot_targets(1, nseg, c, seed=c, p_extra=0.3)
    mask = synth.region_mask(spx, nseg, rho, seed=ch)
    ref = ol.pseudo_label_generation(full.float(), logits, trg, mask, spx)
    feats = full if source == "full_bf16" else low
    got = labeller.pseudo_label_generation(None, feats.to(DEV), logits.to(DEV), trg.to(DEV), mask.to(DEV), spx.to(DEV)).cpu()
    # low-resolution sources: the interpolation is evaluated with the expression of torch's CUDA kernel, the oracle's
    # features come from torch's CPU kernel -- an ulp apart at most, which can flip a pixel only where two similarities
    # (or a similarity and its threshold) agree to rounding
    assert int((got != ref).sum()) <= (0 if source == "full_bf16" else 1)
    assert float((ref != 255).float().mean()) > 0.05


def test_voc_multiscale_variant_matches_oracle():
    """..._includeonehot_voc_ms.py: features averaged over scales (+ flipped copies) and re-normalised, then the same
    pseudo_label_generation; VOC-shaped (21 classes, ~150 superpixels)."""
    import types
    from mulactseg_b200.trainer import eval_save_cosplbl_prop_includeonehot_voc_ms as ms
    h, w, nseg, c, ch = 60, 84, 24, 21, 32
    scales = [(30, 42), (60, 84), (90, 126)]
    feat_list, out_list = [], []
    for k, (hk, wk) in enumerate(scales + scales):              # second half: computed on flipped images
        f = synth.features(1, ch, hk * 4, wk * 4, seed=10 + k)[:, :, ::4, ::4].contiguous()
        o = synth.logits(1, c, hk, wk, "normal", seed=20 + k, coherent=2)
        feat_list.append(f)
        out_list.append(o)
    feats, logits = ms.fuse_multiscale(feat_list, out_list, (h, w))
    assert feats.shape == (1, ch, h, w) and logits.shape == (1, c, h, w)
    np.testing.assert_allclose(feats.norm(dim=1).numpy(), 1.0, rtol=1e-5)
    spx = synth.superpixel_map(1, h, w, nseg, "jitter", seed=3)
    trg = synth.multihot_targets(1, nseg, c, seed=4, p_extra=0.15)
    mask = synth.region_mask(spx, nseg, 0.4, seed=5)
    ref = ol.pseudo_label_generation(feats, logits, trg, mask, spx, only_multihot=False, threshold="median")

    class Trainer(ms.LabellerMixin):
        args = types.SimpleNamespace(nseg=nseg, cosprop_threshold_method="median")

    got = Trainer().pseudo_label_generation(torch.zeros_like(spx), feats.to(DEV), logits.to(DEV), trg.to(DEV), mask.to(DEV),
                                            spx.to(DEV)).cpu()
    assert int((got != ref).sum()) == 0


def test_top_labeller_matches_oracle_and_quirk():
    from mulactseg_b200 import labeller
    n, c, h, w, nseg = 3, 9, 31, 45, 14
    logits = synth.logits(n, c, h, w, "normal", seed=1)
    spx = synth.pad_border(synth.superpixel_map(n, h, w, nseg, "jitter", seed=2), nseg, 2)
    trg = synth.multihot_targets(n, nseg, c, seed=3, p_extra=0.2)
    mask = synth.region_mask(spx, nseg, 0.6, seed=4)
    ref = ol.top_pseudo_label_generation(logits, trg, mask, spx.clamp(max=nseg - 1))
    got = labeller.top_pseudo_label_generation(None, logits.to(DEV), trg.to(DEV), mask.to(DEV), spx.to(DEV)).cpu()
    np.testing.assert_array_equal(got.numpy(), ref.numpy())
    # all candidate logits negative: a NON-candidate class (value 0) wins, exactly like the reference's multiply-then-max
    neg = -logits.abs() - 0.1
    ref = ol.top_pseudo_label_generation(neg, trg, mask, spx.clamp(max=nseg - 1))
    got = labeller.top_pseudo_label_generation(None, neg.to(DEV), trg.to(DEV), mask.to(DEV), spx.to(DEV)).cpu()
    np.testing.assert_array_equal(got.numpy(), ref.numpy())


def test_selected_superpixel_without_candidates_raises():
    from mulactseg_b200 import labeller
    h, w, nseg, c, ch = 16, 24, 6, 5, 8
    feats = synth.features(1, ch, h, w, seed=1).to(DEV)
    logits = synth.logits(1, c, h, w, "normal", seed=2).to(DEV)
    spx = synth.superpixel_map(1, h, w, nseg, "grid", seed=3).to(DEV)
    trg = synth.multihot_targets(1, nseg, c, seed=4).to(DEV)
    trg[0, 2] = 0
    mask = torch.ones((1, h, w), dtype=torch.bool, device=DEV)
    with pytest.raises(RuntimeError):
        labeller.pseudo_label_generation(None, feats, logits, trg, mask, spx)
    with pytest.raises(RuntimeError):
        labeller.pseudo_label_generation(None, feats.cpu(), logits.cpu(), trg.cpu(), mask.cpu(), spx.cpu())


def test_full_size_properties():
    """Cityscapes-shaped image (BASELINE config 5): invariants that hold whatever the data."""
    from mulactseg_b200 import labeller
    h, w, nseg, c, ch = 1024, 2048, 2048, 20, 256
    feats = synth.features(1, ch, h, w, seed=1, device=DEV)
    logits = synth.logits(1, c, h, w, "normal", seed=2, device=DEV, coherent=4)
    spx = synth.superpixel_map(1, h, w, nseg, "jitter", seed=3, device=DEV)
    trg = synth.multihot_targets(1, nseg, c, seed=4, device=DEV, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, 0.08, seed=5)
    out = labeller.pseudo_label_generation(None, feats, logits, trg, mask, spx)
    torch.cuda.synchronize()
    assert out.shape == (1, h, w)
    rows = trg[0][spx[0]]                                   # (h,w,c) candidate sets per pixel
    sel = mask[0]
    # selected pixels: labelled, and always with a candidate class of their own superpixel
    lab_sel = out[0][sel]
    assert int((lab_sel == 255).sum()) == 0
    assert bool(rows[sel].gather(1, lab_sel.view(-1, 1)).all())
    # unselected pixels are labelled only inside superpixels that touch a selected one (3x3 dilation)
    chosen = torch.zeros(nseg, dtype=torch.bool, device=DEV)
    chosen[spx[0][sel]] = True
    near = torch.nn.functional.max_pool2d(chosen[spx[0]].float()[None, None], 3, 1, 1)[0, 0] > 0
    touched = torch.zeros(nseg, dtype=torch.bool, device=DEV)
    touched[spx[0][near]] = True
    labelled_unsel = (out[0] != 255) & ~sel
    assert bool(touched[spx[0][labelled_unsel]].all())
    # idempotent / deterministic
    again = labeller.pseudo_label_generation(None, feats, logits, trg, mask, spx)
    assert torch.equal(out, again)


def test_batched_call_runs_images_on_side_streams_and_equals_single_image_calls():
    """A loader batch is labelled image by image on up to four side streams (one workspace per stream): same labels as
    one call per image on the caller's stream, for more images than streams, and the result is ordered after the join."""
    from mulactseg_b200 import labeller
    n, c, h, w, nseg, f = 6, 8, 40, 72, 30, 16
    feats = synth.features(n, f, h, w, seed=1, device=DEV)
    logits = synth.logits(n, c, h, w, "normal", seed=2, device=DEV)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=3, device=DEV)
    trg = synth.multihot_targets(n, nseg, c, seed=4, device=DEV, p_extra=0.2, p_ignore=0.0)
    mask = synth.region_mask(spx, nseg, 0.3, seed=5)
    for _ in range(3):
        batch = labeller.pseudo_label_generation(None, feats, logits, trg, mask, spx)
        for i in range(n):
            one = labeller.pseudo_label_generation(None, feats[i:i + 1], logits[i:i + 1], trg[i:i + 1], mask[i:i + 1], spx[i:i + 1])
            assert torch.equal(batch[i], one[0]), i
