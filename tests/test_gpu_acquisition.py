"""GPU parity of the acquisition path (kernels called through the C ABI) against the CPU oracle and
against the vectors produced by the unmodified reference (tests/golden/acquisition.npz)."""
import importlib
import os

import numpy as np
import pytest
import torch

from helpers import PoolSet, assert_scores_close, batches, fake_trainer, selector_args, tie_free
from mulactseg_b200 import synth
from oracle import acquisition as oa

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
ACQ = np.load(os.path.join(GOLDEN, "acquisition.npz"))
DEV = "cuda:0"

METHODS = ["my_bvsb", "my_bvsb_banignore", "my_bvsb_predclsbal_pwr", "my_bvsb_predclsbal_pwr_banignore",
           "my_bvsb_clsbal_v2", "my_bvsb_clsbal_v2_banignore"]


def _engine_scores(method, logits, spx, nseg, temp, coeff, bs, predignore=False, chunk=None):
    from mulactseg_b200 import acquisition as acq
    spec = acq.SELECTORS[method]
    x = logits.to(DEV)
    ids = spx.to(DEV, torch.int32)
    if spec.slice_ignore and predignore:
        x = x[:, :-1]
    stats = acq.RegionStats(x.shape[0], nseg, x.shape[1], DEV, need_prob=spec.weighting == "predclsbal")
    chunk = chunk or bs
    for i in range(0, x.shape[0], chunk):
        stats.add_batch(i, x[i:i + chunk], ids[i:i + chunk], temp)
    score, dom = acq.finalize(stats, spec, coeff, bs)
    torch.cuda.synchronize()
    return score.cpu(), dom.cpu(), stats


@pytest.mark.parametrize("case", ["city_small", "voc_small", "adversarial"])
def test_selectors_match_reference_golden(case):
    logits = torch.from_numpy(ACQ[f"{case}/logits"])
    spx = torch.from_numpy(ACQ[f"{case}/spx"])
    nseg, bs = (int(v) for v in ACQ[f"{case}/meta"])
    temp, coeff = (float(v) for v in ACQ[f"{case}/temp_coeff"])
    for method in METHODS:
        score, _, stats = _engine_scores(method, logits, spx, nseg, temp, coeff, bs, chunk=3)
        assert_scores_close(score.numpy(), ACQ[f"{case}/{method}"], "pwr" not in method, f"{case}/{method}")
        if method == "my_bvsb_banignore":  # integer histogram: bit-exact
            np.testing.assert_array_equal(stats.cls_cnt.cpu().numpy().astype(np.int64), ACQ[f"{case}/hist"])
    score, _, _ = _engine_scores("my_bvsb", logits, spx, nseg, temp, coeff, bs, predignore=True)
    assert_scores_close(score.numpy(), ACQ[f"{case}/my_bvsb/predignore"], True, f"{case}/my_bvsb/predignore")


@pytest.mark.parametrize("shape", [(3, 20, 128, 256, 512, "jitter"), (2, 21, 65, 77, 40, "jitter"),
                                   (2, 19, 64, 128, 2048, "random"), (1, 5, 33, 36, 7, "grid"),
                                   (2, 32, 40, 64, 30, "jitter"), (1, 2, 16, 32, 4, "grid"),
                                   (5, 20, 24, 200, 64, "jitter"), (3, 22, 37, 132, 150, "jitter"),
                                   (9, 19, 8, 260, 16, "grid")])
@pytest.mark.parametrize("method", ["my_bvsb_predclsbal_pwr_banignore", "my_bvsb_clsbal_v2"])
@pytest.mark.parametrize("path", ["tma", "ldg", "abreast"])
def test_selectors_match_oracle(shape, method, path, monkeypatch):
    """Every data path of the scorer: TMA ring where rows are 16-byte aligned, the abreast kernel otherwise (or when
    forced), the plain register path when forced ("tma" = the default choice for the shape)."""
    monkeypatch.setenv("MAS_SCORER_PATH", path)
    n, c, h, w, nseg, kind = shape
    logits = synth.logits(n, c, h, w, "cosine", seed=n * c + h)
    spx = synth.superpixel_map(n, h, w, nseg, kind, seed=7, drop_ids=1 if nseg > 4 else 0)
    pool = batches(logits, spx, 2)
    if "pwr" in method:
        ref = oa.scores_predclsbal_pwr(pool, nseg, 0.1, 6.0, ban_ignore=True)
    else:
        ref = oa.scores_clsbal_v2(pool, nseg, 0.1, ban_ignore=False)
    score, _, stats = _engine_scores(method, logits, spx, nseg, 0.1, 6.0, 2)
    assert_scores_close(score.numpy(), ref.numpy(), "pwr" not in method, str(shape))
    hist = oa.region_histograms(pool, nseg, 0.1)
    np.testing.assert_array_equal(stats.cls_cnt.cpu().numpy().astype(np.int64), hist.numpy())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_unaligned_rows_and_unaligned_batch_pointers(dtype):
    """Odd width and odd plane size (every plane / row has its own alignment phase) with batches that are single-image SLICES
    of one allocation, so their base pointers are only element-aligned: the abreast kernel (scalar loads) takes them."""
    from mulactseg_b200 import acquisition as acq
    n, c, h, w, nseg = 5, 22, 37, 41, 12
    logits = tie_free(synth.logits(n, c, h, w, "cosine", seed=8), 0.1, bump=0.05, dtype=dtype if dtype != torch.float32 else None)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=9)
    x, ids = logits.to(DEV, dtype), spx.to(DEV, torch.int32)
    assert x[1:2].data_ptr() % 16 != 0 or x[3:4].data_ptr() % 16 != 0
    spec = acq.SELECTORS["my_bvsb_predclsbal_pwr_banignore"]
    stats = acq.RegionStats(n, nseg, c, DEV, need_prob=True)
    for i in range(n):
        stats.add_batch(i, x[i:i + 1], ids[i:i + 1], 0.1)
    score, _ = acq.finalize(stats, spec, 12.0, 1)
    pool = batches(logits.float(), spx, 1)
    ref = oa.scores_predclsbal_pwr(pool, nseg, 0.1, 12.0, ban_ignore=True)
    np.testing.assert_array_equal(stats.cls_cnt.cpu().numpy().astype(np.int64), oa.region_histograms(pool, nseg, 0.1).numpy())
    assert_scores_close(score.cpu().numpy(), ref.numpy(), False, str(dtype))


def test_bf16_logits_match_oracle_on_rounded_inputs():
    n, c, h, w, nseg = 2, 20, 64, 128, 64
    logits = synth.logits(n, c, h, w, "cosine", seed=3).to(torch.bfloat16)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=4)
    # bf16 rounding creates exact top-2 ties; keep pixels without them by nudging the arg-max plane
    ref_in = logits.float()
    top2 = ref_in.topk(2, dim=1).values
    tie = top2[:, 0] == top2[:, 1]
    assert tie.float().mean() < 0.2
    ref = oa.scores_predclsbal_pwr(batches(ref_in, spx, 2), nseg, 0.1, 6.0, ban_ignore=False)
    score, _, _ = _engine_scores("my_bvsb_predclsbal_pwr", logits, spx, nseg, 0.1, 6.0, 2)
    # ties make top1 ambiguous in the oracle (topk order), so compare the tie-free statistic: region means
    ref_plain = oa.scores_my_bvsb(batches(ref_in, spx, 2), nseg, 0.1, predignore=False)
    plain, _, _ = _engine_scores("my_bvsb", logits, spx, nseg, 0.1, 6.0, 2)
    assert_scores_close(plain.numpy(), ref_plain.numpy(), True, "bf16 my_bvsb")
    assert np.isfinite(score.numpy()).all() and ref.shape == score.shape


@pytest.mark.parametrize("path,stages,warps", [("tma", "", ""), ("tma", "3", "5"), ("ldg", "", "")])
def test_full_size_properties(path, stages, warps, monkeypatch):
    """Cityscapes-shaped images: size-independent invariants instead of the (slow) oracle."""
    monkeypatch.setenv("MAS_SCORER_PATH", path)
    monkeypatch.setenv("MAS_SCORER_STAGES", stages)
    monkeypatch.setenv("MAS_SCORER_WARPS", warps)
    n, c, h, w, nseg = 2, 20, 1024, 2048, 2048
    logits = synth.logits(n, c, h, w, "cosine", seed=1, device=DEV)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=2, device=DEV, dtype=torch.int32)
    from mulactseg_b200 import acquisition as acq
    stats = acq.RegionStats(n, nseg, c, DEV, need_prob=True)
    stats.add_batch(0, logits, spx, 0.1)
    torch.cuda.synchronize()
    cnt = stats.cls_cnt.long()
    # every pixel lands in exactly one (superpixel, class) bin
    assert cnt.sum(dim=(1, 2)).tolist() == [h * w] * n
    # superpixel sizes == bincount of the id map (bit-exact)
    for i in range(n):
        assert torch.equal(cnt[i].sum(dim=1), torch.bincount(spx[i].reshape(-1).long(), minlength=nseg))
    # class marginals == histogram of the arg-max plane (bit-exact away from exact logit ties)
    top1 = logits.argmax(dim=1)
    for i in range(n):
        assert torch.equal(cnt[i].sum(dim=0), torch.bincount(top1[i].reshape(-1), minlength=c))
    # the softmax probabilities of a pixel sum to one
    np.testing.assert_allclose(stats.prob_sum.sum(dim=1).cpu().numpy(), [h * w] * n, rtol=1e-6)
    # total bvsb mass is conserved by the segmented reduction
    top2 = logits.topk(2, dim=1).values
    bvsb = torch.exp((top2[:, 1] - top2[:, 0]).double() / 0.1) + 1e-8
    np.testing.assert_allclose(stats.cls_sum.double().sum(dim=(1, 2)).cpu().numpy(),
                               bvsb.sum(dim=(1, 2)).cpu().numpy(), rtol=1e-5)
    # linearity in the id map: merging superpixels pairwise adds their tables
    stats2 = acq.RegionStats(n, nseg // 2, c, DEV, need_prob=False)
    stats2.add_batch(0, logits, (spx // 2).contiguous(), 0.1)
    torch.cuda.synchronize()
    assert torch.equal(stats2.cls_cnt, stats.cls_cnt.view(n, nseg // 2, 2, c).sum(dim=2).int())


@pytest.mark.parametrize("shape", [(4, 22, 513, 513, 150), (3, 21, 375, 500, 150)])
def test_voc_size_properties(shape):
    """BASELINE config 3 shapes (odd width -> scalar loads; 500 x 375 native): the same size-independent invariants."""
    from mulactseg_b200 import acquisition as acq
    n, c, h, w, nseg = shape
    logits = synth.logits(n, c, h, w, "cosine", seed=3, device=DEV)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=4, device=DEV, dtype=torch.int32, drop_ids=3)
    stats = acq.RegionStats(n, nseg, c, DEV, need_prob=True)
    stats.add_batch(0, logits[:3], spx[:3], 0.1)               # a batch of 3 and, when there is one, a short last batch
    if n > 3:
        stats.add_batch(3, logits[3:], spx[3:], 0.1)
    cnt = stats.cls_cnt.long()
    assert cnt.sum(dim=(1, 2)).tolist() == [h * w] * n
    top1 = logits.argmax(dim=1)
    for i in range(n):
        assert torch.equal(cnt[i].sum(dim=1), torch.bincount(spx[i].reshape(-1).long(), minlength=nseg))
        assert torch.equal(cnt[i].sum(dim=0), torch.bincount(top1[i].reshape(-1), minlength=c))
    np.testing.assert_allclose(stats.prob_sum.sum(dim=1).cpu().numpy(), [h * w] * n, rtol=1e-6)
    top2 = logits.topk(2, dim=1).values
    bvsb = torch.exp((top2[:, 1] - top2[:, 0]).double() / 0.1) + 1e-8
    np.testing.assert_allclose(stats.cls_sum.double().sum(dim=(1, 2)).cpu().numpy(), bvsb.sum(dim=(1, 2)).cpu().numpy(), rtol=1e-5)
    # the ban-ignore selector end to end on this shape: banned regions are exactly those dominated by the last channel
    score, dom = acq.finalize(stats, acq.SELECTORS["my_bvsb_predclsbal_pwr_banignore"], 12.0, 4)
    dominant = cnt.argmax(dim=2)
    assert torch.equal(dom.long(), dominant)
    assert float(score[dominant == c - 1].abs().max() if (dominant == c - 1).any() else 0.0) == 0.0


@pytest.mark.parametrize("path", ["tma", "ldg", "abreast"])
@pytest.mark.parametrize("shape", [(11, 19, 24, 128, 40, 1), (21, 22, 37, 132, 150, 2), (8, 6, 9, 33, 7, 3), (70, 5, 8, 64, 6, 1)])
def test_grouped_launches_equal_single_launches(shape, path, monkeypatch):
    """Several loader batches (separate allocations, short last batch, more batches than one launch takes) folded by
    grouped launches == one launch per batch: identical histograms, sums equal up to the order of the fp32 atomics."""
    from mulactseg_b200 import acquisition as acq
    monkeypatch.setenv("MAS_SCORER_PATH", path)
    n, c, h, w, nseg, bs = shape
    logits = synth.logits(n, c, h, w, "cosine", seed=n)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=n + 1, dtype=torch.int32)
    parts = [(logits[i:i + bs].clone().to(DEV), spx[i:i + bs].clone().to(DEV)) for i in range(0, n, bs)]   # own allocations
    one = acq.RegionStats(n, nseg, c, DEV, need_prob=True, group_bytes=0)
    grouped = acq.RegionStats(n, nseg, c, DEV, need_prob=True, group_bytes=1 << 40)
    first = 0
    for x, ids in parts:
        one.add_batch(first, x, ids, 0.1)
        grouped.add_batch(first, x, ids, 0.1)
        first += x.shape[0]
    from mulactseg_b200 import _lib
    per = _lib.MAS_MAX_SEGMENTS
    assert one.launches == len(parts) and grouped.launches == len(parts) // per   # only full groups went out so far ...
    assert torch.equal(one.cls_cnt, grouped.cls_cnt)                     # ... reading a table flushes the rest
    assert grouped.launches == -(-len(parts) // per)
    np.testing.assert_allclose(grouped.cls_sum.cpu().numpy(), one.cls_sum.cpu().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(grouped.prob_sum.cpu().numpy(), one.prob_sum.cpu().numpy(), rtol=1e-6)   # fp32 per-thread partials
    # a gap in the image rows or another temperature starts a new launch instead of corrupting the group
    odd = acq.RegionStats(n, nseg, c, DEV, need_prob=False, group_bytes=1 << 40)
    odd.add_batch(0, parts[0][0], parts[0][1], 0.1)
    odd.add_batch(2 * bs, parts[2][0], parts[2][1], 0.1)
    odd.add_batch(bs, parts[1][0], parts[1][1], 0.5)
    assert odd.launches == 2
    ref = acq.RegionStats(n, nseg, c, DEV, need_prob=False, group_bytes=0)
    ref.add_batch(0, parts[0][0], parts[0][1], 0.1)
    ref.add_batch(2 * bs, parts[2][0], parts[2][1], 0.1)
    ref.add_batch(bs, parts[1][0], parts[1][1], 0.5)
    assert torch.equal(odd.cls_cnt, ref.cls_cnt)


@pytest.mark.parametrize("n,c,bs", [(1, 2, 4), (7, 19, 4), (2975, 20, 4), (1031, 22, 3), (300, 32, 1)])
def test_class_weights_kernel_matches_the_reference_sequence(n, c, bs):
    """mas_class_weights_dev == `cumulated += batch mean` in loader order in fp32 (my_bvsb_predclsbal_pwr.py:36-47),
    short last batch and more batches than one shared-memory chunk included."""
    from mulactseg_b200 import ops
    g = torch.Generator().manual_seed(n * c)
    pixels = 1000
    prob = torch.rand((n, c), generator=g, dtype=torch.float64) * pixels / c
    cumulated = torch.zeros(c, dtype=torch.float32)
    n_batches = 0
    for b0 in range(0, n, bs):
        part = prob[b0:b0 + bs]
        cumulated += (part.sum(dim=0) / (part.shape[0] * pixels)).to(torch.float32)
        n_batches += 1
    want = (6.0 * (cumulated / n_batches) + 1.0) ** (-2)
    got = ops.class_weights(prob.to(DEV), pixels, bs, 6.0).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-6)    # 1 / (x * x) vs torch pow(-2): an ulp or two


def test_ids_outside_range_are_ignored_and_empty_batch():
    from mulactseg_b200 import acquisition as acq
    n, c, h, w, nseg = 1, 6, 16, 32, 5
    logits = synth.logits(n, c, h, w, "normal", seed=9, device=DEV)
    spx = synth.superpixel_map(n, h, w, nseg, "grid", seed=1, device=DEV, dtype=torch.int32)
    spx[:, :3] = nseg          # crop-padding id
    spx[:, -1] = -1
    stats = acq.RegionStats(n, nseg, c, DEV, need_prob=False)
    stats.add_batch(0, logits, spx, 1.0)
    torch.cuda.synchronize()
    assert int(stats.cls_cnt.sum()) == (h - 4) * w
    stats.add_batch(0, logits[:0], spx[:0], 1.0)  # empty batch is a no-op


@pytest.mark.parametrize("path", ["tma", "ldg", "abreast"])
@pytest.mark.parametrize("bad", [-1, -7, 2 ** 31 - 1, "nseg"])
def test_invalid_ids_before_valid_ones_do_not_leak(path, bad, monkeypatch):
    """A thread that meets out-of-range ids BEFORE its first valid superpixel (top rows of -1 / pad / garbage) must not
    carry anything into that superpixel: tables == the tables of the same image with those rows cut off."""
    from mulactseg_b200 import acquisition as acq
    monkeypatch.setenv("MAS_SCORER_PATH", path)
    n, c, h, w, nseg = 2, 7, 40, 128, 12
    logits = synth.logits(n, c, h, w, "cosine", seed=4, device=DEV)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=5, device=DEV, dtype=torch.int32)
    dirty = spx.clone()
    dirty[:, :5] = nseg if bad == "nseg" else bad
    dirty[:, 20:22, ::3] = nseg if bad == "nseg" else bad          # and some in the middle of a column
    got = acq.RegionStats(n, nseg, c, DEV, need_prob=False, group_bytes=0)
    got.add_batch(0, logits, dirty, 0.1)
    valid = (dirty >= 0) & (dirty < nseg)
    top1 = logits.argmax(dim=1)
    want = torch.zeros((n, nseg, c), dtype=torch.int64, device=DEV)
    for i in range(n):
        key = (dirty[i][valid[i]].long() * c + top1[i][valid[i]])
        want[i] = torch.bincount(key, minlength=nseg * c).view(nseg, c)
    assert torch.equal(got.cls_cnt.long(), want)
    top2 = logits.topk(2, dim=1).values
    bvsb = torch.exp((top2[:, 1] - top2[:, 0]) / 0.1) + 1e-8
    np.testing.assert_allclose(float(got.cls_sum.double().sum()), float(bvsb[valid].double().sum()), rtol=1e-5)


def test_argument_errors_raise():
    from mulactseg_b200 import acquisition as acq, ops
    stats = acq.RegionStats(1, 4, 6, DEV, need_prob=False)
    with pytest.raises(RuntimeError):
        ops.bvsb_segment_stats(torch.zeros(1, 6, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int32), 4, 1.0,
                               stats.cls_sum, stats.cls_cnt, None)      # CPU tensors: no CPU path
    with pytest.raises(RuntimeError):
        stats.add_batch(0, torch.zeros(1, 6, 4, 4, device=DEV), torch.zeros(1, 4, 4, dtype=torch.int32, device=DEV), 0.0)
    with pytest.raises(RuntimeError):
        big = acq.RegionStats(1, 4, 40, DEV, need_prob=False)            # > MAS_MAX_CLASSES channels
        big.add_batch(0, torch.zeros(1, 40, 4, 4, device=DEV), torch.zeros(1, 4, 4, dtype=torch.int32, device=DEV), 1.0)


@pytest.mark.parametrize("n,k", [(1000, 10), (5000, 5000), (300000, 100001), (70000, 1), (50, 80)])
def test_topk_and_sort_match_numpy(n, k):
    from mulactseg_b200 import ops
    rng = np.random.RandomState(n + k)
    keys = rng.randint(1, 2 ** 62, size=n, dtype=np.int64)
    keys[rng.rand(n) < 0.3] = 0                              # "not in pool"
    keys[: n // 4] = (keys[: n // 4] & 0xFFFFFFFF) | (0x3F800000 << 32)  # many equal scores, distinct ties
    keys = np.unique(keys[keys != 0])
    rng.shuffle(keys)
    full = np.concatenate([keys, np.zeros(n // 3, dtype=np.int64)])
    out, count = ops.topk_keys(torch.from_numpy(full).to(DEV), k, sort=True)
    torch.cuda.synchronize()
    want = np.sort(keys.view(np.uint64))[::-1][:k]
    assert int(count) == len(want)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint64)[: len(want)], want)


@pytest.mark.parametrize("n,k,ties", [(10, 3, False), (5000, 5000, False), (70001, 4097, False), (70001, 4097, True),
                                      (300000, 100001, False), (20000, 30000, False), (300000, 1000, True)])
def test_fast_topk_sorted_matches_numpy(n, k, ties):
    """Bucket + compaction + sort path (mas_topk_sorted_u64_dev): identical to numpy wherever it reports success; on
    heavily tied scores it must either succeed or report the overflow (-1) that sends the caller to the exact path."""
    from mulactseg_b200 import ops
    g = torch.Generator().manual_seed(n + k)
    if ties:
        hi = torch.randint(5, 9, (n,), generator=g, dtype=torch.int64) << 40      # 4 distinct high parts: one bucket
    else:
        hi = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64) << 32
    keys = (hi | torch.randperm(n, generator=g)) * (torch.rand(n, generator=g) < 0.9)
    got, count = ops.topk_sorted(keys.to(DEV), k)
    cnt = int(count.item())
    nz = np.sort(keys.numpy().astype(np.uint64)[keys.numpy() != 0])[::-1]
    if cnt < 0:
        assert ties and len(nz) > ops.sort_capacity(k)      # only an overflowing bucket may give up
        got, count = ops.topk_keys(keys.to(DEV), k, sort=True)
        cnt = int(count.item())
    assert cnt == min(k, len(nz))
    np.testing.assert_array_equal(got.cpu().numpy().astype(np.uint64)[:cnt], nz[:cnt])


@pytest.mark.parametrize("method,predignore", [("my_bvsb", True), ("my_bvsb_predclsbal_pwr_banignore", True),
                                              ("my_bvsb_clsbal_v2", False)])
def test_plugin_select_next_batch_matches_oracle(method, predignore, tmp_path):
    """Drop-in plugin API end to end: same sorted prefix and the same datalist as the reference semantics."""
    n, c, h, w, nseg, bs = 7, 8, 48, 64, 24, 2
    logits = synth.logits(n, c, h, w, "cosine", seed=5)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=6, drop_ids=1)
    im_idx, suppix = synth.pool_lists(n, nseg, spx, labelled_frac=0.2, seed=8)
    num_classes = c - 1 if predignore else c
    args = selector_args(method, nseg, num_classes, predignore, 0.1, 6.0, bs)
    mod = importlib.import_module(f"mulactseg_b200.active_selection.{method}")
    selector = mod.RegionSelector(args)
    pool = PoolSet(logits, spx.long(), im_idx, suppix)
    got = selector.calculate_scores(fake_trainer(DEV), pool)

    pb = batches(logits, spx.long(), bs)
    if method == "my_bvsb":
        ref_t = oa.scores_my_bvsb(pb, nseg, 0.1, predignore)
    elif "pwr" in method:
        ref_t = oa.scores_predclsbal_pwr(pb, nseg, 0.1, 6.0, ban_ignore=True)
    else:
        ref_t = oa.scores_clsbal_v2(pb, nseg, 0.1, ban_ignore=False)
    ref = oa.score_list(im_idx, suppix, ref_t)
    assert [(p, i) for _, p, i in got] == [(p, i) for _, p, i in ref]
    assert_scores_close([s for s, _, _ in got], [s for s, _, _ in ref], "pwr" not in method)

    # ranking + selection: compare with sorted() of OUR scores (the order is exact given the scores) ...
    import types
    budget = 20
    label = types.SimpleNamespace(im_idx=[], suppix={})
    pool_ds = PoolSet(logits, spx.long(), im_idx, suppix)
    picked = {}

    class ActiveSet:
        args = types.SimpleNamespace(fair_counting=False, or_labeling=False)
        trg_pool_dataset = pool_ds
        trg_label_dataset = label

        def expand_training_set(self, ranked, count, name):
            picked["n"] = oa.expand_training_set(ranked, count, label.im_idx, label.suppix, pool_ds.im_idx, pool_ds.suppix)
            picked["ranked"] = ranked

    selector.select_next_batch(fake_trainer(DEV), ActiveSet(), budget)
    assert picked["n"] == budget + 1
    want = sorted(got, reverse=True)[: budget + 1]
    assert [(p, i) for _, p, i in picked["ranked"][: budget + 1]] == [(p, i) for _, p, i in want]
    # ... and with the oracle's ranking wherever adjacent oracle scores are separated by more than the tolerance
    ref_sorted = sorted(ref, reverse=True)[: budget + 2]
    gaps = np.diff([s for s, _, _ in ref_sorted])
    if np.all(np.abs(gaps) > 1e-4):
        assert [(p, i) for _, p, i in picked["ranked"][: budget + 1]] == [(p, i) for _, p, i in ref_sorted[: budget + 1]]


@pytest.mark.parametrize("n,budget,zero_frac", [(1000, 300, 0.0), (5000, 20000, 0.0), (100001, 100000, 0.0), (3000, 50, 0.9),
                                               (10, 0, 0.0), (1, 5, 0.0)])
def test_prefix_cut_matches_numpy(n, budget, zero_frac):
    """mas_prefix_cut_dev == the strict '>' walk of expand_training_set (numpy: selection.cumulative_cut)."""
    from mulactseg_b200 import ops, selection
    rng = np.random.RandomState(n + budget)
    n_cost = 4 * n + 7
    table = rng.randint(1, 5, size=n_cost).astype(np.uint8)
    table[rng.rand(n_cost) < zero_frac] = 0
    ties = rng.permutation(n_cost)[:n].astype(np.uint64)
    keys = ((np.arange(n, 0, -1).astype(np.uint64) << np.uint64(32)) | ties).view(np.int64)     # already descending
    buf = torch.zeros(ops.sort_capacity(n), dtype=torch.int64, device=DEV)
    buf[:n] = torch.from_numpy(keys).to(DEV)
    got = int(ops.prefix_cut(buf, torch.tensor([n], dtype=torch.int32, device=DEV), torch.from_numpy(table).to(DEV), budget).item())
    assert got == selection.cumulative_cut(table[ties.astype(np.int64)], budget)
    assert int(ops.prefix_cut(buf, torch.tensor([-1], dtype=torch.int32, device=DEV), torch.from_numpy(table).to(DEV), budget).item()) == -1


@pytest.mark.parametrize("zero_cost_frac,budget", [(0.0, 30), (0.6, 25), (1.0, 10)])
def test_plugin_fair_counting_cut_matches_oracle(zero_cost_frac, budget):
    """--fair_counting --or_labeling: the plugin cuts the ranked list on the device by multi-hot class counts; the
    reference walk (oracle expand_training_set with the same costs) consumes exactly that prefix."""
    import types
    n, c, h, w, nseg, bs = 6, 8, 48, 64, 24, 2
    logits = synth.logits(n, c, h, w, "cosine", seed=25)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=26)
    im_idx, suppix = synth.pool_lists(n, nseg, spx, labelled_frac=0.1, seed=28)
    args = selector_args("my_bvsb_predclsbal_pwr_banignore", nseg, c - 1, True, 0.1, 6.0, bs)
    selector = importlib.import_module("mulactseg_b200.active_selection.my_bvsb_predclsbal_pwr_banignore").RegionSelector(args)
    g = torch.Generator().manual_seed(29)
    multi_hot = (torch.rand((n + 2, nseg, c), generator=g) < 0.25).to(torch.uint8)
    multi_hot[(torch.rand((n + 2, nseg), generator=g) < zero_cost_frac)] = 0          # zero-cost regions: the prefix must grow
    order = torch.randperm(n + 2, generator=g).tolist()                                  # label rows are not in pool order
    stem = lambda key: key[2].split("/")[-1].split(".")[0]
    label = types.SimpleNamespace(im_idx=[], suppix={}, multi_hot_cls=multi_hot.numpy(),
                                  id_to_index={stem(k): order[i] for i, k in enumerate(im_idx)})
    pool_ds = PoolSet(logits, spx.long(), im_idx, suppix)
    scores = selector.calculate_scores(fake_trainer(DEV), PoolSet(logits, spx.long(), im_idx, suppix))
    cost = lambda spx_path, sid: int(label.multi_hot_cls[label.id_to_index[spx_path.split("/")[-1].split(".")[0]], sid].sum())
    seen = {}

    class ActiveSet:
        args = types.SimpleNamespace(fair_counting=True, or_labeling=True)
        trg_pool_dataset = pool_ds
        trg_label_dataset = label

        def expand_training_set(self, ranked, count, name):
            seen["ranked"] = list(ranked)
            seen["n"] = oa.expand_training_set(ranked, count, label.im_idx, label.suppix, pool_ds.im_idx, pool_ds.suppix, cost)

    selector.select_next_batch(fake_trainer(DEV), ActiveSet(), budget)
    full = sorted(scores, reverse=True)
    want_n = oa.expand_training_set(full, budget, [], {}, [list(k) for k in im_idx], {k: list(v) for k, v in suppix.items()}, cost)
    assert seen["n"] == want_n == len(seen["ranked"])                 # handed exactly the prefix the walk consumes
    assert [(p, i) for _, p, i in seen["ranked"]] == [(p, i) for _, p, i in full[:want_n]]


def test_host_entry_matches_device_path():
    from mulactseg_b200 import _lib
    n, c, h, w, nseg = 5, 20, 64, 128, 96
    logits = synth.logits(n, c, h, w, "cosine", seed=15).contiguous()
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=16, dtype=torch.int32).contiguous()
    score = np.empty(n * nseg, dtype=np.float32)
    dom = np.empty(n * nseg, dtype=np.int32)
    prob = np.empty(n * c, dtype=np.float64)
    _lib.call("mas_acquisition_host", logits.data_ptr(), 0, spx.data_ptr(), n, c, h, w, nseg, 0.1, 1, 6.0, 2, 0, c - 1, 0, 2,
              score.ctypes.data, dom.ctypes.data, prob.ctypes.data)
    ref = oa.scores_predclsbal_pwr(batches(logits, spx.long(), 2), nseg, 0.1, 6.0, ban_ignore=True)
    assert_scores_close(score.reshape(n, nseg), ref.numpy(), False, "host entry")
    # selection through the host entry == numpy on the same scores
    rank = np.arange(n, dtype=np.int32)[::-1].copy()
    mask = (np.random.RandomState(0).rand(n, nseg) < 0.8).astype(np.uint8)
    k = 50
    keys = np.zeros(k, dtype=np.uint64)
    cnt = np.zeros(1, dtype=np.int32)
    _lib.call("mas_select_topk_host", score.ctypes.data, mask.ctypes.data, rank.ctypes.data, n, nseg, k,
              keys.ctypes.data, cnt.ctypes.data)
    from mulactseg_b200 import selection
    im_idx, _ = synth.pool_lists(n, nseg)
    rank_sorted = selection.image_ranks(im_idx)
    cand = [(float(score[i * nseg + s]), int(rank[i]), s) for i in range(n) for s in range(nseg) if mask[i, s]]
    cand.sort(reverse=True)
    assert int(cnt[0]) == k
    tie = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert list(zip((tie // nseg).tolist(), (tie % nseg).tolist())) == [(r, s) for _, r, s in cand[:k]]
    assert rank_sorted.tolist() == list(range(n))


@pytest.mark.parametrize("shape", [(3, 20, 16, 32, 64, 128, 96, torch.float32), (2, 22, 33, 33, 129, 129, 40, torch.float32),
                                   (2, 19, 10, 13, 37, 50, 12, torch.float32), (3, 21, 24, 32, 94, 125, 30, torch.bfloat16),
                                   (1, 6, 8, 8, 8, 8, 4, torch.float32), (2, 20, 64, 128, 256, 512, 512, torch.float32)])
def test_lowres_entry_matches_oracle_on_the_interpolated_logits(shape):
    """the scorer fed with the head's low-resolution logits must equal the reference path applied to
    F.interpolate(logits, size, mode='bilinear', align_corners=False) (models/segmentation/utils.py:28-34) -- integer x4
    (Cityscapes: 256x512 -> 1024x2048), the VOC ratio 129 -> 513, odd sizes, bf16 source, identity size."""
    from mulactseg_b200 import acquisition as acq
    n, c, h_in, w_in, h, w, nseg, dtype = shape
    low = synth.logits(n, c, h_in, w_in, "cosine", seed=h + w).to(dtype)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=9, drop_ids=1 if nseg > 8 else 0)
    full = torch.nn.functional.interpolate(low.float(), size=(h, w), mode="bilinear", align_corners=False)
    pool = batches(full, spx, 2)
    ref = oa.scores_predclsbal_pwr(pool, nseg, 0.1, 6.0, ban_ignore=True)
    spec = acq.SELECTORS["my_bvsb_predclsbal_pwr_banignore"]
    stats = acq.RegionStats(n, nseg, c, DEV, need_prob=True)
    x, ids = low.to(DEV), spx.to(DEV, torch.int32)
    for i in range(0, n, 2):
        stats.add_batch_lowres(i, x[i:i + 2], ids[i:i + 2], 0.1)
    score, _ = acq.finalize(stats, spec, 6.0, 2)
    # Where the two largest interpolated logits of a pixel (nearly) coincide the arg-max is not defined across
    # implementations: border pixels copy a single tap (bf16 sources tie exactly there; torch's topk order among equal
    # probabilities is arbitrary), and an ulp of difference between this kernel's interpolation and torch's CPU kernel can
    # flip a near-tie.  Regions holding such a pixel are left out; everything else is held to the north_star tolerance.
    top2 = full.topk(2, dim=1).values
    unsafe_px = (top2[:, 0] - top2[:, 1]) <= 1e-5
    unsafe = torch.zeros((n, nseg + 1), dtype=torch.bool)
    for i in range(n):
        unsafe[i, spx[i][unsafe_px[i]].clamp(max=nseg)] = True
    safe = ~unsafe[:, :nseg].numpy()
    assert safe.mean() > 0.7
    cnt = stats.cls_cnt.cpu().long()
    assert cnt.sum(dim=(1, 2)).tolist() == [int((spx[i] < nseg).sum()) for i in range(n)]
    np.testing.assert_array_equal(cnt.numpy()[safe], oa.region_histograms(pool, nseg, 0.1).numpy()[safe])
    assert_scores_close(score.cpu().numpy()[safe], ref.numpy()[safe], False, str(shape))
    # and equal to our own full-resolution kernels on the interpolated tensor (same tolerance, independent path)
    stats_full = acq.RegionStats(n, nseg, c, DEV, need_prob=True)
    stats_full.add_batch(0, full.to(DEV), ids, 0.1)
    score_full, _ = acq.finalize(stats_full, spec, 6.0, 2)
    np.testing.assert_allclose(score.cpu().numpy()[safe], score_full.cpu().numpy()[safe], rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(stats.prob_sum.cpu().numpy(), stats_full.prob_sum.cpu().numpy(), rtol=1e-5)


def test_selector_plugin_scores_from_the_low_resolution_head_when_asked():
    """RegionSelector with args.b200_lowres and a net exposing forward_lowres: same scores (1e-5) as the default path on
    the up-sampled logits, without the up-sampled tensor."""
    import types
    from mulactseg_b200.active_selection import my_bvsb_predclsbal_pwr as plugin
    n, c, h, w, nseg = 4, 19, 64, 128, 32
    low = synth.logits(n, c, h // 4, w // 4, "cosine", seed=3)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=4)
    im_idx, suppix = synth.pool_lists(n, nseg, spx)

    class Net:
        def eval(self):
            return self

        def forward_lowres(self, x):
            return x

        def __call__(self, x):
            return torch.nn.functional.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)

    pool = PoolSet(low, spx, im_idx, suppix)        # the 'images' handed to the fake net are the low-resolution logits
    trainer = types.SimpleNamespace(net=Net(), device=torch.device(DEV), model_save_dir="/tmp", selection_iter=1)
    args = selector_args("my_bvsb_predclsbal_pwr", nseg, c, False, 0.1, 6.0, 2)
    full = plugin.RegionSelector(args).score_regions(trainer, pool).scores.cpu().numpy()
    args.b200_lowres = True
    fused = plugin.RegionSelector(args).score_regions(trainer, pool).scores.cpu().numpy()
    np.testing.assert_allclose(fused, full, rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("c,dtype", [(22, torch.float32), (24, torch.float32), (31, torch.float32), (22, torch.bfloat16)])
def test_split_id_ring_equals_ids_in_the_stages(c, dtype, monkeypatch):
    """TMA path, wide class counts: the id rows ride in one buffer per warp with its own barrier when that buys a warp
    (C' >= 22 fp32).  Same histograms / sums as with the ids inside the logits stages, and as the oracle's histograms."""
    from mulactseg_b200 import acquisition as acq
    n, h, w, nseg = 5, 37, 384, 40                                      # several strips, rows not a multiple of anything
    logits = synth.logits(n, c, h, w, "cosine", seed=c).to(dtype)
    spx = synth.superpixel_map(n, h, w, nseg, "jitter", seed=c + 1, dtype=torch.int32)
    out = {}
    for split in ("0", "1"):
        monkeypatch.setenv("MAS_SCORER_SPLIT_IDS", split)
        stats = acq.RegionStats(n, nseg, c, DEV, need_prob=True, group_bytes=0)
        stats.add_batch(0, logits.to(DEV), spx.to(DEV), 0.1)
        out[split] = (stats.cls_cnt.clone(), stats.cls_sum.clone(), stats.prob_sum.clone())
    assert torch.equal(out["0"][0], out["1"][0])
    np.testing.assert_allclose(out["1"][1].cpu().numpy(), out["0"][1].cpu().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(out["1"][2].cpu().numpy(), out["0"][2].cpu().numpy(), rtol=1e-6)
    ref = oa.region_histograms(batches(logits.float(), spx, n), nseg, 0.1).numpy()
    if dtype == torch.float32:
        np.testing.assert_array_equal(out["1"][0].cpu().numpy().astype(np.int64), ref)
