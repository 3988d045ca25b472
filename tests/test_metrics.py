"""mIoU counters and dominant label assignment: oracle vs the vectors produced by the unmodified
reference classes (CPU tier) and the CUDA kernels vs both (GPU tier).  Integer outputs: bit-exact."""
import os

import numpy as np
import pytest
import torch

from mulactseg_b200 import synth
from oracle import metrics as om

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MET = np.load(os.path.join(GOLDEN, "metrics.npz"))
MIOU_CASES = sorted({k.split("/")[0] for k in MET.files if "/_after" in k})
DOM_CASES = sorted({k.split("/")[0] for k in MET.files if k.startswith("dominant")})
HOOKS = ["_after_step", "_after_step_within_predregion"]


@pytest.mark.parametrize("case", MIOU_CASES)
@pytest.mark.parametrize("hook", HOOKS)
def test_miou_oracle_matches_reference_golden(case, hook):
    c, ignore = (int(v) for v in MET[f"{case}/meta"])
    counts = om.miou_counts(MET[f"{case}/outputs"], MET[f"{case}/targets"], c, ignore, hook.endswith("predregion"))
    np.testing.assert_array_equal(counts, MET[f"{case}/{hook}/counts"])
    np.testing.assert_array_equal(np.asarray(om.ious(counts), dtype=np.float64), MET[f"{case}/{hook}/ious"])
    np.testing.assert_array_equal(np.asarray(om.ious(counts, [0, c - 1]), dtype=np.float64), MET[f"{case}/{hook}/ious_skip"])
    np.testing.assert_array_equal(np.asarray(om.ious_precisions_recalls(counts), dtype=np.float64), MET[f"{case}/{hook}/ipr"])


@pytest.mark.parametrize("case", DOM_CASES)
def test_dominant_oracle_matches_reference_golden(case):
    got = om.dominant_target(MET[f"{case}/target"], MET[f"{case}/spx"], MET[f"{case}/ids"].tolist())
    np.testing.assert_array_equal(got, MET[f"{case}/labels"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", MIOU_CASES)
@pytest.mark.parametrize("hook", HOOKS)
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.uint8, "numpy"])
def test_miou_kernel_matches_reference_golden(case, hook, dtype):
    from mulactseg_b200.miou import MeanIoU
    c, ignore = (int(v) for v in MET[f"{case}/meta"])
    outputs, targets = MET[f"{case}/outputs"], MET[f"{case}/targets"]
    helper = MeanIoU(c, ignore)
    helper._before_epoch()
    for i in range(outputs.shape[0]):                      # one call per image, like the reference run that made the vectors
        if dtype == "numpy":
            batch = {"outputs": outputs[i:i + 1], "targets": targets[i:i + 1]}
        else:
            batch = {"outputs": torch.from_numpy(outputs[i:i + 1]).to("cuda", dtype), "targets": torch.from_numpy(targets[i:i + 1]).to("cuda", dtype)}
        getattr(helper, hook)(batch)
    got = np.stack([helper.total_seen, helper.total_correct, helper.total_positive])
    assert got.dtype == np.float64                         # the reference keeps float counters
    np.testing.assert_array_equal(got.astype(np.int64), MET[f"{case}/{hook}/counts"])
    np.testing.assert_array_equal(np.asarray(helper._after_epoch(), dtype=np.float64), MET[f"{case}/{hook}/ious"])
    np.testing.assert_array_equal(np.asarray(helper._after_epoch([0, c - 1]), dtype=np.float64), MET[f"{case}/{hook}/ious_skip"])
    np.testing.assert_array_equal(np.asarray(helper._after_epoch_ipr(), dtype=np.float64), MET[f"{case}/{hook}/ipr"])
    helper._before_epoch()
    assert helper.total_seen.sum() == 0


@pytest.mark.gpu
def test_miou_full_size_and_mixed_dtypes():
    """Cityscapes validation shape: counts equal a bincount-based evaluation; argmax output (int64) vs uint8 labels."""
    from mulactseg_b200 import ops
    from mulactseg_b200.miou import MeanIoU
    n, c, h, w = 2, 19, 1024, 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    targets = torch.randint(0, c, (n, h, w), device="cuda", generator=g)
    targets[torch.rand((n, h, w), device="cuda", generator=g) < 0.1] = 255
    outputs = torch.where(torch.rand((n, h, w), device="cuda", generator=g) < 0.7, targets.clamp(max=c - 1),
                          torch.randint(0, c, (n, h, w), device="cuda", generator=g))
    helper = MeanIoU(c, 255)
    helper._before_epoch()
    helper._after_step({"outputs": outputs, "targets": targets.to(torch.uint8)})       # promoted to a common dtype
    keep = targets != 255
    conf = torch.bincount(targets[keep] * c + outputs[keep], minlength=c * c).view(c, c).cpu().numpy()
    np.testing.assert_array_equal(helper.total_seen.astype(np.int64), conf.sum(1))
    np.testing.assert_array_equal(helper.total_positive.astype(np.int64), conf.sum(0))
    np.testing.assert_array_equal(helper.total_correct.astype(np.int64), np.diag(conf))
    with pytest.raises(RuntimeError):
        ops.miou_counts(outputs.cpu(), targets.cpu(), c, 255, False, torch.zeros(3 * c, dtype=torch.int64))   # no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("case", DOM_CASES)
@pytest.mark.parametrize("id_dtype", [torch.int64, torch.int32])
def test_dominant_kernel_matches_reference_golden(case, id_dtype):
    from mulactseg_b200 import label_assignment
    nseg, c = (int(v) for v in MET[f"{case}/meta"])
    got = label_assignment.dominant_target(torch.from_numpy(MET[f"{case}/target"]), torch.from_numpy(MET[f"{case}/spx"]).to(id_dtype),
                                           MET[f"{case}/ids"].tolist(), nseg, c)
    np.testing.assert_array_equal(got.numpy(), MET[f"{case}/labels"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(97, 131, 60, 19, "jitter"), (64, 64, 30, 21, "grid"), (50, 70, 2048, 19, "random"),
                                   (256, 512, 512, 19, "jitter")])
def test_dominant_kernel_matches_oracle(shape):
    from mulactseg_b200 import label_assignment
    h, w, nseg, c, kind = shape
    spx = synth.superpixel_map(1, h, w, nseg, kind, seed=h, drop_ids=2)[0]
    g = torch.Generator().manual_seed(w)
    target = torch.randint(0, c, (h, w), generator=g)
    target[torch.rand((h, w), generator=g) < 0.1] = 255
    ids = sorted(set(torch.unique(spx).tolist()[::2] + [nseg - 1]))
    ref = om.dominant_target(target.numpy().astype(np.uint8), spx.numpy(), ids)
    got = label_assignment.dominant_target(target, spx, ids, nseg, c)
    np.testing.assert_array_equal(got.numpy(), ref)
    # idempotent: a dominant-labelled map is its own dominant labelling
    again = label_assignment.dominant_target(got, spx, ids, nseg, c)
    assert torch.equal(again, got)
