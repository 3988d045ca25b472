"""Operator row of the drop-in boundary: mulactseg_b200.torch_scatter_compat against the restatement of
torch_scatter 2.0.9 the oracle uses (oracle/scatter_ref.py), on the index / src shapes of the reference's call sites."""
import numpy as np
import pytest
import torch

from mulactseg_b200 import torch_scatter_compat as ts
from oracle import scatter_ref as ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cases():
    g = torch.Generator().manual_seed(0)
    b, hw, c, s = 3, 500, 7, 40
    yield "mean (B,HW) like my_bvsb.py:73", torch.rand((b, hw), generator=g), torch.randint(0, s - 3, (b, hw), generator=g), 1, s
    onehot = torch.nn.functional.one_hot(torch.randint(0, c, (b, hw), generator=g), c)
    yield "int64 one-hot sum, shared index like my_bvsb_banignore.py:44-45", onehot, torch.randint(0, s, (b, hw), generator=g), 1, s
    yield "max (HW',C) with (HW',1) index like utils/loss.py:122", torch.rand((hw, c), generator=g), torch.randint(0, s - 5, (hw, 1), generator=g), 0, s
    yield "1-d index along dim 1", torch.randn((4, 30), generator=g), torch.randint(0, 6, (30,), generator=g), 1, 6
    yield "full-shape index, middle dim", torch.randn((2, 25, 3), generator=g), torch.randint(0, 9, (2, 25, 3), generator=g), 1, 9
    yield "dim_size inferred", torch.randn((50,), generator=g), torch.randint(0, 11, (50,), generator=g), 0, None


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_scatter_matches_the_torch_scatter_restatement(case):
    name, src, index, dim, size = case
    reduces = ["sum"] if not src.is_floating_point() else ["sum", "mean", "max"]
    for reduce in reduces:
        want = ref.scatter(src, index, dim=dim, dim_size=size, reduce=reduce)
        got = ts.scatter(src.to(DEV), index.to(DEV), dim=dim, dim_size=size, reduce=reduce)
        assert got.dtype == want.dtype and tuple(got.shape) == tuple(want.shape)
        if src.is_floating_point():
            np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6, err_msg=f"{name}/{reduce}")
        else:
            assert torch.equal(got.cpu(), want)
    if src.is_floating_point():
        want_v, want_a = ref.scatter_max(src, index, dim=dim, dim_size=size)
        got_v, got_a = ts.scatter_max(src.to(DEV), index.to(DEV), dim=dim, dim_size=size)
        assert torch.equal(got_v.cpu(), want_v) and torch.equal(got_a.cpu(), want_a)       # values and FIRST arg, empty -> (0, n)


def test_gradients_and_ties():
    g = torch.Generator().manual_seed(1)
    src = torch.randn((2, 60, 4), generator=g)
    src[0, 5] = src[0, 3]                              # an exact tie inside a segment: the first element takes the gradient
    index = torch.randint(0, 8, (2, 60), generator=g)
    index[0, 3] = index[0, 5] = 2
    for fn in ("sum", "mean", "max"):
        a = src.clone().requires_grad_(True)
        b = src.clone().to(DEV).requires_grad_(True)
        w = torch.randn((2, 10, 4), generator=g)
        (ref.scatter(a, index, dim=1, dim_size=10, reduce=fn) * w).sum().backward()
        (ts.scatter(b, index.to(DEV), dim=1, dim_size=10, reduce=fn) * w.to(DEV)).sum().backward()
        np.testing.assert_allclose(b.grad.cpu().numpy(), a.grad.numpy(), rtol=1e-5, atol=1e-6, err_msg=fn)
    with pytest.raises(RuntimeError):
        ts.scatter(src, index, dim=1)                   # CPU tensors: no CPU path
