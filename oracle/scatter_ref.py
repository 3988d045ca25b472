"""Restatement of the ``torch_scatter`` 2.0.9 operators the reference calls.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``torch_scatter`` is a
third-party dependency that is absent from the reference tree and from this
image (pinned by the reference at ``actsegmul.yml:99``:
``pytorch-scatter=2.0.9=py38_torch_1.11.0_cu113``).  Its published semantics,
which the reference's own comments rely on (``active_selection/my_bvsb.py:73``
"value for non-existing spx id == 0"; ``utils/loss.py:202-204`` "value == 0.0,
index == (max_index + 1)"), are restated here with torch-native CPU ops:

* ``scatter(reduce='sum')``  : zeros(dim_size).scatter_add_(dim, index*, src)
* ``scatter(reduce='mean')`` : sum / clamp(count, min=1) (floor-div for ints)
* ``scatter(reduce='max')``  : ``scatter_max(...)[0]``
* ``scatter_max``            : segments never written hold value 0 and
  arg == src.size(dim); on CPU the FIRST element wins ties (strict ``>``
  update in index order); backward routes the gradient to the arg element only.

``index*`` is ``index`` broadcast to ``src`` the torch_scatter way: leading
singleton dims up to ``dim`` if it is 1-D, then trailing singleton dims, then
``expand_as(src)``.

Call sites on the hot path (reference file:line): ``my_bvsb.py:73``,
``my_bvsb_banignore.py:43,45``, ``my_bvsb_predclsbal_pwr[_banignore].py:65,69``,
``my_bvsb_clsbal_v2*.py:44-47``, ``utils/loss.py:122``,
``trainer/active_joint_multi_predignore.py:109``,
``trainer/active_joint_multi_predignore_mclossablation2.py:60``,
``trainer/eval_save_cosplbl_prop*.py:178,213``.
"""
from __future__ import annotations

import torch


def _expand_index(index: torch.Tensor, src: torch.Tensor, dim: int) -> torch.Tensor:
    if dim < 0:
        dim += src.dim()
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def _out_size(src, index, dim, dim_size):
    size = list(src.size())
    if dim_size is not None:
        size[dim] = int(dim_size)
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _expand_index(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter_sum(src, index, dim, out, dim_size)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    total = scatter_sum(src, index, dim, out, dim_size)
    dim_size = total.size(dim)
    index_dim = dim + src.dim() if dim < 0 else dim
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _expand_index(count, total, dim)
    if total.is_floating_point():
        total.true_divide_(count)
    else:
        total.div_(count, rounding_mode="floor")
    return total


class _ScatterMax(torch.autograd.Function):
    """value/arg of the per-segment maximum; grad goes to the arg element only."""

    @staticmethod
    def forward(ctx, src, index, dim, dim_size):
        if dim < 0:
            dim += src.dim()
        idx = _expand_index(index, src, dim).contiguous()
        size = _out_size(src, idx, dim, dim_size)
        n = src.size(dim)
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
        out.scatter_reduce_(dim, idx, src.detach(), "amax", include_self=False)
        # first index attaining the maximum (CPU torch_scatter: strict '>' in order)
        shape = [1] * src.dim()
        shape[dim] = n
        pos = torch.arange(n, device=src.device).view(shape).expand_as(src)
        hit = src.detach() == out.gather(dim, idx)
        cand = torch.where(hit, pos, torch.full_like(pos, n))
        arg = torch.full(size, n, dtype=torch.long, device=src.device)
        arg.scatter_reduce_(dim, idx, cand, "amin", include_self=True)
        ctx.dim = dim
        ctx.src_shape = src.shape
        ctx.save_for_backward(arg)
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        dim = ctx.dim
        shape = list(ctx.src_shape)
        shape[dim] += 1  # slot n swallows the gradient of empty segments
        grad_src = grad_out.new_zeros(shape)
        grad_src.scatter_(dim, arg, grad_out)
        grad_src = grad_src.narrow(dim, 0, shape[dim] - 1)
        return grad_src, None, None, None


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    if out is not None:
        raise NotImplementedError("out= is not used by the reference hot path")
    return _ScatterMax.apply(src, index, dim, dim_size)


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    val, arg = scatter_max(-src, index, dim, out, dim_size)
    return -val, arg


def scatter_mul(src, index, dim=-1, out=None, dim_size=None):
    index = _expand_index(index, src, dim)
    if out is None:
        out = torch.ones(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_reduce_(dim, index, src, "prod", include_self=True)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    if reduce == "min":
        return scatter_min(src, index, dim, out, dim_size)[0]
    if reduce == "mul":
        return scatter_mul(src, index, dim, out, dim_size)
    raise ValueError(reduce)
