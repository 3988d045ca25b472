"""Operator-level drop-in for the ``torch_scatter`` calls of the reference ("what actually gets
replaced"): ``scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'|'add'|'mean'|'max')`` and
``scatter_max(src, index, dim=-1, out=None, dim_size=None) -> (out, arg)`` with the semantics of torch_scatter 2.0.9
(``actsegmul.yml:99``), computed by the segmented-reduction kernels of ``csrc/scatter.cu``:

  * index broadcast the torch_scatter way (1-D index -> leading singleton dims up to ``dim``; missing trailing dims are
    shared, e.g. a (B, HW) index against a (B, HW, C') one-hot -- ``my_bvsb_banignore.py:44-45``);
  * segments nobody writes hold 0; ``scatter_max``'s arg is the FIRST element attaining the maximum and ``src.size(dim)``
    for an empty segment (``utils/loss.py:202-204`` relies on both); 'mean' divides by ``clamp(count, min=1)``;
  * differentiable w.r.t. ``src``: sum / mean gather the incoming gradient, max routes it to the arg element only.

The twelve rebuilt hot-path plugins do NOT call these (their fused kernels never materialise the operands); this module is
for the reference's remaining call sites (loss ablations, other selectors), switched with

    import mulactseg_b200.torch_scatter_compat as torch_scatter          # or sys.modules["torch_scatter"] = ...

CUDA tensors only -- there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .ops import _on, _stream


def _view(src: torch.Tensor, index: torch.Tensor, dim: int):
    """-> (dim, outer, n, inner, index (int64, contiguous), index_has_inner)."""
    if not isinstance(src, torch.Tensor) or not src.is_cuda:
        raise RuntimeError("torch_scatter_compat: expected CUDA tensors (there is no CPU path)")
    if dim < 0:
        dim += src.dim()
    if not 0 <= dim < src.dim():
        raise IndexError(f"dim {dim} out of range for a {src.dim()}-d src")
    if index.dim() == 1 and dim > 0:
        index = index.view((1,) * dim + (-1,))
    if index.dim() > src.dim():
        raise RuntimeError("index has more dimensions than src")
    shape = tuple(src.shape)
    n = shape[dim]
    outer = 1
    for s in shape[:dim]:
        outer *= s
    inner = 1
    for s in shape[dim + 1:]:
        inner *= s
    lead = tuple(index.shape[: dim + 1]) + (1,) * max(0, dim + 1 - index.dim())
    trailing_shared = index.dim() <= dim + 1 or all(s == 1 for s in index.shape[dim + 1:])
    if trailing_shared:
        idx = index.reshape(lead[: dim + 1]).expand(shape[: dim + 1]).contiguous().view(outer, n)
        has_inner = 0
    else:
        full = index
        while full.dim() < src.dim():
            full = full.unsqueeze(-1)
        idx = full.expand(shape).contiguous()
        has_inner = 1
    return dim, outer, n, inner, idx.long(), has_inner


def _dim_size(idx: torch.Tensor, dim_size: Optional[int]) -> int:
    if dim_size is not None:
        return int(dim_size)
    return int(idx.max()) + 1 if idx.numel() else 0


def _out_shape(src, dim, size):
    shape = list(src.shape)
    shape[dim] = size
    return shape


def _raw_sum(src: torch.Tensor, idx: torch.Tensor, outer, n, inner, has_inner, size, out: torch.Tensor) -> None:
    dtype = _lib.MAS_F32 if src.dtype == torch.float32 else _lib.MAS_SCATTER_I64
    with _on(src):
        _lib.call("mas_scatter_sum_dev", src.data_ptr(), dtype, idx.data_ptr(), outer, n, inner, has_inner, size, out.data_ptr(),
                  _stream(src))


class _ScatterSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, dim, dim_size):
        dim, outer, n, inner, idx, has_inner = _view(src, index, dim)
        size = _dim_size(idx, dim_size)
        work = src.contiguous()
        if work.dtype not in (torch.float32, torch.int64):
            work = work.float() if work.is_floating_point() else work.long()
        out = torch.zeros(_out_shape(src, dim, size), dtype=work.dtype, device=src.device)
        _raw_sum(work, idx, outer, n, inner, has_inner, size, out)
        ctx.dim, ctx.shape = dim, tuple(src.shape)
        ctx.save_for_backward(idx)
        ctx.has_inner = has_inner
        return out if out.dtype == src.dtype else out.to(src.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        dim, shape = ctx.dim, ctx.shape
        gather_index = idx.view(shape) if ctx.has_inner else idx.view(shape[: dim + 1] + (1,) * (len(shape) - dim - 1)).expand(shape)
        valid = (gather_index >= 0) & (gather_index < grad_out.shape[dim])
        grad = grad_out.gather(dim, gather_index.clamp(0, max(grad_out.shape[dim] - 1, 0)))
        return grad * valid, None, None, None


class _ScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, dim, dim_size):
        dim, outer, n, inner, idx, has_inner = _view(src, index, dim)
        size = _dim_size(idx, dim_size)
        work = src.contiguous().float()
        shape = _out_shape(src, dim, size)
        out = torch.empty(shape, dtype=torch.float32, device=src.device)
        arg = torch.full(shape, n, dtype=torch.int64, device=src.device)
        key = torch.empty(max(out.numel(), 1), dtype=torch.int32, device=src.device)
        with _on(src):
            _lib.call("mas_scatter_max_dev", work.data_ptr(), idx.data_ptr(), outer, n, inner, has_inner, size, out.data_ptr(),
                      arg.data_ptr(), key.data_ptr(), _stream(src))
        ctx.dim, ctx.shape = dim, tuple(src.shape)
        ctx.save_for_backward(arg)
        ctx.mark_non_differentiable(arg)
        return out.to(src.dtype), arg

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        (arg,) = ctx.saved_tensors
        dim = ctx.dim
        shape = list(ctx.shape)
        shape[dim] += 1                       # slot n swallows the gradient of empty segments
        grad = grad_out.new_zeros(shape)
        grad.scatter_(dim, arg, grad_out)
        return grad.narrow(dim, 0, shape[dim] - 1), None, None, None


def scatter_sum(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None):
    if out is not None:
        raise NotImplementedError("out= is not used by the reference")
    return _ScatterSum.apply(src, index, dim, dim_size)


scatter_add = scatter_sum


def scatter_mean(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None):
    total = scatter_sum(src, index, dim, out, dim_size)
    d, outer, n, inner, idx, has_inner = _view(src, index, dim)
    size = total.shape[d]
    count = torch.zeros(_out_shape(src, d, size) if has_inner else list(src.shape[:d]) + [size], dtype=torch.int64, device=src.device)
    ones = torch.ones(idx.shape, dtype=torch.int64, device=src.device)
    _raw_sum(ones, idx, outer, n, inner if has_inner else 1, has_inner, size, count)
    count = count.clamp_(min=1)
    if not has_inner:
        count = count.view(list(src.shape[:d]) + [size] + [1] * (src.dim() - d - 1))
    if total.is_floating_point():
        return total / count.to(total.dtype)
    return torch.div(total, count, rounding_mode="floor")


def scatter_max(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    if out is not None:
        raise NotImplementedError("out= is not used by the reference")
    return _ScatterMax.apply(src, index, dim, dim_size)


def scatter_min(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None):
    val, arg = scatter_max(-src, index, dim, out, dim_size)
    return -val, arg


def scatter(src, index, dim: int = -1, out=None, dim_size: Optional[int] = None, reduce: str = "sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    if reduce == "min":
        return scatter_min(src, index, dim, out, dim_size)[0]
    raise ValueError(f"reduce={reduce!r} is not used by the reference")
