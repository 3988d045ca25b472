"""Drop-in stage-1 losses (same class names, constructors and ``forward`` signatures as the reference's
``nn.Module``s), computed by the fused sm_100a kernels of ``csrc/losses.cu``.

Reference classes mirrored (file:line relative to the reference checkout):
  ``GroupMultiLabelCE``            utils/loss.py:81-141            (targets[..., :-1]; net without ignore channel)
  ``GroupMultiLabelCE_``           trainer/active_joint_multi_predignore.py:74-128
  ``GroupMultiLabelCE_onlymulti``  trainer/active_joint_multi_predignore_mclossablation2.py:17-79
  ``MultiChoiceCE``                utils/loss.py:535-588           (targets[..., :-1], empty rows dropped)
  ``MultiChoiceCE_``               trainer/active_joint_multi_predignore.py:17-73
  ``OnehotCEMultihotChoice``       trainer/active_joint_multi_predignore_lossdecomp.py:16-72  (``strict_multihot=False``)
  ``OnehotCEMultihotChoiceVOC``    trainer/active_joint_multi_lossdecomp.py:17-74             (multi-hot := row-sum > 1)
``forward(inputs (N,C',H,W) f32, targets (N,S,Ct) u8, superpixels (N,H,W) i64|i32, spmasks (N,H,W) bool)``.

All of them are one autograd function, ``segmented_loss_sums``: one pass over the logits of the masked
pixels yields the four bucket sums / counts (one-hot, multi-hot, empty-row, group) and the packed
max-pool table; ``backward`` is one pass writing the dense gradient.  Modules that are called on the same
``inputs`` one after the other (the trainer computes the group loss and the decomposed loss back to back,
``..._lossdecomp.py:101-104``) share a single pass when their temperatures agree -- see ``SharedPass``.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib, ops


# Forward softmax in the reference's operation order (x / T, max-subtract, expf, divide by the sum) instead of the fast
# ex2 / reciprocal form (<= 4 ulp apart, both far inside the 1e-5 tolerance of the losses).  The exact form costs several
# hundred extra instructions per SELECTED pixel and makes the forward kernel issue-bound when most pixels are labelled;
# the stage-2 labeller, where the arg-max PIXEL must agree with torch on near-ties, always uses it.
EXACT_SOFTMAX = False


class _SegmentedLossSums(torch.autograd.Function):
    """sums (4,) f32 = [one-hot, multi-hot, empty-row, group] bucket sums (differentiable w.r.t. inputs);
    counts (4,) f64 (not differentiable)."""

    @staticmethod
    def forward(ctx, inputs, spx, mask, info, tiles, nseg, temperature, flags):
        x = inputs.contiguous()
        acc, gmax = ops.multihot_loss_forward(x, spx, mask, info, nseg, temperature, flags, tiles)
        ctx.save_for_backward(x, spx, mask, info, tiles, gmax if gmax is not None else torch.empty(0, device=x.device))
        ctx.nseg, ctx.temperature, ctx.flags = nseg, temperature, flags
        sums = acc[0::2].to(torch.float32)
        counts = acc[1::2].clone()
        ctx.mark_non_differentiable(counts)
        return sums, counts

    @staticmethod
    def backward(ctx, grad_sums, _grad_counts):
        x, spx, mask, info, tiles, gmax = ctx.saved_tensors
        coef = grad_sums.to(torch.float32).contiguous()
        grad = ops.multihot_loss_backward(x, spx, mask, info, gmax if gmax.numel() else None, coef, ctx.nseg,
                                          ctx.temperature, ctx.flags & ~_lib.MAS_LOSS_EXACT_SOFTMAX, tiles)
        return grad, None, None, None, None, None, None, None


class _SegmentedLosses(torch.autograd.Function):
    """The same fused pass as ``_SegmentedLossSums`` with the reference's normalisations folded in on the device:
    returns the six 0-dim losses of ``mas_multihot_loss_finish_dev`` (views of one (6,) tensor) and the bucket
    counts (4,) f64.  The whole forward side is ONE library call on one workspace (``mas_stage1_loss_fwd_dev``: candidate
    words, active-tile list, fused pass, group reduce, finish) and so is the backward side (coefficients, zero sweep,
    fused pass): at a few percent of labelled pixels the host work of a step, not its kernels, sets the pace."""

    @staticmethod
    def forward(ctx, inputs, trg, spx, mask, nseg, temperature, group_mode, flags):
        x = inputs.contiguous()
        ws, acc, losses = ops.stage1_forward(x, spx, mask, trg, temperature, group_mode, flags)
        ctx.save_for_backward(x, spx, mask, ws)
        ctx.nseg, ctx.temperature, ctx.flags = nseg, temperature, flags
        ctx.set_materialize_grads(False)
        counts = acc[1::2]
        ctx.mark_non_differentiable(counts)
        return (*losses.unbind(0), counts)

    @staticmethod
    def backward(ctx, *grads):
        x, spx, mask, ws = ctx.saved_tensors
        grads = grads[:6]
        if all(g is None for g in grads):
            return (None,) * 8
        grad = ops.stage1_backward(x, spx, mask, ws, ctx.nseg, ctx.temperature, ctx.flags, grads)
        return grad, None, None, None, None, None, None, None


# positions in the tuple returned by ``segmented_losses``
ONE_HOT, MULTI_STRICT, MULTI_NOT_ONE, CHOICE_ALL, GROUP, COUNTS = 0, 1, 2, 3, 4, 6


def _prepare(inputs, targets, superpixels, spmasks):
    if not inputs.is_cuda:
        raise RuntimeError("mulactseg_b200 losses need CUDA tensors (there is no CPU path)")
    n, c = inputs.shape[:2]
    if c > _lib.MAS_MAX_LOSS_CLASSES:
        raise RuntimeError(f"at most {_lib.MAS_MAX_LOSS_CLASSES} channels are supported, got {c}")
    trg = targets if targets.dtype == torch.uint8 else targets.to(torch.uint8)
    spx = superpixels if superpixels.dtype in (torch.int32, torch.int64) else superpixels.long()
    mask = spmasks if spmasks.dtype in (torch.bool, torch.uint8) else spmasks.bool()
    return trg.contiguous(), spx.contiguous(), mask.contiguous()


def segmented_loss_sums(inputs, targets, superpixels, spmasks, temperature: float, group_mode: Optional[int],
                        want_choice: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """One fused pass.  ``group_mode``: None (no group loss), MAS_GROUP_ALL or MAS_GROUP_ONLYMULTI.
    The candidate sets are the first ``inputs.shape[1]`` target channels (== ``targets[..., :C']``)."""
    trg, spx, mask = _prepare(inputs, targets, superpixels, spmasks)
    if inputs.dtype != torch.float32:
        inputs = inputs.float()
    nseg = trg.shape[1]
    info = ops.multihot_info(trg, inputs.shape[1], _lib.MAS_GROUP_ALL if group_mode is None else group_mode)
    flags = (_lib.MAS_LOSS_CHOICE if want_choice else 0) | (_lib.MAS_LOSS_GROUP if group_mode is not None else 0)
    if EXACT_SOFTMAX:
        flags |= _lib.MAS_LOSS_EXACT_SOFTMAX
    return _SegmentedLossSums.apply(inputs, spx, mask, info, ops.multihot_tiles(mask), nseg, float(temperature), flags)


def segmented_losses(inputs, targets, superpixels, spmasks, temperature: float, group_mode: Optional[int], want_choice: bool):
    """Like ``segmented_loss_sums`` but normalised on the device: -> tuple indexed by ONE_HOT .. GROUP (0-dim losses)
    and COUNTS ((4,) f64 bucket counts)."""
    trg, spx, mask = _prepare(inputs, targets, superpixels, spmasks)
    if inputs.dtype != torch.float32:
        inputs = inputs.float()        # bf16 / fp16 logits (autocast heads): widened once, autograd casts the gradient back
    if trg.shape[2] < inputs.shape[1]:
        raise RuntimeError(f"targets carry {trg.shape[2]} channels, fewer than the {inputs.shape[1]} logit channels")
    flags = (_lib.MAS_LOSS_CHOICE if want_choice else 0) | (_lib.MAS_LOSS_GROUP if group_mode is not None else 0)
    if EXACT_SOFTMAX:
        flags |= _lib.MAS_LOSS_EXACT_SOFTMAX
    return _SegmentedLosses.apply(inputs, trg, spx, mask, trg.shape[1], float(temperature),
                                  _lib.MAS_GROUP_ALL if group_mode is None else group_mode, flags)


class SharedPass:
    """Lets several loss modules called on the SAME tensors in a row reuse one fused pass.

    The first module asks for everything any member needs (group mode + choice buckets); members called
    next with identical arguments (same tensor objects and versions, same temperature) get the cached
    sums.  Anything else falls back to its own pass, so results never depend on the sharing."""

    def __init__(self):
        self._key = None
        self._value = None
        self._refs = ()
        self.members = []

    def register(self, module):
        self.members.append(weakref.ref(module))

    def _wanted(self, temperature):
        group_mode, choice = None, False
        for ref in self.members:
            m = ref()
            if m is None or float(m.temp) != float(temperature):
                continue
            if m.group_mode is not None:
                if group_mode is not None and group_mode != m.group_mode:
                    return None
                group_mode = m.group_mode
            choice = choice or m.wants_choice
        return group_mode, choice

    def sums(self, module, inputs, targets, superpixels, spmasks):
        tensors = (inputs, targets, superpixels, spmasks)
        key = (tuple(t._version for t in tensors), float(module.temp), torch.is_grad_enabled(), inputs.requires_grad)
        if self._key == key and self._value is not None and all(r() is t for r, t in zip(self._refs, tensors)):
            gm, choice, value = self._value
            if (module.group_mode is None or module.group_mode == gm) and (choice or not module.wants_choice):
                return value
        wanted = self._wanted(module.temp)
        if wanted is None:
            wanted = (module.group_mode, module.wants_choice)
        value = segmented_losses(inputs, targets, superpixels, spmasks, module.temp, wanted[0], wanted[1])
        # identity through weak references: a recycled id() of a dead tensor can never hit the cache
        self._refs = tuple(weakref.ref(t) for t in tensors)
        self._key, self._value = key, (wanted[0], wanted[1], value)
        return value


class _SegmentedLoss(nn.Module):
    group_mode: Optional[int] = None
    wants_choice = False

    def __init__(self, temperature=1.0, reduction="mean"):
        super().__init__()
        self.eps = 1e-8
        self.temp = temperature
        self.reduction = reduction
        self.shared: Optional[SharedPass] = None

    def share(self, shared: SharedPass):
        self.shared = shared
        shared.register(self)
        return self

    def _losses(self, inputs, targets, superpixels, spmasks):
        if self.shared is not None:
            return self.shared.sums(self, inputs, targets, superpixels, spmasks)
        return segmented_losses(inputs, targets, superpixels, spmasks, self.temp, self.group_mode, self.wants_choice)


class GroupMultiLabelCE(_SegmentedLoss):
    """utils/loss.py:81-141.  The net has ``targets.shape[-1] - 1`` channels; candidates = targets[..., :-1]."""
    group_mode = _lib.MAS_GROUP_ALL

    def __init__(self, args, num_class, num_superpixel, temperature=1.0, reduction="mean"):
        super().__init__(temperature, reduction)
        self.args = args
        self.num_class = num_class
        self.num_superpixel = num_superpixel

    def forward(self, inputs, targets, superpixels, spmasks):
        if self.reduction == "mean":
            return self._losses(inputs, targets, superpixels, spmasks)[GROUP]
        if self.reduction == "none":                     # (loss sum, num_valid), utils/loss.py:137-139
            sums, counts = segmented_loss_sums(inputs, targets, superpixels, spmasks, self.temp, self.group_mode, False)
            return sums[3], 1 + counts[3]
        raise NotImplementedError


class GroupMultiLabelCE_(GroupMultiLabelCE):
    """trainer/active_joint_multi_predignore.py:74-128 (targets not sliced: the net predicts the ignore channel)."""


class GroupMultiLabelCE_onlymulti(GroupMultiLabelCE_):
    """trainer/active_joint_multi_predignore_mclossablation2.py:17-79: only pixels of multi-hot superpixels."""
    group_mode = _lib.MAS_GROUP_ONLYMULTI


class MultiChoiceCE(_SegmentedLoss):
    """utils/loss.py:535-588: -log of the candidate-set probability mass over labelled masked pixels."""
    wants_choice = True

    def __init__(self, num_class, temperature=1.0, reduction="mean"):
        super().__init__(temperature, reduction)
        self.num_class = num_class

    def forward(self, inputs, targets, superpixels, spmasks):
        if self.reduction == "mean":
            return self._losses(inputs, targets, superpixels, spmasks)[CHOICE_ALL]
        if self.reduction == "none":                     # (loss sum, num_valid), utils/loss.py:583-586; empty rows dropped (:574-577)
            sums, counts = segmented_loss_sums(inputs, targets, superpixels, spmasks, self.temp, None, True)
            return sums[0] + sums[1], 1 + counts[0] + counts[1]
        raise NotImplementedError


class MultiChoiceCE_(MultiChoiceCE):
    """trainer/active_joint_multi_predignore.py:17-73."""


class OnehotCEMultihotChoice(MultiChoiceCE):
    """trainer/active_joint_multi_predignore_lossdecomp.py:16-72 -> (one-hot CE, multi-hot multi-choice).

    The reference defines multi-hot := not one-hot and asserts that this equals row-sum > 1 (:65-67), i.e. it
    raises on a selected pixel whose superpixel has no candidate class.  ``assert_partition=True`` (default)
    keeps that check (one device sync per call, as in the reference); ``'deferred'`` copies the counter to pinned host
    memory asynchronously and raises at the NEXT call (or ``check_partition()``), so the host keeps running ahead of
    the GPU; False skips it and counts such pixels in the multi-hot bucket like the reference's arithmetic would."""
    strict_multihot = False

    def __init__(self, num_class, temperature=1.0, reduction="mean", assert_partition=True):
        super().__init__(num_class, temperature, reduction)
        assert self.reduction == "mean"
        self.assert_partition = assert_partition
        self._pending = None        # (pinned host copy of the empty-row count, event) of the previous call
        self._pinned = None         # the pinned buffer, allocated once

    def check_partition(self):
        """Raise if the previous call saw a selected pixel whose superpixel has no candidate class
        (``assert_partition='deferred'``: same assertion as the reference, without stalling the stream)."""
        if self._pending is not None:
            host, event = self._pending
            self._pending = None
            event.synchronize()
            assert float(host[0]) == 0.0   # ..._lossdecomp.py:67

    def forward(self, inputs, targets, superpixels, spmasks):
        if self.assert_partition == "deferred":
            self.check_partition()
        out = self._losses(inputs, targets, superpixels, spmasks)
        if self.strict_multihot:
            return out[ONE_HOT], out[MULTI_STRICT]
        if self.assert_partition == "deferred":
            if self._pinned is None:
                self._pinned = torch.empty(1, dtype=torch.float64).pin_memory()
            host = self._pinned
            host.copy_(out[COUNTS][2:3], non_blocking=True)
            event = torch.cuda.Event()
            event.record()
            self._pending = (host, event)
        elif self.assert_partition:
            assert float(out[COUNTS][2]) == 0.0   # ..._lossdecomp.py:67 (one device sync, as in the reference)
        return out[ONE_HOT], out[MULTI_NOT_ONE]


class OnehotCEMultihotChoiceVOC(OnehotCEMultihotChoice):
    """trainer/active_joint_multi_lossdecomp.py:17-74 (multi-hot := row-sum > 1, no assertion)."""
    strict_multihot = True


def stage1_criterion(args, num_classes: int, voc: bool = False, assert_partition="deferred"):
    """The two criteria ``ActiveTrainer.get_criterion`` installs (``..._lossdecomp.py:78-81``), wired to share
    one fused pass per step when ``group_ce_temp == multi_ce_temp``.  The reference's partition assertion is kept
    but checked one step late (``assert_partition='deferred'``) so that a training step never waits for the GPU."""
    shared = SharedPass()
    group = GroupMultiLabelCE_onlymulti(args=args, num_class=num_classes, num_superpixel=args.nseg,
                                        temperature=args.group_ce_temp).share(shared)
    cls = OnehotCEMultihotChoiceVOC if voc else OnehotCEMultihotChoice
    multi = cls(num_class=num_classes, temperature=args.multi_ce_temp, assert_partition=assert_partition).share(shared)
    return group, multi
