"""The cycle-consistency step that consumes the fusion path's MGFM output.

The reference trainer sums ``f4_global_fusion[view]`` over the spatial axes and feeds the [frames, C] result to
``Trainer.seg_cycle`` / ``Trainer.dense_seg_cycle`` (R/main.py:229-235, :650-717, :719-798) — some 60 tiny
repeat / gather / softmax launches per start position and as many again in autograd.  Here:

* ``spatial_sum(x)``        one kernel, reads the channels-last view the fusion path returns coalesced along C, fp32
                            accumulation; its backward is a broadcast that ``GlobalLocalFusion``'s backward consumes
                            without materialising it per view (``views_to_tokens``, token stride 0);
* ``seg_cycle`` / ``dense_seg_cycle``   same names, arguments and random draw as the reference methods; one launch
                            evaluates the loss AND its gradient (one CTA per start position), one more adds the
                            positions in a fixed order.  The loss is a scalar, so backward is ``grad_output * dfeat``.

CUDA only (libglf_sm100a); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from .tpavi import _io_dtype, _stream_ptr


class _SpatialSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if x.dim() != 4:
            raise ValueError("spatial_sum expects [B, C, h, w]")
        B, Cn, h, w = x.shape
        sb, sc, sh, sw = x.stride()
        if sh != w * sw or min(sb, sc, sw) < 1:          # h, w must collapse into one axis
            x = x.contiguous()
            sb, sc, sh, sw = x.stride()
        ctx.shape, ctx.dtype = (B, Cn, h, w), x.dtype
        out = torch.empty((B, Cn), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.load().glf_spatial_sums(L.ptr(x), _io_dtype(x), B, Cn, h * w, sb, sc, sw, L.ptr(out),
                                              _stream_ptr()))
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cn, h, w = ctx.shape
        return g.to(ctx.dtype).contiguous().view(B, Cn, 1, 1).expand(B, Cn, h, w)


def spatial_sum(x: torch.Tensor) -> torch.Tensor:
    """``x.sum(dim=(2, 3))`` of a [B, C, h, w] CUDA tensor (R/main.py:229), returned in fp32 whatever the storage
    dtype (the reference sums fp32 activations)."""
    if not x.is_cuda:
        raise L.GlfError("glfusion_b200.cycle.spatial_sum needs CUDA tensors (there is no CPU path)")
    return _SpatialSum.apply(x)


class _CycleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, target_region, cyc_off, chunk_size, temperature, start, step, n_starts, soft_label, scale):
        if not feat.is_cuda:
            raise L.GlfError("glfusion_b200.cycle needs CUDA tensors (there is no CPU path)")
        if feat.dim() != 2:
            raise ValueError("cycle loss expects [frames, C] features")
        f = feat.detach().float().contiguous()
        T, Cn = f.shape
        lib = L.load()
        loss = torch.empty((), dtype=torch.float32, device=f.device)
        dfeat = torch.empty_like(f)
        nbytes = int(lib.glf_cycle_loss_scratch_bytes(T, Cn, n_starts))
        scratch = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=f.device)
        with torch.cuda.device(f.device):
            L.check(lib.glf_cycle_loss(L.ptr(f), T, Cn, int(target_region), int(cyc_off), int(chunk_size),
                                       float(temperature), int(start), int(step), int(n_starts), int(bool(soft_label)),
                                       float(scale), L.ptr(loss), L.ptr(dfeat), L.ptr(scratch), _stream_ptr()))
        ctx.save_for_backward(dfeat)
        ctx.dtype = feat.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (dfeat,) = ctx.saved_tensors
        return ((g * dfeat).to(ctx.dtype),) + (None,) * 9


def positions(target_region: int, cyc_off: int, chunk_size: int) -> int:
    """Number of start positions = number of logits (R/main.py:655)."""
    return target_region - (chunk_size + cyc_off) + 1


def seg_cycle(feat_out: torch.Tensor, target_region: int, cyc_off: int, chunk_size: int, temperature: float,
              target_strtpt: Optional[int] = None) -> torch.Tensor:
    """``Trainer.seg_cycle`` (R/main.py:650-717).  ``target_strtpt`` None draws the start position exactly as the
    reference does (one ``np.random.choice`` call, main.py:655), so a seeded run consumes the same random stream."""
    n = positions(target_region, cyc_off, chunk_size)
    if n < 1:
        raise ValueError("target_region too small for chunk_size + cyc_off")
    if target_strtpt is None:
        target_strtpt = int(np.random.choice(n))
    return _CycleLoss.apply(feat_out, target_region, cyc_off, chunk_size, temperature, int(target_strtpt), 1, 1, False,
                            1.0)


def dense_seg_cycle(feat_out: torch.Tensor, target_region: int, cyc_off: int, chunk_size: int, temperature: float,
                    soft_label: bool = False, is_overlap: bool = True) -> torch.Tensor:
    """``Trainer.dense_seg_cycle`` (R/main.py:719-798): every start position (stride 1, or ``chunk_size`` without
    overlap), divided by the number of positions (main.py:798)."""
    n = positions(target_region, cyc_off, chunk_size)
    if n < 1:
        raise ValueError("target_region too small for chunk_size + cyc_off")
    step = 1 if is_overlap else chunk_size
    n_starts = len(range(0, n, step))
    return _CycleLoss.apply(feat_out, target_region, cyc_off, chunk_size, temperature, 0, step, n_starts, soft_label,
                            1.0 / n)
