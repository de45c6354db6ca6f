"""Negative SI-SDR with 2-speaker permutation-invariant training, on the GPU (csrc/loss.cu).

Mirrors ``models.io.loss.Loss(loss_func=neg_si_sdr, pit=True)`` of the reference (models/io/loss.py:21-29,95-118,
configs/SpatialNet.yaml:33-37): ``forward(yr_hat, yr, reorder=None, reduce_batch=True) -> (loss, perms, yr_hat)``.
The arithmetic lives in torchmetrics there (absent from the reference tree; algorithm restated in the oracle).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib, ops


class _SiSdrPitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, est: Tensor, ref: Tensor, zero_mean: bool):
        if not est.is_cuda:
            raise _lib.NbssError("nbss_b200.loss runs on CUDA tensors only (there is no CPU path)")
        est_c, ref_c = ops._f32c(est), ops._f32c(ref)
        B, S, Ts = est_c.shape
        sums = torch.empty(B * 12, dtype=torch.float64, device=est.device)
        loss = torch.empty(1, dtype=torch.float32, device=est.device)
        loss_b = torch.empty(B, dtype=torch.float32, device=est.device)
        perm = torch.empty(B, 2, dtype=torch.int32, device=est.device)
        coef = torch.empty(B, 2, 5, dtype=torch.float32, device=est.device)
        _lib.check(ops._K("nbss_sisdr_pit_fwd")(_lib.ptr(est_c), _lib.ptr(ref_c), B, S, ctypes.c_longlong(Ts), int(zero_mean),
                                                 _lib.ptr(sums), _lib.ptr(loss), _lib.ptr(loss_b), _lib.ptr(perm), _lib.ptr(coef),
                                                 _lib.stream_ptr()), "nbss_sisdr_pit_fwd")
        ctx.save_for_backward(est_c, ref_c, coef)
        ctx.mark_non_differentiable(loss_b, perm)
        return loss.reshape(()), loss_b, perm

    @staticmethod
    def backward(ctx, gout: Tensor, _g1, _g2):
        est, ref, coef = ctx.saved_tensors
        B, S, Ts = est.shape
        dest = torch.empty_like(est)
        g = gout.reshape(1).float().contiguous()
        _lib.check(ops._K("nbss_sisdr_pit_bwd")(_lib.ptr(est), _lib.ptr(ref), _lib.ptr(coef), _lib.ptr(g), _lib.ptr(dest), B, S,
                                                 ctypes.c_longlong(Ts), _lib.stream_ptr()), "nbss_sisdr_pit_bwd")
        return dest, None, None


def neg_si_sdr_pit(est: Tensor, ref: Tensor, zero_mean: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """est, ref: [B,2,Ts] -> (mean loss (scalar, differentiable wrt est), per-utterance loss [B], perms [B,2])."""
    return _SiSdrPitFn.apply(est, ref, zero_mean)


class NegSiSdrPitLoss(nn.Module):
    """Drop-in for ``Loss(loss_func=neg_si_sdr, pit=True)`` on time-domain signals."""

    is_scale_invariant_loss = True
    name = "neg_si_sdr"
    mask = None

    def to_CC(self, out: Tensor, Xr: Tensor, stft, XrMM: Tensor):
        """Loss.to_CC for a non-mask loss (models/io/loss.py:120-126): the network output is the estimate itself."""
        return out, {"out": out, "Xr": Xr, "stft": stft, "XrMM": XrMM}

    def forward(self, yr_hat: Tensor, yr: Tensor, reorder: Optional[bool] = None, reduce_batch: bool = True, **kwargs):
        loss, loss_b, perms = neg_si_sdr_pit(yr_hat, yr)
        if reorder:
            yr_hat = torch.gather(yr_hat, 1, perms.long()[:, :, None].expand_as(yr_hat))
        return (loss if reduce_batch else loss_b), perms.long(), yr_hat
