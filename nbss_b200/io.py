"""Drop-in ``STFT`` and ``Norm`` for the reference's ``models.io.stft.STFT`` / ``models.io.norm.Norm`` plus the fused
wave -> wave path of ``TrainModule.forward`` (SharedTrainer.py:104-132).  All arithmetic is in libnbss_b200.so.

* ``STFT.stft(x[..., Ts]) -> (X[..., F, T] complex64, Ts)`` and ``STFT.istft(X, Ts)`` (models/io/stft.py:49-97): one
  kernel launch each (the reference loops ``torch.istft`` over B*S items in Python).  ``istft`` is differentiable.
* ``Norm.norm(X, ref_channel=) -> (X, (Xr, XrMM))`` / ``Norm.inorm`` for ``mode='frequency', online=True`` — the only mode
  the SpatialNet configs use (configs/SpatialNet.yaml:40-43); other modes raise ``NotImplementedError``.
* ``SeparationPipeline``: stft + norm + pack fused into one kernel, the network, and unpack + inverse-norm + iSTFT fused
  into one kernel; this is what ``bench.py`` measures end to end.
"""
from __future__ import annotations

import ctypes
from typing import Any, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib, ops
from ._lib import check, ptr, stream_ptr


class _IstftFn(torch.autograd.Function):
    """iSTFT of a [B,S,F,T] complex tensor (optionally scaled by XrMM[B,F,T]) with a hand-written backward."""

    @staticmethod
    def forward(ctx, Xri: Tensor, scale: Optional[Tensor], n_fft: int, hop: int, length: int):
        # Xri: real view [B,S,F,T,2] with arbitrary strides (last dim stride 1)
        B, S, F, T, _ = Xri.shape
        sb, ss, sf, st, s2 = Xri.stride()
        assert s2 == 1
        y = ops.istft_strided(Xri, (sb, ss, sf, st), scale, B, S, F, T, n_fft, hop, length)
        ctx.save_for_backward(scale) if scale is not None else None
        ctx.meta = (B, S, F, T, n_fft, hop, scale is not None)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        B, S, F, T, n_fft, hop, has_scale = ctx.meta
        scale = ctx.saved_tensors[0] if has_scale else None
        d = torch.empty(B, S, F, T, 2, dtype=torch.float32, device=dy.device)
        ops.istft_bwd_strided(dy, scale, d, (S * F * T * 2, F * T * 2, T * 2, 2), B, S, F, T, n_fft, hop)
        return d, None, None, None, None


class STFT(nn.Module):
    """models/io/stft.py:21-103.  Periodic Hann window of length n_fft (the only configuration the reference YAMLs
    use); center=True, reflect padding, onesided, unnormalised — torch.stft defaults."""

    def __init__(self, n_fft: int, n_hop: int, win_len: Optional[int] = None, win: str = "hann_window") -> None:
        super().__init__()
        self.n_fft, self.n_hop, self.win_len = n_fft, n_hop, win_len if win_len is not None else n_fft
        self.repr = str((n_fft, n_hop, win, win_len))
        if win != "hann_window" or self.win_len != n_fft:
            raise NotImplementedError("nbss_b200.STFT supports the periodic Hann window with win_len == n_fft")
        self.register_buffer("window", torch.hann_window(n_fft))  # kept for state_dict compatibility

    def stft(self, x: Tensor) -> Tuple[Tensor, int]:
        shape = list(x.shape)
        x2 = x.reshape(1, -1, shape[-1]).float()
        X = ops.stft(x2, self.n_fft, self.n_hop)[0]  # [N, F, T]
        return X.reshape(shape[:-1] + list(X.shape[-2:])), shape[-1]

    def istft(self, X: Tensor, original_len: int = None) -> Tensor:
        shape = list(X.shape)
        Xr = torch.view_as_real(X.reshape(1, -1, *shape[-2:]))  # [1, N, F, T, 2], keeps strides where possible
        y = _IstftFn.apply(Xr, None, self.n_fft, self.n_hop, int(original_len))
        return y.reshape(shape[:-2] + [original_len])

    def forward(self, X: Tensor, original_len: int = None, inverse: bool = False) -> Any:
        return self.istft(X, original_len) if inverse else self.stft(X)

    def extra_repr(self) -> str:
        return self.repr

    def _load_from_state_dict(self, *args, **kwargs):  # models/io/stft.py:102-103: the window is never loaded
        return


class Norm(nn.Module):
    """models/io/norm.py:47-111 for mode='frequency', online=True (and 'none')."""

    def __init__(self, mode: Optional[str], online: bool = True) -> None:
        super().__init__()
        self.mode, self.online = mode, online
        if mode not in ("frequency", "none", None) or (mode == "frequency" and not online):
            raise NotImplementedError("nbss_b200.Norm implements mode='frequency' (online) and 'none'")

    def norm(self, X: Tensor, norm_paras: Any = None, ref_channel: int = None, eps: float = 1e-6):
        if self.mode in ("none", None):
            return X, (X[:, [ref_channel]].clone(), None)
        if norm_paras is not None:
            raise NotImplementedError("re-using norm_paras is not supported")
        B, C, F, T = X.shape
        Xc = X if X.is_contiguous() else X.contiguous()
        Xri = torch.view_as_real(Xc)
        xrmm = torch.empty(B, 1, F, T, dtype=torch.float32, device=X.device)
        xr = torch.empty(B, 1, F, T, 2, dtype=torch.float32, device=X.device)
        st = _lib.lib().nbss_norm_freq_online(ptr(Xri), B, C, ctypes.c_longlong(F * T), int(ref_channel), ctypes.c_float(eps),
                                              ptr(xrmm), ptr(xr), stream_ptr())
        check(st, "nbss_norm_freq_online")
        if Xc is not X:
            X.copy_(Xc)  # the reference normalises in place (norm.py:94)
        return X, (torch.view_as_complex(xr), xrmm)

    def inorm(self, X: Tensor, norm_paras: Any) -> Tensor:
        Xr, XrMM = norm_paras
        if XrMM is None:
            return X
        if X.requires_grad:  # keep autograd: a broadcast multiply is all the reference does here (norm.py:108)
            return X * XrMM
        B, S, F, T = X.shape
        Xc = torch.view_as_real(X.contiguous())
        Y = torch.empty_like(Xc)
        st = _lib.lib().nbss_inorm(ptr(Xc), ptr(Y), B, S, ctypes.c_longlong(F * T), ptr(XrMM.contiguous()), stream_ptr())
        check(st, "nbss_inorm")
        return torch.view_as_complex(Y)

    def forward(self, X: Tensor, norm_paras: Any = None, inverse: bool = False) -> Any:
        return self.inorm(X, norm_paras) if inverse else self.norm(X, norm_paras=norm_paras)

    def extra_repr(self) -> str:
        return f"{self.mode}, online={self.online}"


class SeparationPipeline(nn.Module):
    """TrainModule.forward (SharedTrainer.py:104-132) for a non-mask loss: wave [B,C,Ts] -> estimates [B,S,Ts].

    Three fused stages: (stft + Norm + pack) -> SpatialNet -> (unpack + inverse Norm + iSTFT).  Differentiable with
    respect to the network parameters (the iSTFT backward kernel feeds the network's backward)."""

    def __init__(self, arch: nn.Module, n_fft: int = 256, n_hop: int = 128, channels=None, ref_channel: int = 0):
        super().__init__()
        self.arch, self.n_fft, self.n_hop = arch, n_fft, n_hop
        self.channels, self.ref_channel = channels, ref_channel

    def forward(self, x: Tensor) -> Tensor:
        if self.channels is not None:
            x = x[:, self.channels]
            ref = list(self.channels).index(self.ref_channel)
        else:
            ref = self.ref_channel
        B, C, Ts = x.shape
        X, xrmm, _ = ops.stft_norm_pack(x.float(), self.n_fft, self.n_hop, ref)
        out = self.arch(X)  # [B,F,T,2S]
        _, F, T, S2 = out.shape
        S = S2 // 2
        out_v = out.reshape(B, F, T, S, 2).permute(0, 3, 1, 2, 4)  # [B,S,F,T,2] view, strides (F*T*2S, 2, T*2S, 2S, 1)
        return _IstftFn.apply(out_v, xrmm, self.n_fft, self.n_hop, Ts)
