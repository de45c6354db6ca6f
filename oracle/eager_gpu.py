"""The reference's OP-SET in PyTorch eager — BASELINE / TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``bench.py --impl torch-gpu`` (and the ``gpu_eager_baseline`` key of the main bench line) times this on the same B200:
it is "the kernel set to beat" of SURVEY.md §2.3 / §8(d) — the library kernels the unmodified reference would launch
(cuDNN grouped convolutions, cuBLAS GEMMs, fused SDPA attention, native LayerNorm / GroupNorm, cuFFT), called the way
the reference's modules call them, on parameters passed as the reference's state-dict (``oracle.spatialnet_oracle``
naming).  /root/reference itself cannot travel to the GPU box, so this file restates the module graph with
``torch.nn.functional`` calls, one per reference module:

    SpatialNetLayer._fconv      models/arch/SpatialNet.py:116-127   permute -> LayerNorm -> Conv1d(groups) -> PReLU -> permute
    SpatialNetLayer._full       :129-146                            LayerNorm -> Conv1d(1)+SiLU -> LinearGroup -> Conv1d(1)+SiLU
    SpatialNetLayer._tsa        :93-100                             LayerNorm -> nn.MultiheadAttention(need_weights=False) = SDPA
    SpatialNetLayer._tconvffn   :102-114                            LayerNorm -> Conv1d(1) SiLU (gconv SiLU) x3 with GroupNorm -> Conv1d(1)
    STFT / Norm / iSTFT         models/io/stft.py:49-97, norm.py:61-108   torch.stft / torch.istft

``tests/test_oracle_golden.py`` pins it to ``spatialnet_oracle`` (hence to the reference's golden vectors) on CPU.
Works on any device; under ``torch.autocast('cuda', torch.bfloat16)`` it is the reference's ``bf16-mixed`` precision.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _ln(x: Tensor, P, pre: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), P[pre + ".weight"], P[pre + ".bias"], 1e-5)


def _fconv(x: Tensor, P, pre: str, groups: int) -> Tensor:
    B, Fq, T, H = x.shape
    h = _ln(x, P, pre + ".0")                                   # LayerNorm(seq_last=True): transpose, LN over H, transpose
    h = h.permute(0, 2, 3, 1).reshape(B * T, H, Fq)
    h = F.conv1d(h, P[pre + ".1.weight"], P[pre + ".1.bias"], padding="same", groups=groups)
    h = F.prelu(h, P[pre + ".2.weight"])
    return h.reshape(B, T, H, Fq).permute(0, 3, 1, 2)


def _full(x: Tensor, P, pre: str) -> Tensor:
    B, Fq, T, H = x.shape
    h = _ln(x, P, pre + "norm_full").permute(0, 2, 3, 1).reshape(B * T, H, Fq)
    h = F.silu(F.conv1d(h, P[pre + "squeeze.0.weight"], P[pre + "squeeze.0.bias"]))
    h = torch.einsum("...gh,gkh->...gk", h, P[pre + "full.weight"]) + P[pre + "full.bias"]  # linear_group.py:31-33
    h = F.silu(F.conv1d(h, P[pre + "unsqueeze.0.weight"], P[pre + "unsqueeze.0.bias"]))
    return h.reshape(B, T, H, Fq).permute(0, 3, 1, 2)


def _tsa(x: Tensor, P, pre: str, num_heads: int) -> Tensor:
    B, Fq, T, H = x.shape
    h = _ln(x, P, pre + "norm_mhsa").reshape(B * Fq, T, H)
    # nn.MultiheadAttention.forward(x, x, x, need_weights=False) -> F.multi_head_attention_forward -> packed in-proj +
    # scaled_dot_product_attention + out-proj (torch/nn/functional.py)
    y, _ = F.multi_head_attention_forward(
        h.transpose(0, 1), h.transpose(0, 1), h.transpose(0, 1), H, num_heads, P[pre + "mhsa.in_proj_weight"],
        P[pre + "mhsa.in_proj_bias"], None, None, False, 0.0, P[pre + "mhsa.out_proj.weight"], P[pre + "mhsa.out_proj.bias"],
        training=False, need_weights=False)
    return y.transpose(0, 1).reshape(B, Fq, T, H)


def _tconvffn(x: Tensor, P, pre: str, groups: int) -> Tensor:
    B, Fq, T, H = x.shape
    t = pre + "tconvffn."
    h = _ln(x.transpose(-1, -2).reshape(B * Fq, H, T).transpose(-1, -2), P, t + "0").transpose(-1, -2)  # LayerNorm(seq_last=True)
    h = F.silu(F.conv1d(h, P[t + "1.weight"], P[t + "1.bias"]))
    h = F.silu(F.conv1d(h, P[t + "3.weight"], P[t + "3.bias"], padding="same", groups=groups))
    h = F.conv1d(h, P[t + "5.weight"], P[t + "5.bias"], padding="same", groups=groups)
    h = F.silu(F.group_norm(h, groups, P[t + "6.weight"], P[t + "6.bias"], 1e-5))
    h = F.silu(F.conv1d(h, P[t + "8.weight"], P[t + "8.bias"], padding="same", groups=groups))
    h = F.conv1d(h, P[t + "10.weight"], P[t + "10.bias"])
    return h.reshape(B, Fq, H, T).transpose(-1, -2)


def spatialnet_forward(P: Dict[str, Tensor], x: Tensor, cfg: dict) -> Tensor:
    """SpatialNet.forward (models/arch/SpatialNet.py:202-220) as eager library calls."""
    B, Fq, T, Cin = x.shape
    gf, gt = cfg["conv_groups"]
    h = F.conv1d(x.reshape(B * Fq, T, Cin).permute(0, 2, 1), P["encoder.weight"], P["encoder.bias"], padding="same").permute(0, 2, 1)
    h = h.reshape(B, Fq, T, -1)
    for i in range(cfg["num_layers"]):
        pre = f"layers.{i}."
        h = h + _fconv(h, P, pre + "fconv1", gf)
        h = h + _full(h, P, pre)
        h = h + _fconv(h, P, pre + "fconv2", gf)
        h = h + _tsa(h, P, pre, cfg["num_heads"])
        h = h + _tconvffn(h, P, pre, gt)
    return F.linear(h, P["decoder.weight"], P["decoder.bias"]).contiguous()


def io_forward(P: Dict[str, Tensor], x: Tensor, cfg: dict, n_fft: int = 256, n_hop: int = 128, ref_channel: int = 0) -> Tensor:
    """TrainModule.forward (SharedTrainer.py:104-132): wave [B,C,Ts] -> [B,S,Ts], fp32 framing (stft.py:59-61)."""
    B, C, Ts = x.shape
    dt = P["encoder.weight"].dtype  # fp32 (the reference); fp64 only to measure the fp32 noise floor in the tests
    win = torch.hann_window(n_fft, device=x.device, dtype=dt)
    with torch.autocast(x.device.type, enabled=False):
        X = torch.stft(x.to(dt).reshape(B * C, Ts), n_fft=n_fft, hop_length=n_hop, window=win, return_complex=True).reshape(B, C, n_fft // 2 + 1, -1)
        Xr = X[:, [ref_channel]].clone()
        XrMM = torch.abs(Xr) + 1e-6
        X = X / XrMM
        Xp = torch.view_as_real(X.permute(0, 2, 3, 1)).reshape(B, X.shape[2], X.shape[3], -1)
    out = spatialnet_forward(P, Xp, cfg)
    with torch.autocast(x.device.type, enabled=False):
        Fq, T = out.shape[1], out.shape[2]
        Y = torch.view_as_complex(out.to(dt).reshape(B, Fq, T, -1, 2)).permute(0, 3, 1, 2) * XrMM
        S = Y.shape[1]
        y = torch.istft(Y.reshape(B * S, Fq, T), n_fft=n_fft, hop_length=n_hop, window=win, length=Ts)
    return y.reshape(B, S, Ts)


def neg_si_sdr_pit2(est: Tensor, ref: Tensor) -> Tensor:
    """Mean over the batch of min over the two speaker permutations of the negative mean SI-SDR (models/io/loss.py:21-29,
    95-118 with torchmetrics' formula, see spatialnet_oracle.si_sdr)."""
    eps = torch.finfo(est.dtype).eps

    def si_sdr(p, t):
        alpha = ((p * t).sum(-1, keepdim=True) + eps) / ((t * t).sum(-1, keepdim=True) + eps)
        ts = alpha * t
        return 10 * torch.log10(((ts * ts).sum(-1) + eps) / (((ts - p) ** 2).sum(-1) + eps))

    a = -si_sdr(est, ref).mean(-1)
    b = -si_sdr(est[:, [1, 0]], ref).mean(-1)
    return torch.minimum(a, b).mean()
