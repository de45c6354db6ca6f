"""CPU restatement of the reference's NBC2 network (BASELINE.json configs[3]).

TEST INFRASTRUCTURE ONLY, like oracle/spatialnet_oracle.py: nothing in the product path may import this module.
Restates ``models/arch/NBC2.py`` of Audio-WestlakeU/NBSS in explicit torch-CPU math:

    NBC2.forward           :277-289   encoder Conv1d(k=5,'same') along T per (b,f) -> blocks -> Linear decoder
    NBC2Block.forward      :196-225   x += MHSA(LN(x));  x += linear2(conv(linear1(GBN(x))))
    NBC2Block.conv         :178-188   SiLU, gconv3, SiLU, gconv3, GBN (transposed), SiLU, gconv3, SiLU
    GroupBatchNorm.forward :111-145   statistics over (group = all F bins of one utterance, H) per frame t
                                      (share_along_sequence_dim=False), biased variance, eps 1e-5, affine per channel

Pinned by ``tests/golden/nbc2_small_f17_t12.npz`` (generated from the unmodified reference by tests/golden/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as Fn
from torch import Tensor

from .spatialnet_oracle import layer_norm, rel_l2  # noqa: F401  (same LayerNorm restatement; rel_l2 for the tests)

Params = Dict[str, Tensor]

NBC2_SMALL = dict(dim_input=16, dim_output=4, n_layers=8, encoder_kernel_size=5, dim_hidden=96, dim_ffn=192, num_freqs=257,
                  n_heads=2, conv_kernel_size=3, n_conv_groups=8)  # NBC2.py:294-311


def param_shapes(cfg: dict) -> Dict[str, Tuple[int, ...]]:
    H, Hf, Cin, Cout, ek, k, g = (cfg["dim_hidden"], cfg["dim_ffn"], cfg["dim_input"], cfg["dim_output"], cfg["encoder_kernel_size"],
                                  cfg["conv_kernel_size"], cfg["n_conv_groups"])
    s: Dict[str, Tuple[int, ...]] = {"encoder.weight": (H, Cin, ek), "encoder.bias": (H,)}
    for i in range(cfg["n_layers"]):
        p = f"sa_layers.{i}."
        s.update({p + "norm1.weight": (H,), p + "norm1.bias": (H,),
                  p + "self_attn.in_proj_weight": (3 * H, H), p + "self_attn.in_proj_bias": (3 * H,),
                  p + "self_attn.out_proj.weight": (H, H), p + "self_attn.out_proj.bias": (H,),
                  p + "norm2.weight": (H,), p + "norm2.bias": (H,),
                  p + "linear1.weight": (Hf, H), p + "linear1.bias": (Hf,)})
        for j in (1, 3, 6):
            s[p + f"conv.{j}.weight"] = (Hf, Hf // g, k)
            s[p + f"conv.{j}.bias"] = (Hf,)
        s.update({p + "conv.4.weight": (Hf, 1), p + "conv.4.bias": (Hf, 1),
                  p + "linear2.weight": (H, Hf), p + "linear2.bias": (H,)})
    s.update({"decoder.weight": (Cout, H), "decoder.bias": (Cout,)})
    return s


def synth_params(cfg: dict, seed: int, dtype=torch.float32) -> Params:
    """Deterministic synthetic parameters: fan-in-scaled normals for weights, small biases, norm gains around 1."""
    g = torch.Generator().manual_seed(seed)
    out: Params = {}
    for name, shp in param_shapes(cfg).items():
        is_norm = ".norm" in name or ".conv.4." in name
        if name.endswith("bias"):
            v = 0.1 * torch.randn(shp, generator=g)
        elif is_norm:
            v = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            v = torch.randn(shp, generator=g) / fan_in ** 0.5
        out[name] = v.to(dtype)
    return out


def group_batch_norm(x: Tensor, w: Tensor, b: Tensor, group: int, eps: float = 1e-5) -> Tensor:
    """x [B*F, T, H] (channel-last view); statistics over (F, H) per (b, t) (NBC2.py:118-128); w, b: [H]."""
    BF, T, H = x.shape
    v = x.reshape(BF // group, group, T, H)
    mean = v.mean(dim=(1, 3), keepdim=True)
    var = ((v - mean) ** 2).mean(dim=(1, 3), keepdim=True)
    return (((v - mean) / torch.sqrt(var + eps)) * w + b).reshape(BF, T, H)


def mhsa(x: Tensor, P: Params, pre: str, nh: int) -> Tensor:
    """nn.MultiheadAttention(batch_first, q=k=v=x) over T per row of [B*F, T, H] (NBC2.py:215-216)."""
    N, T, H = x.shape
    dh = H // nh
    qkv = x @ P[pre + "in_proj_weight"].t() + P[pre + "in_proj_bias"]
    q, k, v = (t.reshape(N, T, nh, dh).transpose(1, 2) for t in qkv.split(H, dim=-1))
    a = torch.softmax((q * dh ** -0.5) @ k.transpose(-1, -2), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(N, T, H)
    return o @ P[pre + "out_proj.weight"].t() + P[pre + "out_proj.bias"]


def _gconv(x: Tensor, w: Tensor, b: Tensor, groups: int) -> Tensor:
    """Conv1d(k, 'same' zero padding, groups) along T on a channel-last [N, T, C] tensor."""
    return Fn.conv1d(x.transpose(1, 2), w, b, padding=w.shape[-1] // 2, groups=groups).transpose(1, 2)


def block(x: Tensor, P: Params, pre: str, cfg: dict) -> Tensor:
    g, F_ = cfg["n_conv_groups"], cfg["num_freqs"]
    x = x + mhsa(layer_norm(x, P[pre + "norm1.weight"], P[pre + "norm1.bias"]), P, pre + "self_attn.", cfg["n_heads"])
    h = group_batch_norm(x, P[pre + "norm2.weight"], P[pre + "norm2.bias"], F_) @ P[pre + "linear1.weight"].t() + P[pre + "linear1.bias"]
    h = _gconv(Fn.silu(h), P[pre + "conv.1.weight"], P[pre + "conv.1.bias"], g)
    h = _gconv(Fn.silu(h), P[pre + "conv.3.weight"], P[pre + "conv.3.bias"], g)
    h = group_batch_norm(h, P[pre + "conv.4.weight"][:, 0], P[pre + "conv.4.bias"][:, 0], F_)
    h = _gconv(Fn.silu(h), P[pre + "conv.6.weight"], P[pre + "conv.6.bias"], g)
    return x + Fn.silu(h) @ P[pre + "linear2.weight"].t() + P[pre + "linear2.bias"]


def nbc2_forward(P: Params, x: Tensor, cfg: dict) -> Tensor:
    """x [B, F, T, dim_input] -> [B, F, T, dim_output] (NBC2.py:277-289)."""
    B, F_, T, Cin = x.shape
    assert F_ == cfg["num_freqs"]
    h = _gconv(x.reshape(B * F_, T, Cin), P["encoder.weight"], P["encoder.bias"], 1)
    for i in range(cfg["n_layers"]):
        h = block(h, P, f"sa_layers.{i}.", cfg)
    return (h @ P["decoder.weight"].t() + P["decoder.bias"]).reshape(B, F_, T, -1)
