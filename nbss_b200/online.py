"""Drop-in ``OnlineSpatialNet`` for ``models.arch.OnlineSpatialNet.OnlineSpatialNet`` with ``attention='mhsa(N)'`` — the causal /
streaming SpatialNet of BASELINE.json configs[4] — plus the chunked API the reference lacks.

* same constructor signature and state-dict keys (the reference registers its modules under the offline SpatialNet's names;
  ``CausalConv1d`` is an ``nn.Conv1d`` with the same parameters), so reference checkpoints load;
* ``forward(x[B,F,T,Cin]) -> [B,F,T,Cout]``: the function the reference executes — causal encoder / T-convs, GroupNorm over
  (24 channels x F) per frame, causal self-attention.  NOTE the reference's call ``mhsa(..., need_weights=False, attn_mask=mask,
  is_causal=True)`` makes torch drop the 251-frame window (oracle/online_oracle.py header), so ``forward`` attends to ALL past
  frames by default (``window=False``); ``window=True`` applies the window the model is named after;
* ``init_state(batch)`` / ``step(x_t[B,F,Cin], state) -> y_t[B,F,Cout]``: one 16 ms frame in, one frame out, constant memory: a
  key/value ring of ``attn_scope`` frames per layer, two previous frames per causal T-conv, four input frames of the encoder
  (csrc/online.cu).  ``forward`` itself is implemented by stepping, so offline and streaming results are the same numbers.

Supported: the SpatialNet-small layer (dim_hidden 96, dim_ffn 192, 4 heads, dim_squeeze 8, kernel (5,3), groups (8,8), norms
LN/LN/GN/LN/LN/LN, rope False), num_freqs <= 256.  The Mamba / retention variants need ``mamba_ssm`` / have no oracle here.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib, ops
from ._lib import check, ptr, stream_ptr
from .spatialnet import SpatialNetLayer

_MAX_SCOPE = 2048


class OnlineState:
    """Everything a stream of frames carries from one step to the next (all device tensors; constant size)."""

    def __init__(self, net: "OnlineSpatialNet", batch: int, device, scope: int):
        L, F, Cin = len(net.layers), net.num_freqs, net.dim_input
        R = batch * F
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        self.batch, self.scope, self.R = batch, scope, R
        self.pos = torch.zeros(1, dtype=torch.int32, device=device)
        self.enc = z(R, 4, Cin)
        self.kcache = [z(R, scope, 96) for _ in range(L)]
        self.vcache = [z(R, scope, 96) for _ in range(L)]
        self.st = [[z(R, 2, 192) for _ in range(3)] for _ in range(L)]
        self.h = z(batch, F, 1, 96)
        self.c2, self.part, self.stats = z(R, 192), z(R, 8, 2), z(batch, 8, 2)
        self.graph, self.x_in, self.y_out = None, None, None  # set by OnlineSpatialNet.capture_step


class OnlineSpatialNet(nn.Module):
    def __init__(self, dim_input: int, dim_output: int, num_layers: int, dim_squeeze: int, num_freqs: int, encoder_kernel_size: int = 5,
                 dim_hidden: int = 192, dim_ffn: int = 384, num_heads: int = 2, dropout: Tuple[float, float, float] = (0, 0, 0),
                 kernel_size: Tuple[int, int] = (5, 3), conv_groups: Tuple[int, int] = (8, 8),
                 norms: List[str] = ("LN", "LN", "GN", "LN", "LN", "LN"), padding: str = "zeros", full_share: int = 0,
                 attention: str = "mhsa(251)", decay=5, chunkwise_recurrent: bool = True, rope=False):
        super().__init__()
        bad = []
        if (dim_hidden, dim_ffn, num_heads, dim_squeeze) != (96, 192, 4, 8): bad.append("dim_hidden/dim_ffn/num_heads/dim_squeeze != 96/192/4/8")
        if tuple(kernel_size) != (5, 3) or tuple(conv_groups) != (8, 8) or encoder_kernel_size != 5: bad.append("kernel sizes / groups")
        if [n.upper() for n in norms] != ["LN", "LN", "GN", "LN", "LN", "LN"]: bad.append("norms")
        if any(d > 0 for d in dropout) or padding != "zeros" or rope not in (False,): bad.append("dropout / padding / rope")
        if not attention.startswith("mhsa("): bad.append(f"attention {attention!r} (only mhsa(N): retention / Mamba have no oracle here)")
        if num_freqs > 256: bad.append("num_freqs > 256")
        if bad:
            raise NotImplementedError("nbss_b200.OnlineSpatialNet: unsupported " + ", ".join(bad))
        arg = attention[5:-1]
        self.attn_scope = _MAX_SCOPE if arg == "inf" else int(arg)
        if not 1 <= self.attn_scope <= _MAX_SCOPE:
            raise NotImplementedError(f"attention scope {self.attn_scope} (1..{_MAX_SCOPE})")
        self.dim_input, self.num_freqs = dim_input, num_freqs
        self.encoder = nn.Conv1d(dim_input, dim_hidden, encoder_kernel_size)  # CausalConv1d: same parameters, left padding at run time
        full, layers = None, []
        for l in range(num_layers):
            layer = SpatialNetLayer(dim_hidden, dim_ffn, dim_squeeze, num_freqs, num_heads, kernel_size, conv_groups, full=full if l > full_share else None)
            full = layer.full
            layers.append(layer)
        self.layers = nn.ModuleList(layers)
        self.decoder = nn.Linear(dim_hidden, dim_output)
        self._packed = None
        self._packed_key = None

    # ------------------------------------------------------------------------------------------------ weights
    def _params(self) -> Dict[str, Tensor]:
        return {n: p.detach() for n, p in self.named_parameters(remove_duplicate=False)}

    def _pack(self, P: Dict[str, Tensor]):
        key = tuple((p.data_ptr(), p._version) for p in P.values())
        if self._packed is not None and key == self._packed_key:  # inference: weights are constants, also under graph capture
            return self._packed
        dev = P["encoder.weight"].device

        def tr(w: Tensor) -> Tensor:  # [rows, cols] -> [cols, rows] (outputs contiguous: coalesced across the output threads)
            w2 = ops._f32c(w.reshape(w.shape[0], -1))
            out = torch.empty(w2.shape[1], w2.shape[0], dtype=torch.float32, device=dev)
            check(ops._K("nbss_transpose")(ptr(w2), ptr(out), w2.shape[0], w2.shape[1], stream_ptr()), "nbss_transpose")
            return out

        enc = torch.empty(5 * self.dim_input, 96, dtype=torch.float32, device=dev)
        check(ops._K("nbss_online_pack_encoder")(ptr(ops._f32c(P["encoder.weight"])), ptr(enc), self.dim_input, stream_ptr()), "nbss_online_pack_encoder")
        layers = []
        for i in range(len(self.layers)):
            pre = f"layers.{i}."
            layers.append(dict(WinT=tr(P[pre + "mhsa.in_proj_weight"]), WoT=tr(P[pre + "mhsa.out_proj.weight"]),
                               W1T=tr(P[pre + "tconvffn.1.weight"]), W2T=tr(P[pre + "tconvffn.10.weight"]),
                               Wc1T=tr(P[pre + "tconvffn.3.weight"]), Wc2T=tr(P[pre + "tconvffn.5.weight"]), Wc3T=tr(P[pre + "tconvffn.8.weight"]),
                               f1=ops.fconv_pack(P[pre + "fconv1.1.weight"]), f2=ops.fconv_pack(P[pre + "fconv2.1.weight"]),
                               lg=ops.lg_pack(P[pre + "full.weight"])))
        self._packed, self._packed_key = dict(enc=enc, layers=layers), key
        return self._packed

    # ------------------------------------------------------------------------------------------------ streaming API
    def init_state(self, batch: int, device=None, scope: Optional[int] = None) -> OnlineState:
        device = device if device is not None else self.encoder.weight.device
        if torch.device(device).type != "cuda":
            raise _lib.NbssError("nbss_b200.OnlineSpatialNet runs on CUDA only (there is no CPU path)")
        return OnlineState(self, batch, device, scope if scope is not None else self.attn_scope)

    @torch.no_grad()
    def capture_step(self, state: OnlineState) -> OnlineState:
        """Records the ~60 launches of one step of this stream into a CUDA graph (every pointer of a step is fixed: the state, the
        ring position lives on the device); afterwards `step` copies the frame into the graph's input and replays.  The returned
        frame is then a static buffer, overwritten by the next step.  Capturing executes nothing: the state is not advanced."""
        dev = state.h.device
        self._pack(self._params())
        ops.device_err_flag(dev)
        state.x_in = torch.zeros(state.batch, self.num_freqs, self.dim_input, dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            state.y_out = self._step_launches(state.x_in, state)
        state.graph = g
        return state

    @torch.no_grad()
    def step(self, x_t: Tensor, state: OnlineState) -> Tensor:
        """One frame: x_t [B,F,Cin] -> y_t [B,F,Cout]; `state` is updated in place."""
        B, F, Cin = x_t.shape
        assert B == state.batch and F == self.num_freqs and Cin == self.dim_input, (x_t.shape, state.batch)
        with _lib.device_of(state.h):
            if state.graph is not None:
                state.x_in.copy_(x_t, non_blocking=True)
                state.graph.replay()
                return state.y_out
            return self._step_launches(x_t, state)

    def _step_launches(self, x_t: Tensor, state: OnlineState) -> Tensor:
        B, F, Cin = x_t.shape
        P = self._params()
        W = self._pack(P)
        R, h, sp = state.R, state.h, stream_ptr
        err = ops.device_err_flag(x_t.device)
        check(ops._K("nbss_online_encoder_step")(ptr(ops._f32c(x_t.float())), ptr(state.enc), ptr(W["enc"]), ptr(ops._f32c(P["encoder.bias"])),
                                                 ptr(h), R, Cin, sp()), "nbss_online_encoder_step")
        for i, Wl in enumerate(W["layers"]):
            pre = f"layers.{i}."
            ops.fconv_tc_fwd(h, P, pre + "fconv1", Wl["f1"], out=h)
            ops.full_fwd_tc(h, P, pre, Wl["lg"], out=h)
            ops.fconv_tc_fwd(h, P, pre + "fconv2", Wl["f2"], out=h)
            check(ops._K("nbss_online_attn_step")(ptr(h), R, ptr(ops._f32c(P[pre + "norm_mhsa.weight"])), ptr(ops._f32c(P[pre + "norm_mhsa.bias"])),
                                                  ptr(Wl["WinT"]), ptr(ops._f32c(P[pre + "mhsa.in_proj_bias"])), ptr(Wl["WoT"]),
                                                  ptr(ops._f32c(P[pre + "mhsa.out_proj.bias"])), ptr(state.kcache[i]), ptr(state.vcache[i]),
                                                  ptr(state.pos), state.scope, sp()), "nbss_online_attn_step")
            t = pre + "tconvffn."
            check(ops._K("nbss_online_ffn_a_step")(ptr(h), R, ptr(ops._f32c(P[t + "0.weight"])), ptr(ops._f32c(P[t + "0.bias"])), ptr(Wl["W1T"]),
                                                   ptr(ops._f32c(P[t + "1.bias"])), ptr(Wl["Wc1T"]), ptr(ops._f32c(P[t + "3.bias"])),
                                                   ptr(Wl["Wc2T"]), ptr(ops._f32c(P[t + "5.bias"])), ptr(state.st[i][0]),
                                                   ptr(state.st[i][1]), ptr(state.c2), ptr(state.part), sp()), "nbss_online_ffn_a_step")
            check(ops._K("nbss_online_gn_stats")(ptr(state.part), B, F, ptr(state.stats), sp()), "nbss_online_gn_stats")
            check(ops._K("nbss_online_ffn_b_step")(ptr(h), R, F, ptr(state.c2), ptr(state.stats), ptr(ops._f32c(P[t + "6.weight"])),
                                                   ptr(ops._f32c(P[t + "6.bias"])), ptr(Wl["Wc3T"]), ptr(ops._f32c(P[t + "8.bias"])),
                                                   ptr(Wl["W2T"]), ptr(ops._f32c(P[t + "10.bias"])), ptr(state.st[i][2]), sp()), "nbss_online_ffn_b_step")
        y = ops.decoder_fwd(h, P)  # [B,F,1,Cout]
        check(ops._K("nbss_online_advance")(ptr(state.pos), sp()), "nbss_online_advance")
        self._last_err = err
        return y[:, :, 0]

    @torch.no_grad()
    def forward(self, x: Tensor, inference: bool = False, return_attn_score: bool = False, window: bool = False):
        """x [B,F,T,Cin] -> [B,F,T,Cout] by stepping through the frames (see the module docstring for `window`)."""
        if not x.is_cuda:
            raise _lib.NbssError("nbss_b200.OnlineSpatialNet runs on CUDA tensors only (there is no CPU path)")
        B, F, T, _ = x.shape
        scope = self.attn_scope if window else max(T, 1)
        if scope > _MAX_SCOPE:
            raise NotImplementedError(f"forward(window=False) attends to all {T} past frames like the reference; supported up to {_MAX_SCOPE} frames "
                                      "(use window=True or step())")
        state = self.init_state(B, x.device, scope=scope)
        xs = x.detach().float().permute(2, 0, 1, 3).contiguous()  # [T,B,F,Cin]
        ys = [self.step(xs[t], state).clone() for t in range(T)]
        y = torch.stack(ys, dim=2).contiguous()
        return (y, [None] * len(self.layers)) if return_attn_score else y

    def check_device_errors(self) -> None:
        for f in ops._ERR_FLAGS.values():
            ops.check_err_flag(f, "nbss_b200 kernel")
