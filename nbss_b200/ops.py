"""Thin Python wrappers over the C-ABI entry points of libnbss_b200.so (include/nbss_b200.h).

Each wrapper takes torch CUDA tensors, passes raw device pointers + sizes + the current CUDA stream, and raises
``NbssError`` on a non-zero status.  No arithmetic happens here and there is no fallback path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

FMT_F16, FMT_BF16, FMT_TF32 = 0, 1, 2
Tensor = torch.Tensor

# Instrumentation used by bench.py: LAUNCHES counts kernel launches issued through this module; when TIMING is a dict,
# every C-ABI call is bracketed by CUDA events on the launching (current) stream.
LAUNCHES = 0
TIMING = None
_NLAUNCH = {"nbss_nbc2_block": 5, "nbss_full_fwd": 3, "nbss_full_bwd": 4, "nbss_full_fwd_tc": 3, "nbss_full_bwd_tc": 4, "nbss_ffn_wgrad": 3, "nbss_mhsa_bwd": 2, "nbss_istft": 2, "nbss_sisdr_pit_fwd": 2, "nbss_clip_adam": 2,
            "nbss_mhsa_fwd_long": 2, "nbss_ffn_fwd_long": 3, "nbss_grad_prescale": 2}


_KCACHE = {}

# Weight-gradient kernels may run on a side stream (set by spatialnet.Engine.backward): they depend only on the data
# gradient kernel that precedes them, so they fill the SMs that the persistent slab kernels leave idle in their last
# partial wave (at 4 utterances per GPU: 516 slabs on 148 SMs = 3.5 waves).  SIDE_KEEP holds the tensors the side stream
# still reads until the engine joins the streams.
WGRAD_STREAM: Optional["torch.cuda.Stream"] = None
SIDE_KEEP: list = []


class _side_stream:
    """Context: run the enclosed launches on WGRAD_STREAM (after everything queued on the current stream so far)."""

    def __init__(self, keep):
        self.keep = keep
        self.ctx = None

    def __enter__(self):
        if WGRAD_STREAM is not None:
            WGRAD_STREAM.wait_stream(torch.cuda.current_stream())
            SIDE_KEEP.append(self.keep)
            self.ctx = torch.cuda.stream(WGRAD_STREAM)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _K(name: str):
    c = _KCACHE.get(name)
    if c is None:
        c = _KCACHE[name] = _make_call(name)
    return c


def _make_call(name: str):
    fn = getattr(_lib.lib(), name)

    def call(*args):
        global LAUNCHES
        LAUNCHES += _NLAUNCH.get(name, 1)
        if TIMING is None:
            return fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = fn(*args)
        e1.record()
        TIMING.setdefault(name[5:], []).append((e0, e1))
        return st

    return call


def _f32c(t: Tensor) -> Tensor:
    assert t.is_cuda and t.dtype == torch.float32, (t.device, t.dtype)
    return t if t.is_contiguous() else t.contiguous()


_ERR_FLAGS = {}


def device_err_flag(device) -> Tensor:
    """One persistent device int per device, shared by every tensor-core kernel launch (set to 0x7001 by a kernel whose
    mbarrier wait timed out; it stays set, so checking it once after a step is enough)."""
    f = _ERR_FLAGS.get(device)
    if f is None:
        f = _ERR_FLAGS[device] = torch.zeros(1, dtype=torch.int32, device=device)
    return f


def check_err_flag(flag: Tensor, what: str) -> None:
    v = int(flag.item())
    if v != 0:
        raise _lib.NbssError(f"{what}: device-side error flag {v:#x} (mbarrier wait timed out)")


def untile(t: Tensor, nslab: int, T: int) -> Tensor:
    """16-bit intermediates are stored in the slab-tile layout [slab][C/8][T][8] (csrc/slab.cuh); returns [nslab*T, C]."""
    C = t.numel() // (nslab * T)
    return t.reshape(nslab, C // 8, T, 8).permute(0, 2, 1, 3).reshape(nslab * T, C)


def layer_image_bytes() -> int:
    L = _lib.lib()
    L.nbss_layer_image_bytes.restype = ctypes.c_uint
    return int(_K("nbss_layer_image_bytes")())


def pack_layer_weights(P: Dict[str, Tensor], pre: str, img: Optional[Tensor] = None, fwd_fmt: int = FMT_F16,
                       bwd_fmt: int = FMT_F16) -> Tensor:
    """Builds the UMMA weight images of one SpatialNet layer (pack.cu) from its fp32 parameters."""
    dev = P[pre + "tconvffn.1.weight"].device
    if img is None:
        img = torch.empty(layer_image_bytes(), dtype=torch.uint8, device=dev)
    t = pre + "tconvffn."
    st = _K("nbss_pack_layer_weights")(
        ptr(_f32c(P[t + "1.weight"])), ptr(_f32c(P[t + "3.weight"])), ptr(_f32c(P[t + "5.weight"])),
        ptr(_f32c(P[t + "8.weight"])), ptr(_f32c(P[t + "10.weight"])), ptr(_f32c(P[pre + "mhsa.in_proj_weight"])),
        ptr(_f32c(P[pre + "mhsa.out_proj.weight"])), ptr(img), fwd_fmt, bwd_fmt, stream_ptr())
    check(st, "nbss_pack_layer_weights")
    return img


def ffn_fwd(x: Tensor, P: Dict[str, Tensor], pre: str, img: Tensor, save: bool = False, fmt: int = FMT_F16,
            out: Optional[Tensor] = None):
    """x: [B,F,T,96] fp32 -> y = x + tconvffn(x).  With save=True also returns the fp16 pre-activations
    (a1, c1, c2, c3) [B*F*T,192] and GroupNorm stats [B*F,8,2] needed by the backward kernels."""
    x = _f32c(x)
    B, F, T, H = x.shape
    assert H == 96
    y = torch.empty_like(x) if out is None else out
    n = B * F * T
    t = pre + "tconvffn."
    if T > 256:  # inference on long utterances: the sub-block cut at its GroupNorm and tiled over T (ffn_fwd.cu MODE 3 / 4)
        if save:
            raise NotImplementedError("nbss_b200: training needs T <= 256 frames per utterance (the backward kernels hold one (b,f) "
                                      "slab per CTA); longer inputs run in inference mode (torch.no_grad())")
        if fmt != FMT_F16:
            raise NotImplementedError("the long-sequence kernels are built for fp16 operands")
        c2 = torch.empty(B * F * 24 * T * 8, dtype=torch.float16, device=x.device)
        part = torch.empty(B * F * int(_lib.lib().nbss_ffn_long_chunks(T, 0)) * 16, dtype=torch.float32, device=x.device)
        stats = torch.empty(B * F * 16, dtype=torch.float32, device=x.device)
        err = device_err_flag(x.device)
        check(_K("nbss_ffn_fwd_long")(
            ptr(x), ptr(y), B * F, T, ptr(_f32c(P[t + "0.weight"])), ptr(_f32c(P[t + "0.bias"])), ptr(_f32c(P[t + "1.bias"])),
            ptr(_f32c(P[t + "3.bias"])), ptr(_f32c(P[t + "5.bias"])), ptr(_f32c(P[t + "8.bias"])), ptr(_f32c(P[t + "6.weight"])),
            ptr(_f32c(P[t + "6.bias"])), ptr(_f32c(P[t + "10.bias"])), ptr(img), ptr(c2), ptr(part), ptr(stats), fmt, ptr(err),
            stream_ptr()), "nbss_ffn_fwd_long")
        return y, err
    saves = [torch.empty(n, 192, dtype=torch.float16, device=x.device) for _ in range(4)] if save else [None] * 4
    stats = torch.empty(B * F, 8, 2, dtype=torch.float32, device=x.device) if save else None
    ln_stats = torch.empty(n, 2, dtype=torch.float32, device=x.device) if save else None
    err = device_err_flag(x.device)
    st = _K("nbss_ffn_fwd")(
        ptr(x), ptr(y), B * F, T, ptr(_f32c(P[t + "0.weight"])), ptr(_f32c(P[t + "0.bias"])), ptr(_f32c(P[t + "1.bias"])),
        ptr(_f32c(P[t + "3.bias"])), ptr(_f32c(P[t + "5.bias"])), ptr(_f32c(P[t + "8.bias"])), ptr(_f32c(P[t + "6.weight"])),
        ptr(_f32c(P[t + "6.bias"])), ptr(_f32c(P[t + "10.bias"])), ptr(img), ptr(saves[0]), ptr(saves[1]), ptr(saves[2]),
        ptr(saves[3]), ptr(stats), ptr(ln_stats), fmt, ptr(err), stream_ptr())
    check(st, "nbss_ffn_fwd")
    if save:
        return y, saves + [ln_stats], stats, err
    return y, err


def nbc2_pack_block(P: Dict[str, Tensor], pre: str, img: Optional[Tensor] = None, fmt: int = FMT_F16) -> Tensor:
    """UMMA weight images of one NBC2Block (models/arch/NBC2.py:152-194): the same image layout as a SpatialNet layer's
    narrow-band block (linear1 / conv.{1,3,6} / linear2 / self_attn take the places of tconvffn.{1,3,5,8,10} / mhsa)."""
    dev = P[pre + "linear1.weight"].device
    if img is None:
        img = torch.empty(layer_image_bytes(), dtype=torch.uint8, device=dev)
    st = _K("nbss_pack_layer_weights")(
        ptr(_f32c(P[pre + "linear1.weight"])), ptr(_f32c(P[pre + "conv.1.weight"])), ptr(_f32c(P[pre + "conv.3.weight"])),
        ptr(_f32c(P[pre + "conv.6.weight"])), ptr(_f32c(P[pre + "linear2.weight"])), ptr(_f32c(P[pre + "self_attn.in_proj_weight"])),
        ptr(_f32c(P[pre + "self_attn.out_proj.weight"])), ptr(img), fmt, fmt, stream_ptr())
    check(st, "nbss_pack_layer_weights")
    return img


def nbc2_block_fwd(x: Tensor, P: Dict[str, Tensor], pre: str, img: Tensor, num_heads: int = 2, fmt: int = FMT_F16,
                   ws: Optional[dict] = None) -> Tensor:
    """One NBC2Block in place on the stream x [B,F,T,96] (inference): x += MHSA(LN(x)); x += FF(GBN(x)) (NBC2.py:196-225).
    Five launches: attention (+ GroupBatchNorm partials), statistics, T-ConvFFN part A, statistics, part B.  `ws` caches the
    scratch tensors across blocks."""
    x = _f32c(x)
    B, F, T, H = x.shape
    assert H == 96
    n, nslab = B * F * T, B * F
    ws = {} if ws is None else ws
    key = (B, F, T, x.device)
    if ws.get("key") != key:
        ws.clear()
        ws.update(key=key, part1=torch.empty(n, 2, dtype=torch.float32, device=x.device),
                  part2=torch.empty(n, 2, 2, dtype=torch.float32, device=x.device),
                  stats=torch.empty(B * T, 2, dtype=torch.float32, device=x.device),
                  c2=torch.empty(n, 192, dtype=torch.float16, device=x.device))
    err = device_err_flag(x.device)
    st = _K("nbss_mhsa_fwd_nh")(
        ptr(x), ptr(x), nslab, T, ptr(_f32c(P[pre + "norm1.weight"])), ptr(_f32c(P[pre + "norm1.bias"])),
        ptr(_f32c(P[pre + "self_attn.in_proj_bias"])), ptr(_f32c(P[pre + "self_attn.out_proj.bias"])), ptr(img), ptr(None), ptr(None),
        ptr(None), ptr(None), ptr(ws["part1"]), num_heads, fmt, ptr(err), stream_ptr())
    check(st, "nbss_mhsa_fwd_nh")
    check(_K("nbss_gbn_reduce")(ptr(ws["part1"]), B, F, T, 1, ctypes.c_longlong(F * 96), ctypes.c_float(1e-5), ptr(ws["stats"]), stream_ptr()),
          "nbss_gbn_reduce")
    st = _K("nbss_nbc2_ffn_a")(
        ptr(x), nslab, T, F, ptr(ws["stats"]), ptr(_f32c(P[pre + "norm2.weight"])), ptr(_f32c(P[pre + "norm2.bias"])),
        ptr(_f32c(P[pre + "linear1.bias"])), ptr(_f32c(P[pre + "conv.1.bias"])), ptr(_f32c(P[pre + "conv.3.bias"])), ptr(img),
        ptr(ws["c2"]), ptr(ws["part2"]), fmt, ptr(err), stream_ptr())
    check(st, "nbss_nbc2_ffn_a")
    check(_K("nbss_gbn_reduce")(ptr(ws["part2"]), B, F, T, 2, ctypes.c_longlong(F * 192), ctypes.c_float(1e-5), ptr(ws["stats"]), stream_ptr()),
          "nbss_gbn_reduce")
    st = _K("nbss_nbc2_ffn_b")(
        ptr(x), ptr(x), nslab, T, F, ptr(ws["stats"]), ptr(_f32c(P[pre + "conv.4.weight"])), ptr(_f32c(P[pre + "conv.4.bias"])),
        ptr(_f32c(P[pre + "conv.6.bias"])), ptr(_f32c(P[pre + "linear2.bias"])), ptr(img), ptr(ws["c2"]), fmt, ptr(err), stream_ptr())
    check(st, "nbss_nbc2_ffn_b")
    return x


def mhsa_fwd(x: Tensor, P: Dict[str, Tensor], pre: str, img: Tensor, save: bool = False, fmt: int = FMT_F16,
             out: Optional[Tensor] = None):
    """x: [B,F,T,96] fp32 -> y = x + MHSA(LN(x)) over T per (b,f).  With save=True also returns fp16 (scaled q|k|v)
    [n,288], O [n,96] and the log2-domain logsumexp [B*F,4,T]."""
    x = _f32c(x)
    B, F, T, H = x.shape
    assert H == 96
    y = torch.empty_like(x) if out is None else out
    n = B * F * T
    if T > 256:  # inference on long utterances: K | V pass + chunked flash-style attention (mhsa_fwd.cu LONG = 1 / 2)
        if save:
            raise NotImplementedError("nbss_b200: training needs T <= 256 frames per utterance (the backward kernels hold one (b,f) "
                                      "slab per CTA); longer inputs run in inference mode (torch.no_grad())")
        if fmt != FMT_F16:
            raise NotImplementedError("the long-sequence kernels are built for fp16 operands")
        kv = torch.empty(n, 288, dtype=torch.float16, device=x.device)
        err = device_err_flag(x.device)
        check(_K("nbss_mhsa_fwd_long")(
            ptr(x), ptr(y), B * F, T, ptr(_f32c(P[pre + "norm_mhsa.weight"])), ptr(_f32c(P[pre + "norm_mhsa.bias"])),
            ptr(_f32c(P[pre + "mhsa.in_proj_bias"])), ptr(_f32c(P[pre + "mhsa.out_proj.bias"])), ptr(img), ptr(kv), fmt, ptr(err),
            stream_ptr()), "nbss_mhsa_fwd_long")
        return y, err
    qkv = torch.empty(n, 288, dtype=torch.float16, device=x.device) if save else None
    o = torch.empty(n, 96, dtype=torch.float16, device=x.device) if save else None
    lse = torch.empty(B * F, 4, T, dtype=torch.float32, device=x.device) if save else None
    ln_stats = torch.empty(n, 2, dtype=torch.float32, device=x.device) if save else None
    err = device_err_flag(x.device)
    st = _K("nbss_mhsa_fwd")(
        ptr(x), ptr(y), B * F, T, ptr(_f32c(P[pre + "norm_mhsa.weight"])), ptr(_f32c(P[pre + "norm_mhsa.bias"])),
        ptr(_f32c(P[pre + "mhsa.in_proj_bias"])), ptr(_f32c(P[pre + "mhsa.out_proj.bias"])), ptr(img), ptr(qkv), ptr(o),
        ptr(lse), ptr(ln_stats), fmt, ptr(err), stream_ptr())
    check(st, "nbss_mhsa_fwd")
    if save:
        return y, (qkv, o, lse, ln_stats), err
    return y, err


# ------------------------------------------------------------------------------------------------ cross-band (fp32)
def fconv_fwd(x: Tensor, P, pre: str, out: Optional[Tensor] = None) -> Tensor:
    """y = x + PReLU(gconv_F(LN(x))); pre = 'layers.i.fconv1' / '...fconv2'."""
    x = _f32c(x)
    B, F, T, H = x.shape
    assert H == 96
    y = torch.empty_like(x) if out is None else out
    st = _K("nbss_fconv_fwd")(ptr(x), ptr(y), B, F, T, ptr(_f32c(P[pre + ".0.weight"])), ptr(_f32c(P[pre + ".0.bias"])),
                          ptr(_f32c(P[pre + ".1.weight"])), ptr(_f32c(P[pre + ".1.bias"])), ptr(_f32c(P[pre + ".2.weight"])),
                          stream_ptr())
    check(st, "nbss_fconv_fwd")
    return y


def fconv_bwd(x: Tensor, dy: Tensor, P, pre: str, G) -> Tensor:
    """Returns dx; accumulates parameter gradients into the fp32 tensors G[name] (same keys as P)."""
    x, dy = _f32c(x), _f32c(dy)
    B, F, T, H = x.shape
    dx = torch.empty_like(x)
    st = _K("nbss_fconv_bwd")(ptr(x), ptr(dy), ptr(dx), B, F, T, ptr(_f32c(P[pre + ".0.weight"])), ptr(_f32c(P[pre + ".0.bias"])),
                          ptr(_f32c(P[pre + ".1.weight"])), ptr(_f32c(P[pre + ".1.bias"])), ptr(_f32c(P[pre + ".2.weight"])),
                          ptr(G[pre + ".1.weight"]), ptr(G[pre + ".1.bias"]), ptr(G[pre + ".2.weight"]),
                          ptr(G[pre + ".0.weight"]), ptr(G[pre + ".0.bias"]), stream_ptr())
    check(st, "nbss_fconv_bwd")
    return dx


def full_fwd(x: Tensor, P, pre: str, out: Optional[Tensor] = None):
    """y = x + full-band branch; returns (y, s, u) with s,u [B,T,8,F] kept for backward."""
    x = _f32c(x)
    B, F, T, H = x.shape
    y = torch.empty_like(x) if out is None else out
    s = torch.empty(B, T, 8, F, dtype=torch.float32, device=x.device)
    u = torch.empty_like(s)
    st = _K("nbss_full_fwd")(ptr(x), ptr(y), ptr(s), ptr(u), B, F, T, ptr(_f32c(P[pre + "norm_full.weight"])),
                         ptr(_f32c(P[pre + "norm_full.bias"])), ptr(_f32c(P[pre + "squeeze.0.weight"])),
                         ptr(_f32c(P[pre + "squeeze.0.bias"])), ptr(_f32c(P[pre + "full.weight"])),
                         ptr(_f32c(P[pre + "full.bias"])), ptr(_f32c(P[pre + "unsqueeze.0.weight"])),
                         ptr(_f32c(P[pre + "unsqueeze.0.bias"])), stream_ptr())
    check(st, "nbss_full_fwd")
    return y, s, u


def full_bwd(x: Tensor, dy: Tensor, s: Tensor, u: Tensor, P, pre: str, G) -> Tensor:
    x, dy = _f32c(x), _f32c(dy)
    B, F, T, H = x.shape
    dx = torch.empty_like(x)
    ws = torch.empty(2 * s.numel(), dtype=torch.float32, device=x.device)
    st = _K("nbss_full_bwd")(ptr(x), ptr(dy), ptr(dx), ptr(s), ptr(u), ptr(ws), B, F, T, ptr(_f32c(P[pre + "norm_full.weight"])),
                         ptr(_f32c(P[pre + "norm_full.bias"])), ptr(_f32c(P[pre + "squeeze.0.weight"])),
                         ptr(_f32c(P[pre + "squeeze.0.bias"])), ptr(_f32c(P[pre + "full.weight"])),
                         ptr(_f32c(P[pre + "unsqueeze.0.weight"])), ptr(_f32c(P[pre + "unsqueeze.0.bias"])),
                         ptr(G[pre + "norm_full.weight"]), ptr(G[pre + "norm_full.bias"]), ptr(G[pre + "squeeze.0.weight"]),
                         ptr(G[pre + "squeeze.0.bias"]), ptr(G[pre + "full.weight"]), ptr(G[pre + "full.bias"]),
                         ptr(G[pre + "unsqueeze.0.weight"]), ptr(G[pre + "unsqueeze.0.bias"]), stream_ptr())
    check(st, "nbss_full_bwd")
    return dx


def lg_image_bytes(F: int) -> int:
    n = int(_lib.lib().nbss_lg_image_bytes(F))
    if n == 0:
        raise _lib.NbssError(f"nbss_lg_image_bytes: unsupported num_freqs {F} (tensor-core LinearGroup needs F <= 256)")
    return n


def lg_pack(Wf: Tensor, img: Optional[Tensor] = None, fmt: int = FMT_F16) -> Tensor:
    """UMMA operand images of the LinearGroup weight full.weight [8,F,F] (fullband_tc.cu)."""
    F = Wf.shape[-1]
    if img is None:
        img = torch.empty(lg_image_bytes(F), dtype=torch.uint8, device=Wf.device)
    check(_K("nbss_lg_pack")(ptr(_f32c(Wf)), ptr(img), F, fmt, stream_ptr()), "nbss_lg_pack")
    return img


def lg_tc_apply(x: Tensor, img: Tensor, bias: Optional[Tensor], mode: int, fmt: int = FMT_F16) -> Tensor:
    """LinearGroup on tensor cores: x [M,8,F] fp32 -> mode 0: x W^T + bias, mode 1: x W (data gradient)."""
    x = _f32c(x)
    M, G8, F = x.shape
    assert G8 == 8
    out = torch.empty_like(x)
    err = device_err_flag(x.device)
    check(_K("nbss_lg_tc_apply")(ptr(x), ptr(out), M, F, ptr(img), ptr(None if bias is None else _f32c(bias)), mode, fmt, ptr(err),
                                 stream_ptr()), "nbss_lg_tc_apply")
    return out


def lg_tc_wgrad(du: Tensor, s: Tensor, dW: Tensor, db: Tensor, fmt: int = FMT_F16) -> None:
    """dW [8,F,F] += du^T s, db [8,F] += sum_m du  (du, s: [M,8,F] fp32)."""
    du, s = _f32c(du), _f32c(s)
    M, G8, F = du.shape
    err = device_err_flag(du.device)
    check(_K("nbss_lg_tc_wgrad")(ptr(du), ptr(s), M, F, ptr(dW), ptr(db), fmt, ptr(err), stream_ptr()), "nbss_lg_tc_wgrad")


def full_fwd_tc(x: Tensor, P, pre: str, img: Tensor, out: Optional[Tensor] = None, fmt: int = FMT_F16):
    """full_fwd with the LinearGroup on tensor cores; img = lg_pack(P[pre + 'full.weight'])."""
    x = _f32c(x)
    B, F, T, H = x.shape
    y = torch.empty_like(x) if out is None else out
    s = torch.empty(B, T, 8, F, dtype=torch.float32, device=x.device)
    u = torch.empty_like(s)
    err = device_err_flag(x.device)
    st = _K("nbss_full_fwd_tc")(ptr(x), ptr(y), ptr(s), ptr(u), B, F, T, ptr(_f32c(P[pre + "norm_full.weight"])),
                                ptr(_f32c(P[pre + "norm_full.bias"])), ptr(_f32c(P[pre + "squeeze.0.weight"])),
                                ptr(_f32c(P[pre + "squeeze.0.bias"])), ptr(_f32c(P[pre + "full.bias"])),
                                ptr(_f32c(P[pre + "unsqueeze.0.weight"])), ptr(_f32c(P[pre + "unsqueeze.0.bias"])), ptr(img), fmt,
                                ptr(err), stream_ptr())
    check(st, "nbss_full_fwd_tc")
    return y, s, u


def full_bwd_tc(x: Tensor, dy: Tensor, s: Tensor, u: Tensor, P, pre: str, img: Tensor, G, fmt: int = FMT_F16) -> Tensor:
    x, dy = _f32c(x), _f32c(dy)
    B, F, T, H = x.shape
    dx = torch.empty_like(x)
    ws = torch.empty(2 * s.numel(), dtype=torch.float32, device=x.device)
    err = device_err_flag(x.device)
    st = _K("nbss_full_bwd_tc")(ptr(x), ptr(dy), ptr(dx), ptr(s), ptr(u), ptr(ws), B, F, T, ptr(_f32c(P[pre + "norm_full.weight"])),
                                ptr(_f32c(P[pre + "norm_full.bias"])), ptr(_f32c(P[pre + "squeeze.0.weight"])),
                                ptr(_f32c(P[pre + "squeeze.0.bias"])), ptr(_f32c(P[pre + "unsqueeze.0.weight"])),
                                ptr(_f32c(P[pre + "unsqueeze.0.bias"])), ptr(img), ptr(G[pre + "norm_full.weight"]),
                                ptr(G[pre + "norm_full.bias"]), ptr(G[pre + "squeeze.0.weight"]), ptr(G[pre + "squeeze.0.bias"]),
                                ptr(G[pre + "full.weight"]), ptr(G[pre + "full.bias"]), ptr(G[pre + "unsqueeze.0.weight"]),
                                ptr(G[pre + "unsqueeze.0.bias"]), fmt, ptr(err), stream_ptr())
    check(st, "nbss_full_bwd_tc")
    return dx


# ------------------------------------------------------------------------------------------------ encoder / decoder
def encoder_fwd(x: Tensor, P) -> Tensor:
    x = _f32c(x)
    B, F, T, Cin = x.shape
    y = torch.empty(B, F, T, 96, dtype=torch.float32, device=x.device)
    st = _K("nbss_encoder_fwd")(ptr(x), ptr(y), B * F, T, Cin, ptr(_f32c(P["encoder.weight"])), ptr(_f32c(P["encoder.bias"])), stream_ptr())
    check(st, "nbss_encoder_fwd")
    return y


def encoder_wgrad(x: Tensor, dy: Tensor, G) -> None:
    x, dy = _f32c(x), _f32c(dy)
    B, F, T, Cin = x.shape
    st = _K("nbss_encoder_wgrad")(ptr(x), ptr(dy), B * F, T, Cin, ptr(G["encoder.weight"]), ptr(G["encoder.bias"]), stream_ptr())
    check(st, "nbss_encoder_wgrad")


def decoder_fwd(x: Tensor, P) -> Tensor:
    x = _f32c(x)
    B, F, T, H = x.shape
    cout = P["decoder.weight"].shape[0]
    y = torch.empty(B, F, T, cout, dtype=torch.float32, device=x.device)
    st = _K("nbss_decoder_fwd")(ptr(x), ptr(y), ctypes.c_longlong(B * F * T), cout, ptr(_f32c(P["decoder.weight"])),
                            ptr(_f32c(P["decoder.bias"])), stream_ptr())
    check(st, "nbss_decoder_fwd")
    return y


def decoder_bwd(x: Tensor, dy: Tensor, P, G) -> Tensor:
    x, dy = _f32c(x), _f32c(dy)
    B, F, T, H = x.shape
    cout = P["decoder.weight"].shape[0]
    dx = torch.empty_like(x)
    st = _K("nbss_decoder_bwd")(ptr(x), ptr(dy), ptr(dx), ctypes.c_longlong(B * F * T), cout, ptr(_f32c(P["decoder.weight"])),
                            ptr(G["decoder.weight"]), ptr(G["decoder.bias"]), stream_ptr())
    check(st, "nbss_decoder_bwd")
    return dx


# ------------------------------------------------------------------------------------------------ framing
def _ll(v):
    return ctypes.c_longlong(int(v))


def stft(x: Tensor, n_fft: int, hop: int) -> Tensor:
    """STFT.stft: [B,C,Ts] fp32 -> complex64 [B,C,F,T] (center, reflect, periodic Hann, onesided)."""
    x = _f32c(x)
    B, C, Ts = x.shape
    F, T = n_fft // 2 + 1, 1 + Ts // hop
    out = torch.empty(B, C, F, T, 2, dtype=torch.float32, device=x.device)
    st = _K("nbss_stft")(ptr(x), B, C, Ts, n_fft, hop, 0, 0, ctypes.c_float(0.0), ptr(out), _ll(C * F * T * 2), _ll(F * T * 2),
                     _ll(T * 2), _ll(2), ptr(None), ptr(None), stream_ptr())
    check(st, "nbss_stft")
    return torch.view_as_complex(out)


def stft_norm_pack(x: Tensor, n_fft: int, hop: int, ref_channel: int, eps: float = 1e-6, want_xr: bool = False):
    """Fused stft + Norm(frequency, online) + pack: [B,C,Ts] -> (X [B,F,T,2C], XrMM [B,F,T], Xr [B,F,T] complex|None)."""
    x = _f32c(x)
    B, C, Ts = x.shape
    F, T = n_fft // 2 + 1, 1 + Ts // hop
    out = torch.empty(B, F, T, 2 * C, dtype=torch.float32, device=x.device)
    xrmm = torch.empty(B, F, T, dtype=torch.float32, device=x.device)
    xr = torch.empty(B, F, T, 2, dtype=torch.float32, device=x.device) if want_xr else None
    st = _K("nbss_stft")(ptr(x), B, C, Ts, n_fft, hop, 1, ref_channel, ctypes.c_float(eps), ptr(out), _ll(F * T * 2 * C), _ll(2),
                     _ll(T * 2 * C), _ll(2 * C), ptr(xrmm), ptr(xr), stream_ptr())
    check(st, "nbss_stft")
    return out, xrmm, (torch.view_as_complex(xr) if want_xr else None)


def istft_strided(real_view: Tensor, strides_bsft, scale: Optional[Tensor], B: int, S: int, F: int, T: int, n_fft: int,
                  hop: int, length: int) -> Tensor:
    """iSTFT of a complex tensor given as a float32 storage + (b,s,f,t) strides in floats (imag at +1)."""
    y = torch.empty(B, S, length, dtype=torch.float32, device=real_view.device)
    ib, is_, if_, it = strides_bsft
    st = _K("nbss_istft")(ptr(real_view), _ll(ib), _ll(is_), _ll(if_), _ll(it), ptr(scale), ptr(y), B, S, length, T, n_fft, hop,
                      stream_ptr())
    check(st, "nbss_istft")
    return y


def istft_bwd_strided(dy: Tensor, scale: Optional[Tensor], out: Tensor, strides_bsft, B, S, F, T, n_fft, hop) -> Tensor:
    dy = _f32c(dy)
    ib, is_, if_, it = strides_bsft
    st = _K("nbss_istft_bwd")(ptr(dy), ptr(scale), ptr(out), _ll(ib), _ll(is_), _ll(if_), _ll(it), B, S, dy.shape[-1], T, n_fft, hop,
                          stream_ptr())
    check(st, "nbss_istft_bwd")
    return out


# ------------------------------------------------------------------------------------------------ narrow-band backward
def ffn_bwd(x: Tensor, dy: Tensor, saves, gn_stats: Tensor, P, pre: str, img: Tensor, G, fmt_g: int = FMT_F16):
    """Backward of y = x + tconvffn(x).  saves = [a1, c1, c2, c3, ln_stats] from ffn_fwd(save=True).
    Returns dx; accumulates every tconvffn.* parameter gradient into G (fp32)."""
    x, dy = _f32c(x), _f32c(dy)
    B, F, T, H = x.shape
    n = B * F * T
    a1, c1, c2, c3, ln_stats = saves
    dt = torch.bfloat16 if fmt_g == FMT_BF16 else torch.float16
    gbuf = [torch.empty(n, 192, dtype=dt, device=x.device) for _ in range(4)]  # g_a1, g_c1, g_c2, g_c3
    dx = torch.empty_like(x)
    err = device_err_flag(x.device)
    t = pre + "tconvffn."
    st = _K("nbss_ffn_bwd")(
        ptr(x), ptr(dy), ptr(dx), B * F, T, ptr(_f32c(P[t + "0.weight"])), ptr(_f32c(P[t + "6.weight"])), ptr(_f32c(P[t + "6.bias"])),
        ptr(ln_stats), ptr(gn_stats), ptr(img), ptr(a1), ptr(c1), ptr(c2), ptr(c3), ptr(gbuf[0]), ptr(gbuf[1]), ptr(gbuf[2]),
        ptr(gbuf[3]), ptr(G[t + "0.weight"]), ptr(G[t + "0.bias"]), ptr(G[t + "6.weight"]), ptr(G[t + "6.bias"]), fmt_g, ptr(err),
        stream_ptr())
    check(st, "nbss_ffn_bwd")
    # the weight gradients recompute their activation operands SiLU(.) from the saved pre-activations (kept alive for the side stream)
    with _side_stream((x, dy, gbuf, (a1, c1, c2, c3), gn_stats)):
        st = _K("nbss_ffn_wgrad")(
            ptr(x), ptr(dy), B * F, T, ptr(_f32c(P[t + "0.weight"])), ptr(_f32c(P[t + "0.bias"])), ptr(gbuf[0]), ptr(gbuf[1]),
            ptr(gbuf[2]), ptr(gbuf[3]), ptr(a1), ptr(c1), ptr(c2), ptr(c3), ptr(gn_stats), ptr(_f32c(P[t + "6.weight"])),
            ptr(_f32c(P[t + "6.bias"])), ptr(G[t + "1.weight"]),
            ptr(G[t + "1.bias"]), ptr(G[t + "3.weight"]), ptr(G[t + "3.bias"]), ptr(G[t + "5.weight"]), ptr(G[t + "5.bias"]),
            ptr(G[t + "8.weight"]), ptr(G[t + "8.bias"]), ptr(G[t + "10.weight"]), ptr(G[t + "10.bias"]), fmt_g, fmt_g, ptr(err),
            stream_ptr())
    check(st, "nbss_ffn_wgrad")
    return dx, err


def mhsa_bwd(x: Tensor, dy: Tensor, msave, P, pre: str, img: Tensor, G, fmt_g: int = FMT_F16):
    """Backward of y = x + MHSA(LN(x)).  msave = (qkv, o, lse, ln_stats) from mhsa_fwd(save=True)."""
    x, dy = _f32c(x), _f32c(dy)
    B, F, T, H = x.shape
    n = B * F * T
    qkv, o, lse, ln_stats = msave
    dt = torch.bfloat16 if fmt_g == FMT_BF16 else torch.float16
    dqkv = torch.empty(n, 288, dtype=dt, device=x.device)
    dx = torch.empty_like(x)
    err = device_err_flag(x.device)
    st = _K("nbss_mhsa_bwd")(ptr(x), ptr(dy), ptr(dx), B * F, T, ptr(_f32c(P[pre + "norm_mhsa.weight"])), ptr(ln_stats), ptr(img),
                         ptr(qkv), ptr(o), ptr(lse), ptr(dqkv), ptr(G[pre + "norm_mhsa.weight"]), ptr(G[pre + "norm_mhsa.bias"]),
                         fmt_g, ptr(err), stream_ptr())
    check(st, "nbss_mhsa_bwd")
    with _side_stream((x, dy, dqkv, o)):
        st = _K("nbss_mhsa_wgrad")(ptr(x), ptr(dy), B * F, T, ptr(_f32c(P[pre + "norm_mhsa.weight"])),
                                   ptr(_f32c(P[pre + "norm_mhsa.bias"])), ptr(dqkv), ptr(o), ptr(G[pre + "mhsa.in_proj_weight"]),
                                   ptr(G[pre + "mhsa.in_proj_bias"]), ptr(G[pre + "mhsa.out_proj.weight"]),
                                   ptr(G[pre + "mhsa.out_proj.bias"]), fmt_g, FMT_F16, ptr(err), stream_ptr())
    check(st, "nbss_mhsa_wgrad")
    return dx, err


# ------------------------------------------------------------------------------------------------ F-conv on tensor cores
def fconv_image_bytes() -> int:
    L = _lib.lib()
    L.nbss_fconv_image_bytes.restype = ctypes.c_uint
    return int(L.nbss_fconv_image_bytes())


def fconv_pack(W: Tensor, img: Optional[Tensor] = None, fmt: int = FMT_F16) -> Tensor:
    """UMMA weight images (forward + transposed) of one F-conv weight [96,12,5]."""
    if img is None:
        img = torch.empty(fconv_image_bytes(), dtype=torch.uint8, device=W.device)
    check(_K("nbss_fconv_pack")(ptr(_f32c(W)), ptr(img), fmt, stream_ptr()), "nbss_fconv_pack")
    return img


def fconv_tc_fwd(x: Tensor, P, pre: str, img: Tensor, out: Optional[Tensor] = None, fmt: int = FMT_F16):
    """y = x + PReLU(gconv_F(LN(x))) on tensor cores (fconv_tc.cu); pre = 'layers.i.fconv1' / '...fconv2'."""
    x = _f32c(x)
    B, F, T, H = x.shape
    assert H == 96
    y = torch.empty_like(x) if out is None else out
    err = device_err_flag(x.device)
    st = _K("nbss_fconv_tc_fwd")(ptr(x), ptr(y), B, F, T, ptr(_f32c(P[pre + ".0.weight"])), ptr(_f32c(P[pre + ".0.bias"])),
                                 ptr(_f32c(P[pre + ".1.bias"])), ptr(_f32c(P[pre + ".2.weight"])), ptr(img), fmt, ptr(err), stream_ptr())
    check(st, "nbss_fconv_tc_fwd")
    return y, err


def fconv_tc_bwd(x: Tensor, dy: Tensor, P, pre: str, img: Tensor, G, fmt: int = FMT_F16):
    x, dy = _f32c(x), _f32c(dy)
    B, F, T, H = x.shape
    dx = torch.empty_like(x)
    err = device_err_flag(x.device)
    st = _K("nbss_fconv_tc_bwd")(ptr(x), ptr(dy), ptr(dx), B, F, T, ptr(_f32c(P[pre + ".0.weight"])), ptr(_f32c(P[pre + ".0.bias"])),
                                 ptr(_f32c(P[pre + ".1.bias"])), ptr(_f32c(P[pre + ".2.weight"])), ptr(img), ptr(G[pre + ".1.weight"]),
                                 ptr(G[pre + ".1.bias"]), ptr(G[pre + ".2.weight"]), ptr(G[pre + ".0.weight"]), ptr(G[pre + ".0.bias"]),
                                 fmt, ptr(err), stream_ptr())
    check(st, "nbss_fconv_tc_bwd")
    return dx, err
