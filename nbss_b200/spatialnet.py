"""Drop-in ``SpatialNet`` for the reference's ``models.arch.SpatialNet.SpatialNet`` (models/arch/SpatialNet.py:152-220).

Same constructor signature, same parameter names / shapes / creation order (so ``seed_everything`` gives the same
initial weights and reference checkpoints load), same ``forward(x[B,F,T,dim_input]) -> [B,F,T,dim_output]``
contract.  The torch submodules created here are ONLY parameter containers — their ``forward`` is never called; all
arithmetic runs in the sm_100a kernels of ``libnbss_b200.so`` through ``nbss_b200.ops``.  There is no fallback: without
the library, or on CPU tensors, ``forward`` raises.

Precision policy: fp32 residual stream, norms, activations, cross-band block, encoder/decoder; 16-bit
tensor-core operands (fp16 forward; fp16 gradient operands with automatic power-of-two loss scaling) with fp32
accumulation in the narrow-band block.
"""
from __future__ import annotations

import math
import weakref
from typing import Dict, List, Optional, Tuple

import ctypes

import torch
import torch.nn as nn
from torch import Tensor

from . import ops


import os as _os

_CHECK_EVERY_STEP = _os.environ.get("NBSS_CHECK_EVERY_STEP", "0") == "1"


class LinearGroup(nn.Module):
    """Parameter container for the full-band linear (models/arch/base/linear_group.py:7-37): weight [G,out,in]."""

    def __init__(self, in_features: int, out_features: int, num_groups: int) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.empty(num_groups, out_features, in_features))
        self.bias = nn.Parameter(torch.empty(num_groups, out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.bias, -bound, bound)


class SpatialNetLayer(nn.Module):
    """Parameters of one layer, registered in the reference's order (models/arch/SpatialNet.py:33-74)."""

    def __init__(self, dim_hidden, dim_ffn, dim_squeeze, num_freqs, num_heads, kernel_size, conv_groups, full=None):
        super().__init__()
        fk, tk = kernel_size
        fg, tg = conv_groups
        self.fconv1 = nn.ModuleList([nn.LayerNorm(dim_hidden),
                                     nn.Conv1d(dim_hidden, dim_hidden, fk, groups=fg, padding="same"), nn.PReLU(dim_hidden)])
        self.norm_full = nn.LayerNorm(dim_hidden)
        self.full_share = full is not None
        self.squeeze = nn.Sequential(nn.Conv1d(dim_hidden, dim_squeeze, 1), nn.SiLU())
        self.full = LinearGroup(num_freqs, num_freqs, dim_squeeze) if full is None else full
        self.unsqueeze = nn.Sequential(nn.Conv1d(dim_squeeze, dim_hidden, 1), nn.SiLU())
        self.fconv2 = nn.ModuleList([nn.LayerNorm(dim_hidden),
                                     nn.Conv1d(dim_hidden, dim_hidden, fk, groups=fg, padding="same"), nn.PReLU(dim_hidden)])
        self.norm_mhsa = nn.LayerNorm(dim_hidden)
        self.mhsa = nn.MultiheadAttention(embed_dim=dim_hidden, num_heads=num_heads, batch_first=True)
        self.tconvffn = nn.ModuleList([
            nn.LayerNorm(dim_hidden), nn.Conv1d(dim_hidden, dim_ffn, 1), nn.SiLU(),
            nn.Conv1d(dim_ffn, dim_ffn, tk, padding="same", groups=tg), nn.SiLU(),
            nn.Conv1d(dim_ffn, dim_ffn, tk, padding="same", groups=tg), nn.GroupNorm(tg, dim_ffn), nn.SiLU(),
            nn.Conv1d(dim_ffn, dim_ffn, tk, padding="same", groups=tg), nn.SiLU(), nn.Conv1d(dim_ffn, dim_hidden, 1)])

    def extra_repr(self) -> str:
        return f"full_share={self.full_share}"


# ----------------------------------------------------------------------------------------------------------------------
# functional engine: forward (optionally saving for backward) and backward over a flat dict of parameter tensors
# ----------------------------------------------------------------------------------------------------------------------
class Engine:
    """Runs the network on CUDA tensors given a name -> tensor dict P (reference state_dict keys)."""

    def __init__(self, num_layers: int, fwd_fmt: int = ops.FMT_F16, grad_fmt: int = ops.FMT_F16):
        self.L = num_layers
        self.fwd_fmt = fwd_fmt
        self.grad_fmt = grad_fmt
        self._imgs: Optional[List[Tensor]] = None
        self._img_key = None
        self._fimgs: Optional[List[List[Tensor]]] = None
        self._fimg_key = None
        self._limgs: Optional[List[Tensor]] = None
        self._limg_key = None
        # NBSS_FULL_SIMT=1 keeps the full-band LinearGroup on the fp32 CUDA-core kernels (cross-check of fullband_tc.cu)
        self.full_tc = _os.environ.get("NBSS_FULL_SIMT", "0") != "1"
        # NBSS_WGRAD_STREAM=0 keeps the weight-gradient kernels on the main stream (ops.WGRAD_STREAM)
        self.use_side = _os.environ.get("NBSS_WGRAD_STREAM", "1") != "0"
        self._side: Optional[torch.cuda.Stream] = None

    def invalidate_images(self) -> None:
        """Force a rebuild of every cached weight image at the next forward / backward (for code that updates parameters
        without bumping their version counters, e.g. nbss_b200.optim.FlatClipAdam)."""
        self._img_key = self._fimg_key = self._limg_key = None

    def images(self, P: Dict[str, Tensor]) -> List[Tensor]:
        """Per-layer UMMA weight images; rebuilt whenever a narrow-band weight changed (tensor version counters)."""
        names = []
        for i in range(self.L):
            pre = f"layers.{i}."
            names += [pre + "tconvffn.1.weight", pre + "tconvffn.3.weight", pre + "tconvffn.5.weight", pre + "tconvffn.8.weight",
                      pre + "tconvffn.10.weight", pre + "mhsa.in_proj_weight", pre + "mhsa.out_proj.weight"]
        key = tuple((P[n].data_ptr(), P[n]._version) for n in names)
        # under CUDA-graph capture the pack kernels must be part of the graph (weights change between replays)
        if self._imgs is None or key != self._img_key or torch.cuda.is_current_stream_capturing():
            old = self._imgs
            self._imgs = [ops.pack_layer_weights(P, f"layers.{i}.", old[i] if old else None, self.fwd_fmt, self.grad_fmt)
                          for i in range(self.L)]
            self._img_key = key
        return self._imgs

    def fconv_images(self, P: Dict[str, Tensor]) -> List[List[Tensor]]:
        """Per-layer UMMA images of the two F-conv weights (fconv_tc.cu), rebuilt when a weight changed."""
        names = [f"layers.{i}.fconv{j}.1.weight" for i in range(self.L) for j in (1, 2)]
        key = tuple((P[n].data_ptr(), P[n]._version) for n in names)
        if self._fimgs is None or key != self._fimg_key or torch.cuda.is_current_stream_capturing():
            old = self._fimgs
            self._fimgs = [[ops.fconv_pack(P[f"layers.{i}.fconv{j}.1.weight"], old[i][j - 1] if old else None, self.fwd_fmt)
                            for j in (1, 2)] for i in range(self.L)]
            self._fimg_key = key
        return self._fimgs

    def lg_images(self, P: Dict[str, Tensor]) -> List[Tensor]:
        """Per-layer UMMA images of the full-band LinearGroup weight (fullband_tc.cu); layers that share the module
        (full_share) share the image.  One image serves the forward and the data gradient, so it is packed in fwd_fmt
        and the backward of this sub-block uses fwd_fmt operands too."""
        names = [f"layers.{i}.full.weight" for i in range(self.L)]
        key = tuple((P[n].data_ptr(), P[n]._version) for n in names)
        if self._limgs is None or key != self._limg_key or torch.cuda.is_current_stream_capturing():
            old = self._limgs
            done: Dict[int, Tensor] = {}
            imgs = []
            for i, n in enumerate(names):
                k = P[n].data_ptr()
                if k not in done:
                    done[k] = ops.lg_pack(P[n], old[i] if old else None, self.fwd_fmt)
                imgs.append(done[k])
            self._limgs, self._limg_key = imgs, key
        return self._limgs

    def forward(self, P: Dict[str, Tensor], x: Tensor, save: bool):
        if not x.is_cuda:
            raise ops._lib.NbssError("nbss_b200.SpatialNet runs on CUDA tensors only (there is no CPU path)")
        imgs = self.images(P)
        fimgs = self.fconv_images(P)
        # the tensor-core LinearGroup holds one F x F weight per CTA with N = F <= 256 (UMMA limit); wider bands (16 kHz: F = 257,
        # models/io/stft.py:8-12) run the full-band sub-block on the fp32 CUDA-core kernels (crossband.cu), 7 % of a layer
        full_tc = self.full_tc and x.shape[1] <= 256
        limgs = self.lg_images(P) if full_tc else None
        ctx = {"x_in": x, "layers": [], "full_tc": full_tc} if save else None
        errs = []
        h = ops.encoder_fwd(x, P)
        for i in range(self.L):
            pre = f"layers.{i}."
            if save:
                lc = {"x0": h}
                h1, e3 = ops.fconv_tc_fwd(h, P, pre + "fconv1", fimgs[i][0], fmt=self.fwd_fmt)
                h2, s, u = (ops.full_fwd_tc(h1, P, pre, limgs[i], fmt=self.fwd_fmt) if full_tc else ops.full_fwd(h1, P, pre))
                h3, e4 = ops.fconv_tc_fwd(h2, P, pre + "fconv2", fimgs[i][1], fmt=self.fwd_fmt)
                h4, msave, e1 = ops.mhsa_fwd(h3, P, pre, imgs[i], save=True, fmt=self.fwd_fmt)
                h5, fsave, gstats, e2 = ops.ffn_fwd(h4, P, pre, imgs[i], save=True, fmt=self.fwd_fmt)
                lc.update(x1=h1, s=s, u=u, x2=h2, x3=h3, msave=msave, x4=h4, fsave=fsave, gstats=gstats)
                ctx["layers"].append(lc)
                h = h5
            else:
                # inference: every sub-block updates the stream in place (each kernel reads a row before writing it)
                h, e3 = ops.fconv_tc_fwd(h, P, pre + "fconv1", fimgs[i][0], out=h, fmt=self.fwd_fmt)
                h, _, _ = (ops.full_fwd_tc(h, P, pre, limgs[i], out=h, fmt=self.fwd_fmt) if full_tc else ops.full_fwd(h, P, pre, out=h))
                h, e4 = ops.fconv_tc_fwd(h, P, pre + "fconv2", fimgs[i][1], out=h, fmt=self.fwd_fmt)
                h, e1 = ops.mhsa_fwd(h, P, pre, imgs[i], fmt=self.fwd_fmt, out=h)
                h, e2 = ops.ffn_fwd(h, P, pre, imgs[i], fmt=self.fwd_fmt, out=h)
            errs += [e1, e2, e3, e4]
        y = ops.decoder_fwd(h, P)
        if save:
            ctx["x_last"] = h
        return y, ctx, errs

    def backward(self, P: Dict[str, Tensor], ctx, dy: Tensor, G: Dict[str, Tensor]):
        """Accumulates parameter gradients into G (fp32 tensors keyed like P).  The network input needs no gradient
        (SharedTrainer.py:113-120: X comes from the STFT of the data)."""
        imgs = self.images(P)
        fimgs = self.fconv_images(P)
        limgs = self.lg_images(P) if ctx["full_tc"] else None
        errs = []
        if self.use_side and self._side is None:
            self._side = torch.cuda.Stream(device=dy.device)
        ops.WGRAD_STREAM = self._side if self.use_side else None
        try:
            d = self._backward_layers(P, ctx, dy, G, imgs, fimgs, limgs, errs)
        finally:
            if ops.WGRAD_STREAM is not None:
                torch.cuda.current_stream().wait_stream(ops.WGRAD_STREAM)
            ops.SIDE_KEEP.clear()
            ops.WGRAD_STREAM = None
        ops.encoder_wgrad(ctx["x_in"], d, G)
        return errs

    def _backward_layers(self, P, ctx, dy, G, imgs, fimgs, limgs, errs):
        d = ops.decoder_bwd(ctx["x_last"], dy, P, G)
        for i in reversed(range(self.L)):
            pre = f"layers.{i}."
            lc = ctx["layers"][i]
            d, e1 = ops.ffn_bwd(lc["x4"], d, lc["fsave"], lc["gstats"], P, pre, imgs[i], G, fmt_g=self.grad_fmt)
            d, e2 = ops.mhsa_bwd(lc["x3"], d, lc["msave"], P, pre, imgs[i], G, fmt_g=self.grad_fmt)
            d, e3 = ops.fconv_tc_bwd(lc["x2"], d, P, pre + "fconv2", fimgs[i][1], G, fmt=self.grad_fmt)
            d = (ops.full_bwd_tc(lc["x1"], d, lc["s"], lc["u"], P, pre, limgs[i], G, fmt=self.fwd_fmt) if ctx["full_tc"]
                 else ops.full_bwd(lc["x1"], d, lc["s"], lc["u"], P, pre, G))
            d, e4 = ops.fconv_tc_bwd(lc["x0"], d, P, pre + "fconv1", fimgs[i][0], G, fmt=self.grad_fmt)
            errs += [e1, e2, e3, e4]
            if ops.WGRAD_STREAM is not None:  # join before this layer's activations and gradient operands are freed
                torch.cuda.current_stream().wait_stream(ops.WGRAD_STREAM)
                ops.SIDE_KEEP.clear()
            ctx["layers"][i] = None  # free this layer's saved activations
        return d


class _SpatialNetFn(torch.autograd.Function):
    @staticmethod
    def forward(fctx, module: "SpatialNet", x: Tensor, *params: Tensor):
        P = module._param_dict()
        module._poll_device_errors()
        y, ctx, errs = module.engine.forward(P, x.detach().float(), save=True)
        fctx.module, fctx.ctx, fctx.errs = module, ctx, errs
        module._live_graphs.add(fctx)  # weak: a graph that is dropped without a backward disappears by itself
        return y

    @staticmethod
    def backward(fctx, dy: Tensor):
        module = fctx.module
        if fctx.ctx is None:
            raise ops._lib.NbssError("nbss_b200.SpatialNet: backward through the same graph twice is not supported (the saved "
                                     "activations are freed layer by layer during the first backward); re-run the forward")
        P = module._param_dict()
        # another forward of this module is still waiting for its backward (two forwards, one loss.backward()): its views
        # of the persistent buffer may already sit in the autograd engine's input buffers, so this node gets a fresh buffer
        fresh = any(c is not fctx and getattr(c, "ctx", None) is not None for c in module._live_graphs)
        flat, G, views = module.make_flat_grads(dy.device, fresh=fresh)
        # fp16 gradient operands: the backward is linear in dy, so scale dy by a power of two that brings its largest
        # element to ~1 and undo it on the flat gradient buffer (exact loss scaling, no host synchronisation)
        dy = dy.contiguous().float()
        dys = torch.empty_like(dy)
        ws = torch.empty(4, dtype=torch.float32, device=dy.device)  # [0]: bits of max |dy|, [2:4]: scale, 1 / scale
        ops.check(ops._K("nbss_grad_prescale")(ops.ptr(dy), ctypes.c_longlong(dy.numel()), ops.ptr(dys), ops.ptr(ws), ops.ptr(ws[2:]),
                                               ops.stream_ptr()), "nbss_grad_prescale")
        errs = module.engine.backward(P, fctx.ctx, dys, G)
        ops.check(ops._K("nbss_grad_unscale")(ops.ptr(flat), ctypes.c_longlong(flat.numel()), ops.ptr(ws[2:]), ops.stream_ptr()), "nbss_grad_unscale")
        module._last_flat_grad = flat  # one contiguous buffer: a single NCCL all-reduce covers every gradient
        # The device error flag is sticky and shared by all launches; reading it needs a host sync, so the hot path does
        # it only on request (NBSS_CHECK_EVERY_STEP=1) — `SpatialNet.check_device_errors()` reads it at any time.
        module._last_errs = fctx.errs + errs
        if _CHECK_EVERY_STEP:
            module.check_device_errors()
        else:
            module._post_device_error_check(dy.device)
        fctx.ctx = None
        return (None, None) + tuple(views)


class SpatialNet(nn.Module):
    """See module docstring.  Tensor-core path supports the reference's small configuration (dim_hidden 96, dim_ffn
    192, 4 heads, conv_groups (8,8), kernel_size (5,3), LN/GN norms, zero padding, no dropout).  Training: T <= 256 frames (the 4 s
    crops of the reference's configs); inference (torch.no_grad()) takes any T up to 65536 (ops.mhsa_fwd / ops.ffn_fwd switch to the
    chunked long-sequence kernels above 256 frames)."""

    def __init__(self, dim_input: int, dim_output: int, dim_squeeze: int, num_layers: int, num_freqs: int,
                 encoder_kernel_size: int = 5, dim_hidden: int = 192, dim_ffn: int = 384, num_heads: int = 2,
                 dropout: Tuple[float, float, float] = (0, 0, 0), kernel_size: Tuple[int, int] = (5, 3),
                 conv_groups: Tuple[int, int] = (8, 8), norms: List[str] = ("LN", "LN", "GN", "LN", "LN", "LN"),
                 padding: str = "zeros", full_share: int = 0):
        super().__init__()
        unsupported = []
        if (dim_hidden, dim_ffn, num_heads) != (96, 192, 4): unsupported.append("dim_hidden/dim_ffn/num_heads != 96/192/4")
        if tuple(kernel_size) != (5, 3) or tuple(conv_groups) != (8, 8) or encoder_kernel_size != 5: unsupported.append("kernel sizes / groups")
        if [n.upper() for n in norms] != ["LN", "LN", "GN", "LN", "LN", "LN"]: unsupported.append("norms")
        if any(d > 0 for d in dropout): unsupported.append("dropout > 0")
        if padding != "zeros": unsupported.append("padding")
        if dim_squeeze != 8: unsupported.append("dim_squeeze != 8")
        if unsupported:
            raise NotImplementedError("nbss_b200.SpatialNet supports the SpatialNet-small layer widths only (DESIGN.md §1; the large model is "
                                      "SURVEY §8f rank-4 work that is not built): " + ", ".join(unsupported))
        self.encoder = nn.Conv1d(dim_input, dim_hidden, encoder_kernel_size, stride=1, padding="same")
        full, layers = None, []
        for l in range(num_layers):
            layer = SpatialNetLayer(dim_hidden, dim_ffn, dim_squeeze, num_freqs, num_heads, kernel_size, conv_groups,
                                    full=full if l > full_share else None)
            full = layer.full
            layers.append(layer)
        self.layers = nn.ModuleList(layers)
        self.decoder = nn.Linear(dim_hidden, dim_output)
        self.engine = Engine(num_layers)
        self._live_graphs = weakref.WeakSet()  # autograd contexts of forwards whose backward has not run yet
        self._aliases: Dict[str, List[str]] = {}
        self._build_aliases()

    def _build_aliases(self) -> None:
        """state_dict repeats the shared ``full.*`` tensor under every layer; map unique parameter -> all its keys."""
        first: Dict[int, str] = {}
        self._aliases = {}
        for name, p in self.named_parameters(remove_duplicate=False):
            if id(p) not in first:
                first[id(p)] = name
                self._aliases[name] = []
            self._aliases[first[id(p)]].append(name)

    def make_flat_grads(self, device, fresh: bool = False):
        """One contiguous zero fp32 buffer holding every parameter gradient; returns (flat, name->view dict covering all
        state-dict aliases of shared tensors, list of views in parameter registration order).

        The buffer is persistent (same storage every step, which is what CUDA-graph replays and a standing NCCL
        registration want) unless some parameter's .grad still lives in it (gradient accumulation without zero_grad):
        then a fresh buffer is used and autograd adds it to the existing gradients."""
        uniq = self._unique_params()
        n = sum(p.numel() for _, p in uniq)
        flat = getattr(self, "_flat_grad_buf", None)
        if fresh:
            flat = torch.zeros(n, dtype=torch.float32, device=device)
        elif flat is None or flat.device != torch.device(device) or flat.numel() != n:
            flat = self._flat_grad_buf = torch.zeros(n, dtype=torch.float32, device=device)
        else:
            base = flat.untyped_storage().data_ptr()
            if any(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for _, p in uniq):
                flat = torch.zeros(n, dtype=torch.float32, device=device)
            else:
                flat.zero_()
        G, off, views = {}, 0, []
        for name, p in uniq:
            v = flat[off:off + p.numel()].view_as(p)
            views.append(v)
            off += p.numel()
            for alias in self._aliases[name]:
                G[alias] = v
        return flat, G, views

    def grads_alias_flat(self) -> bool:
        """True when every parameter's .grad is a view of the last flat gradient buffer (autograd adopted the views
        returned by the backward instead of cloning them), i.e. one all-reduce of that buffer reduces every gradient."""
        flat = getattr(self, "_last_flat_grad", None)
        if flat is None:
            return False
        base = flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.parameters())

    def _unique_params(self):
        return list(self.named_parameters())  # duplicates removed, registration order

    def _param_dict(self) -> Dict[str, Tensor]:
        # p.detach() shares the parameter's version counter (p.data would start a fresh one): an in-place update by any
        # torch optimizer, load_state_dict() or p.copy_() changes the keys of the cached weight images (Engine.images)
        return {n: p.detach() for n, p in self.named_parameters(remove_duplicate=False)}

    def forward(self, x: Tensor, return_attn_score: bool = False):
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        with ops._lib.device_of(x):  # launches go to the current stream of x's device (autograd sets it for the backward)
            if needs_grad:
                y = _SpatialNetFn.apply(self, x, *[p for _, p in self._unique_params()])
            else:
                self._poll_device_errors()
                y, _, errs = self.engine.forward(self._param_dict(), x.detach().float().contiguous(), save=False)
                self._last_errs = errs
                self._post_device_error_check(x.device)
        if return_attn_score:
            return y, [None] * len(self.layers)  # the reference also returns None here (SpatialNet.py:97 quirk)
        return y

    def _post_device_error_check(self, device) -> None:
        """Queues an asynchronous copy of the device error flag into pinned host memory (no host sync); the value is looked
        at by the NEXT forward (`_poll_device_errors`), i.e. one step late, which keeps a timed-out barrier from silently
        corrupting a whole run.  Skipped under CUDA-graph capture (replays cannot raise)."""
        if torch.cuda.is_current_stream_capturing():
            return
        host = getattr(self, "_err_host", None)
        if host is None:
            host = self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._err_event = torch.cuda.Event()
        host.copy_(ops.device_err_flag(torch.device(device)), non_blocking=True)
        self._err_event.record()
        self._err_posted = True

    def _poll_device_errors(self) -> None:
        if not getattr(self, "_err_posted", False) or torch.cuda.is_current_stream_capturing():
            return
        if not self._err_event.query():  # the copy has not finished yet: look again at the next call
            return
        self._err_posted = False
        v = int(self._err_host.item())
        if v != 0:
            for f in list(ops._ERR_FLAGS.values()):
                f.zero_()  # reset, so the caller can recover (e.g. re-run the step)
            raise ops._lib.NbssError(f"nbss_b200 kernel: device-side error flag {v:#x} was raised during the previous step (an mbarrier "
                                     "wait timed out: GPU time-slicing, preemption or a debugger); that step's results are invalid")

    def check_device_errors(self) -> None:
        """Raises if any tensor-core kernel launched so far reported an mbarrier time-out (host sync)."""
        seen = set()
        for e in getattr(self, "_last_errs", []):
            if id(e) not in seen:
                seen.add(id(e))
                ops.check_err_flag(e, "nbss_b200 kernel")
