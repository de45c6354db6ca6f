"""clip_grad_norm_ + Adam over the network's flat gradient buffer in two launches (csrc/loss.cu: nbss_clip_adam).

Mirrors what the reference's trainer does after backward: Lightning ``gradient_clip_val: 5`` (configs/SpatialNet.yaml:4) then
``torch.optim.Adam(lr=1e-3)`` (configs/SpatialNet.yaml:44, models/utils/general_steps.py:243-271) — same update rule
(no weight decay, no amsgrad), same clipping formula (``max_norm / (total_norm + 1e-6)``, clamped to 1).
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import torch

from . import _lib, ops


class FlatClipAdam:
    def __init__(self, module, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, max_norm: float = 5.0):
        self.module = module
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        uniq = [p for _, p in module._unique_params()]
        dev = uniq[0].device
        if dev.type != "cuda":
            raise _lib.NbssError("FlatClipAdam runs on CUDA parameters only (there is no CPU path)")
        for p in uniq:
            assert p.dtype == torch.float32 and p.is_contiguous()
        self._params = uniq
        sizes = [p.numel() for p in uniq]
        self.n = sum(sizes)
        self.ptrs = torch.tensor([p.data_ptr() for p in uniq], dtype=torch.int64, device=dev)
        self.offsets = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64, device=dev)
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self.module.zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self) -> None:
        """Uses the flat gradient buffer of the last backward (every p.grad is a view of it)."""
        flat = getattr(self.module, "_last_flat_grad", None)
        if flat is None or flat.numel() != self.n:
            raise _lib.NbssError("FlatClipAdam.step(): no flat gradient buffer (run a backward through nbss_b200.SpatialNet first)")
        if not self.module.grads_alias_flat():
            # gradient accumulation without zero_grad, a DDP bucket copy-back, user hooks ...: the p.grad tensors are the
            # truth, so gather them into the flat buffer (one small copy per tensor) instead of stepping on a stale buffer
            if any(p.grad is None for p in self._params):
                raise _lib.NbssError("FlatClipAdam.step(): some parameters have no gradient")
            off = 0
            for p in self._params:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                off += p.numel()
        b1, b2 = self.betas
        _lib.check(ops._K("nbss_clip_adam")(_lib.ptr(self.ptrs), _lib.ptr(self.offsets), len(self._params), ctypes.c_longlong(self.n),
                                            _lib.ptr(flat), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), _lib.ptr(self.gnorm_sq),
                                            _lib.ptr(self.step_count), ctypes.c_float(self.max_norm), ctypes.c_float(self.lr),
                                            ctypes.c_float(b1), ctypes.c_float(b2), ctypes.c_float(self.eps), _lib.stream_ptr()),
                   "nbss_clip_adam")
        # the parameters were updated through raw pointers: their tensor version counters did not move, so the cached
        # UMMA weight images must be rebuilt explicitly
        self.module.engine.invalidate_images()

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm before clipping of the last step (device scalar), what clip_grad_norm_ returns."""
        return self.gnorm_sq.sqrt().float()
