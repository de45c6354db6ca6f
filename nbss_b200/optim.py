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

    # ------------------------------------------------------------------------------------------------ checkpoint / resume
    # The reference's trainer saves and restores ``torch.optim.Adam``'s state inside the Lightning checkpoint
    # (``optimizer_states``; models/utils/general_steps.py:243-271 builds the optimizer over ``self.parameters()``, i.e. the
    # arch's parameters in registration order).  state_dict() / load_state_dict() speak that format, so a run can move between
    # the reference's optimizer and this one in either direction.
    def state_dict(self) -> dict:
        state, off = {}, 0
        step = self.step_count.detach().cpu().reshape(())  # torch.optim.Adam keeps `step` as a CPU scalar (non-capturable default)
        for i, p in enumerate(self._params):
            n = p.numel()
            state[i] = {"step": step.clone(),
                        "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
            off += n
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(self._params)))}
        return {"state": state, "param_groups": [group], "max_norm": self.max_norm}

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self._params):
            raise ValueError(f"FlatClipAdam.load_state_dict: expected one param group over {len(self._params)} tensors "
                             f"(the arch's parameters in registration order), got {[len(g['params']) for g in groups]}")
        g = groups[0]
        if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("FlatClipAdam implements plain Adam (no weight decay / amsgrad / maximize)")
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(float(b) for b in g["betas"]), float(g["eps"])
        if "max_norm" in sd:
            self.max_norm = float(sd["max_norm"])
        state = sd.get("state", {})
        if not state:  # an optimizer that has not stepped yet
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            self.step_count.zero_()
            return
        steps, off = set(), 0
        for i, (idx, p) in enumerate(zip(g["params"], self._params)):
            st = state.get(idx, state.get(str(idx)))
            n = p.numel()
            if st is None or tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"FlatClipAdam.load_state_dict: state of parameter {i} is missing or has the wrong shape")
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(float(st["step"]))
            off += n
        if len(steps) != 1:
            raise ValueError(f"FlatClipAdam.load_state_dict: the parameters carry different step counts {sorted(steps)}")
        self.step_count.fill_(steps.pop())

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm before clipping of the last step (device scalar), what clip_grad_norm_ returns."""
        return self.gnorm_sq.sqrt().float()
