"""Test / predict-time callers around the hot path:

* ``recover_scale`` + peak normalisation — ``models/utils/metrics.py:192-218`` and ``SharedTrainer.py:293-305`` — on the GPU
  (csrc/predict.cu, three launches);
* ``predict_step`` — the reference's ``TrainModule.predict_step`` (``SharedTrainer.py:277-307``) for a scale-invariant loss:
  forward, scale recovery against the reference-channel mixture, optional PIT re-ordering against given targets, peak
  normalisation;
* ``ensemble_state_dict`` / ``load_reference_checkpoint`` — checkpoint averaging and key handling of
  ``models/utils/ensemble.py:40-52`` and ``models/utils/general_steps.py:163-214`` (``arch.`` prefix of the LightningModule,
  ``_orig_mod.`` of compiled checkpoints), so released reference checkpoints load into the drop-in modules.
"""
from __future__ import annotations

import ctypes
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Union

import torch
from torch import Tensor

from . import _lib, ops


@torch.no_grad()
def recover_scale(preds: Tensor, mixture: Tensor, scale_src_together: bool = False, norm_if_exceed_1: bool = True,
                  return_scales: bool = False):
    """preds [B,S,Ts], mixture [B,Ts] -> scale-recovered preds (models/utils/metrics.py:192-218)."""
    if scale_src_together:
        raise NotImplementedError("scale_src_together=True (neg_sa_sdr) is not on the SpatialNet path")
    if not preds.is_cuda:
        raise _lib.NbssError("nbss_b200.predict runs on CUDA tensors only (there is no CPU path)")
    p, m = ops._f32c(preds), ops._f32c(mixture)
    B, S, Ts = p.shape
    out = torch.empty_like(p)
    ws_sums = torch.empty(B * 14, dtype=torch.float64, device=p.device)
    ws_peak = torch.empty(B * S, dtype=torch.int32, device=p.device)
    scales = torch.empty(B, S, dtype=torch.float32, device=p.device)
    _lib.check(ops._K("nbss_predict_post")(_lib.ptr(p), _lib.ptr(m), _lib.ptr(out), B, S, ctypes.c_longlong(Ts), 1, int(norm_if_exceed_1),
                                           _lib.ptr(ws_sums), _lib.ptr(ws_peak), _lib.ptr(scales), _lib.stream_ptr()), "nbss_predict_post")
    return (out, scales) if return_scales else out


@torch.no_grad()
def predict_step(pipeline, x: Tensor, yr: Optional[Tensor] = None, ref_channel: int = 0, norm_if_exceed_1: bool = True) -> Tensor:
    """TrainModule.predict_step (SharedTrainer.py:277-307) for is_scale_invariant_loss: x [B,C,Ts] -> ys_hat [B,S,Ts].
    `pipeline`: nbss_b200.io.SeparationPipeline (or any callable wave -> estimates)."""
    from .loss import neg_si_sdr_pit

    yr_hat = pipeline(x)
    yr_hat = recover_scale(yr_hat, x[:, ref_channel, :], scale_src_together=False, norm_if_exceed_1=False)
    if yr is not None:  # pit(metric=si_sdr, eval_func='max') == the permutation that minimises the negative SI-SDR
        perms = neg_si_sdr_pit(yr_hat, yr)[2].long()
        yr_hat = torch.gather(yr_hat, 1, perms[:, :, None].expand_as(yr_hat))
    if norm_if_exceed_1:
        mx = yr_hat.abs().amax(dim=-1, keepdim=True)
        yr_hat = yr_hat / torch.where(mx > 1, mx, torch.ones_like(mx))
    return yr_hat


def ensemble_state_dict(ckpts: Iterable[Union[str, Path]]) -> Dict[str, Tensor]:
    """Average of the 'state_dict' entries of the given Lightning checkpoints (models/utils/ensemble.py:40-52)."""
    paths: List[str] = sorted({Path(c).name: str(c) for c in ckpts}.values())
    if not paths:
        raise ValueError("no checkpoints to ensemble")
    out: Dict[str, Tensor] = {}
    for path in paths:
        data = torch.load(path, map_location="cpu", weights_only=False)
        for k, v in data["state_dict"].items():
            out[k] = out[k] + v / len(paths) if k in out else v / len(paths)
    return out


def load_reference_checkpoint(arch: torch.nn.Module, ckpt: Union[str, Path, Dict[str, Tensor]], prefix: str = "arch.",
                              ensemble: Optional[Iterable[Union[str, Path]]] = None, strict: bool = True):
    """Loads a reference (Lightning) checkpoint — a path, or an already loaded state dict — into a drop-in arch module: keeps
    the keys under `prefix` (the LightningModule attribute name, SharedTrainer.py:66), strips it and the `_orig_mod.` that
    torch.compile'd checkpoints carry (general_steps.py:189-199); `ensemble` = further checkpoints to average with."""
    if isinstance(ckpt, dict):
        sd = ckpt.get("state_dict", ckpt)
    elif ensemble:
        sd = ensemble_state_dict(list(ensemble) + [ckpt])
    else:
        sd = torch.load(ckpt, map_location="cpu", weights_only=False)["state_dict"]
    clean = {}
    for k, v in sd.items():
        k2 = k.replace("_orig_mod.", "")
        if prefix and not k2.startswith(prefix):
            continue  # stft.window, loss buffers ... belong to other attributes of the LightningModule
        clean[k2[len(prefix):]] = v
    return arch.load_state_dict(clean, strict=strict)
