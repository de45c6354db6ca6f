#!/usr/bin/env python
"""CUDA-event timing of every sub-block kernel alone at the bench shape (batch 32 by default), for whichever build NBSS_LIB
selects — the quick A/B instrument between full bench runs:
    for L in "" _nopf; do NBSS_LIB=nbss_b200/lib/libnbss_b200$L.so python tools/time_kernels.py; done
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nbss_b200 import ops  # noqa: E402
from oracle import spatialnet_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default=None)
a = ap.parse_args()
B, F, T = a.batch, 129, 250
P = {k: v.cuda() for k, v in O.synth_params(O.SMALL_CFG, 5).items()}
pre = "layers.1."
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, F, T, 96, generator=g, device="cuda")
dy = torch.randn(B, F, T, 96, generator=g, device="cuda")
G = {k: torch.zeros_like(v) for k, v in P.items()}
img = ops.pack_layer_weights(P, pre)
fimg = ops.fconv_pack(P[pre + "fconv1.1.weight"])
limg = ops.lg_pack(P[pre + "full.weight"])
ops.TIMING = {}
for _ in range(a.reps + 1):
    y, fsave, gst, _ = ops.ffn_fwd(x, P, pre, img, save=True)
    ops.ffn_bwd(x, dy, fsave, gst, P, pre, img, G)
    y, msave, _ = ops.mhsa_fwd(x, P, pre, img, save=True)
    ops.mhsa_bwd(x, dy, msave, P, pre, img, G)
    ops.fconv_tc_fwd(x, P, pre + "fconv1", fimg)
    ops.fconv_tc_bwd(x, dy, P, pre + "fconv1", fimg, G)
    y, s, u = ops.full_fwd_tc(x, P, pre, limg)
    ops.full_bwd_tc(x, dy, s, u, P, pre, limg, G)
torch.cuda.synchronize()
ops.check_err_flag(ops.device_err_flag(x.device), "time_kernels")
res = {k: round(min(e0.elapsed_time(e1) for e0, e1 in v[1:]), 4) for k, v in ops.TIMING.items()}
print(os.environ.get("NBSS_LIB", "default lib"), json.dumps(res))
if a.out:
    json.dump({"lib": os.environ.get("NBSS_LIB", "default"), "batch": B, "ms_per_launch_best": res}, open(a.out, "w"), indent=1)
