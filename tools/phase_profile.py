#!/usr/bin/env python
"""Per-phase timing of the persistent slab kernels (profiling build: `make -C nbss_b200/csrc prof`).

    NBSS_LIB=nbss_b200/lib/libnbss_b200_prof.so python tools/phase_profile.py [--batch 8] [--out profiles/r02_phases.json]

Thread 0 of CTA 0 stamps clock64() at every phase boundary of its SECOND work item (common.cuh: NBSS_TICK); this script
runs each kernel alone at the bench shape (F=129, T=250), reads the stamps back and prints the cycles spent in each
phase — where a slab's time goes (staging / MMA wait / epilogue), which neither ncu's whole-kernel counters nor the
CUDA-event timing of a launch can show.  Numbers are cycles of ONE CTA's second slab at whatever clock the GPU runs.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NBSS_LIB", os.path.join(ROOT, "nbss_b200", "lib", "libnbss_b200_prof.so"))

import torch  # noqa: E402

from nbss_b200 import _lib, ops  # noqa: E402
from oracle import spatialnet_oracle as O  # noqa: E402

NAMES = {
    "ffn_fwd": {0: ["stage LN(x)", "pw1 MMA wait", "E1 SiLU(a1)", "conv1 MMA wait", "E2 SiLU(c1)", "conv2 MMA wait", "E3 GroupNorm+SiLU",
                    "conv3 MMA wait", "E4 SiLU(c3)", "pw2 MMA wait", "E5a D->smem", "E5b residual out"]},
    "ffn_bwd": {0: ["stage dy", "B1 MMA wait", "E1 SiLU'(c3)", "B2 MMA wait", "E2 GN bwd", "B3 MMA wait", "E3 SiLU'(c1)", "B4 MMA wait",
                    "E4 SiLU'(a1)", "B5 MMA wait", "E5a D->smem", "E5b LN bwd + out"]},
    "mhsa_fwd": {0: ["stage LN(x)", "KV MMA wait", "E1 K|V out", "Q MMA wait", "(heads)", "out-proj D->smem", "residual out"]},
    "mhsa_bwd": {0: ["stage dy", "dO MMA wait", "E0 dO, delta"], 1: ["dQKV load + MMA wait", "D->smem", "LN bwd + out"]},
    "fconv_tc": {0: ["stage LN(x)", "conv MMA wait", "E1 PReLU", "E2 residual out"],
                 1: ["stage LN(x) (warp 0)", "stage dy + sync", "conv MMA wait", "E-A dc", "wgrad MMA wait", "wgrad read-out", "dgrad MMA wait", "E-B1 dh", "E-B2 LN bwd + out"]},
}


def read(unit):
    buf = (ctypes.c_ulonglong * 256)()
    fn = getattr(_lib.lib(), "nbss_debug_phases_" + unit)
    torch.cuda.synchronize()
    assert fn(buf) == 0
    return [list(buf[64 * k:64 * k + 64]) for k in range(4)]


def deltas(st):
    idx = [i for i, v in enumerate(st) if v]
    return idx, [st[b] - st[a] for a, b in zip(idx[:-1], idx[1:])]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
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
    out = {}

    def report(unit, st):
        out[unit] = {}
        for kid, names in NAMES[unit].items():
            idx, d = deltas(st[kid])
            if not d:
                continue
            tot = st[kid][idx[-1]] - st[kid][idx[0]]
            print(f"--- {unit} kernel {kid}: {tot} cycles per work item")
            rows = []
            if unit == "mhsa_fwd":
                s = st[kid]
                S = lambda h, b: s[9 + 10 * h + 2 * b]   # after softmax (+ PV issue) of query tile b of head h
                R = lambda h, b: s[13 + 10 * h + 2 * b]  # after its read-out (+ next S issue)
                rows = [(names[0], s[1] - s[0]), (names[1], s[2] - s[1]), (names[2], s[3] - s[2]), (names[3], s[4] - s[3]), ("EQ all heads", s[8] - s[4])]
                sm = sum((S(h, 0) - (s[8] if h == 0 else R(h - 1, 1))) + (S(h, 1) - S(h, 0)) for h in range(4))
                ro = sum((R(h, 0) - S(h, 1)) + (R(h, 1) - R(h, 0)) for h in range(4))
                rows += [("8x S wait + softmax + P->TMEM", sm), ("8x PV wait + read-out", ro)]
                rows += [("out-proj load + MMA wait", s[5] - R(3, 1)), (names[5], s[6] - s[5]), (names[6], s[7] - s[6])]
            elif unit == "mhsa_bwd" and kid == 0:
                s = st[kid]
                rows = [(names[0], s[1] - s[0]), (names[1], s[2] - s[1]), (names[2], s[3] - s[2])]
                ld = sum(s[4 + 12 * h] - (s[3] if h == 0 else s[14 + 12 * (h - 1)]) for h in range(4))
                wt = sum(s[5 + 12 * h + 2 * b] - (s[4 + 12 * h] if b == 0 else s[6 + 12 * h + 2 * (b - 1)]) for h in range(4) for b in range(4))
                ep = sum(s[6 + 12 * h + 2 * b] - s[5 + 12 * h + 2 * b] for h in range(4) for b in range(4))
                lw = sum(s[13 + 12 * h] - s[12 + 12 * h] for h in range(4))
                ro = sum(s[14 + 12 * h] - s[13 + 12 * h] for h in range(4))
                rows += [("4x q,k,v TMA load", ld), ("16x S/dP(+grad) MMA wait", wt), ("16x P, dS epilogue", ep), ("4x last grad MMA wait", lw), ("4x dQ,dK,dV read-out", ro)]
            else:
                rows = [(names[i] if i < len(names) else f"phase {i}", v) for i, v in zip(idx, d)]
            for n, v in rows:
                print(f"    {n:28s} {v:8d}  {100.0 * v / tot:5.1f} %")
            out[unit][str(kid)] = {"cycles_per_item": tot, "phases": rows}

    for _ in range(2):
        y, fsave, gst, _ = ops.ffn_fwd(x, P, pre, img, save=True)
    report("ffn_fwd", read("ffn_fwd"))
    for _ in range(2):
        ops.ffn_bwd(x, dy, fsave, gst, P, pre, img, G)
    report("ffn_bwd", read("ffn_bwd"))
    for _ in range(2):
        y, msave, _ = ops.mhsa_fwd(x, P, pre, img, save=True)
    report("mhsa_fwd", read("mhsa_fwd"))
    for _ in range(2):
        ops.mhsa_bwd(x, dy, msave, P, pre, img, G)
    report("mhsa_bwd", read("mhsa_bwd"))
    for _ in range(2):
        ops.fconv_tc_fwd(x, P, pre + "fconv1", fimg)
        ops.fconv_tc_bwd(x, dy, P, pre + "fconv1", fimg, G)
    report("fconv_tc", read("fconv_tc"))
    ops.check_err_flag(ops.device_err_flag(x.device), "phase profile")
    if a.out:
        json.dump({"batch": B, "F": F, "T": T, "unit": "SM cycles of CTA 0's second work item", "kernels": out}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
