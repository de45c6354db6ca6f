"""GPU: pins the tcgen05 descriptor / smem-layout conventions of nbss_b200/csrc/umma.cuh against torch matmul.

Every tensor-core kernel in the library builds its operands with exactly these conventions (chunk-column smem tile,
K-major / MN-major views, row-shifted views for conv taps), so this is the first thing to go green on a B200.
"""
import ctypes

import pytest
import torch

from nbss_b200 import _lib

FMT = {"f16": 0, "bf16": 1, "tf32": 2}


def _round(x, fmt):
    if fmt == "bf16":
        return x.to(torch.bfloat16).float()
    if fmt == "f16":
        return x.to(torch.float16).float()
    return x


def _view(X, mn_major, shift, off, n_mn, n_k):
    """Logical [n_mn, n_k] operand as the kernel sees it (rows beyond the array read as zero)."""
    rows, feats = X.shape
    Xp = torch.zeros(rows + 8, feats, dtype=X.dtype)
    Xp[:rows] = X
    if not mn_major:  # MN index = row, K index = feature
        return Xp[shift:shift + n_mn, off:off + n_k]
    return Xp[shift:shift + n_k, off:off + n_mn].t()


CASES = [
    # name, fmt, A shape, B shape, N, K, a_mn, b_mn, a_shift, b_shift, a_off, b_off, passes, tmem_col
    ("kk_bf16_n192_k96", "bf16", (128, 96), (192, 96), 192, 96, 0, 0, 0, 0, 0, 0, 1, 0),
    ("kk_bf16_n144", "bf16", (128, 96), (144, 96), 144, 96, 0, 0, 0, 0, 0, 0, 1, 0),
    ("kk_bf16_n256_k32", "bf16", (128, 32), (256, 32), 256, 32, 0, 0, 0, 0, 0, 0, 1, 0),
    ("kk_bf16_n96_k192", "bf16", (128, 192), (96, 192), 96, 192, 0, 0, 0, 0, 0, 0, 1, 0),
    ("kk_bf16_n16_k16", "bf16", (128, 16), (16, 16), 16, 16, 0, 0, 0, 0, 0, 0, 1, 0),
    ("kk_f16_n192_k96", "f16", (128, 96), (192, 96), 192, 96, 0, 0, 0, 0, 0, 0, 1, 0),
    ("kk_tf32_n192_k96", "tf32", (128, 96), (192, 96), 192, 96, 0, 0, 0, 0, 0, 0, 1, 0),
    ("kk_bf16_shift1", "bf16", (136, 192), (192, 192), 192, 192, 0, 0, 1, 0, 0, 0, 1, 0),
    ("kk_bf16_shift2", "bf16", (136, 192), (192, 192), 192, 192, 0, 0, 2, 0, 0, 0, 1, 0),
    ("kk_bf16_shift7", "bf16", (136, 96), (96, 96), 96, 96, 0, 0, 7, 0, 0, 0, 1, 0),
    ("kk_bf16_koff24_k32", "bf16", (128, 104), (256, 128), 256, 32, 0, 0, 0, 0, 24, 32, 1, 0),
    ("kk_bf16_passes2_col192", "bf16", (128, 96), (192, 96), 192, 96, 0, 0, 0, 0, 0, 0, 2, 192),
    ("mm_bf16_wgrad", "bf16", (256, 192), (256, 96), 96, 256, 1, 1, 0, 0, 64, 0, 1, 0),
    ("mm_bf16_wgrad_bshift1", "bf16", (248, 128), (250, 96), 96, 240, 1, 1, 0, 1, 0, 0, 1, 0),
    ("mm_bf16_wgrad_ashift2", "bf16", (250, 128), (250, 96), 96, 240, 1, 1, 2, 0, 0, 0, 1, 0),
    ("km_bf16_pv", "bf16", (128, 256), (256, 104), 32, 256, 0, 1, 0, 0, 0, 24, 1, 0),
    ("km_f16_pv", "f16", (128, 128), (128, 104), 32, 128, 0, 1, 0, 0, 0, 72, 1, 0),
    ("mk_bf16", "bf16", (64, 128), (96, 64), 96, 64, 1, 0, 0, 0, 0, 0, 1, 0),
    ("mm_f16_wgrad_bshift1", "f16", (248, 128), (250, 96), 96, 240, 1, 1, 0, 1, 0, 0, 1, 0),
    # NOTE: mixing fp16 and bf16 operands in one kind::f16 MMA raises cudaErrorIllegalInstruction on sm_100a (measured)
    # a_mn = 2: the A operand lives in tensor memory (tcgen05.st by the row threads), B K-major / MN-major in smem
    ("ts_f16_kmajor_b", "f16", (128, 64), (96, 64), 96, 64, 2, 0, 0, 0, 0, 0, 1, 0),
    ("ts_f16_pv", "f16", (128, 256), (256, 104), 32, 256, 2, 1, 0, 0, 0, 24, 1, 0),
    ("ts_bf16_passes2", "bf16", (128, 32), (16, 32), 16, 32, 2, 0, 0, 0, 0, 0, 2, 64),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_umma_selftest(case):
    name, fmt, ash, bsh, N, K, a_mn, b_mn, a_shift, b_shift, a_off, b_off, passes, tmem_col = case
    L = _lib.lib()
    g = torch.Generator().manual_seed(hash(name) % (2**31))
    A = torch.randn(*ash, generator=g)
    B = torch.randn(*bsh, generator=g)
    fmt_b = fmt
    if ":" in fmt:
        fmt, fmt_b = fmt.split(":")
    Ar, Br = _round(A, fmt), _round(B, fmt_b)
    Av = _view(Ar, a_mn == 1, a_shift, a_off, 128, K).double()
    Bv = _view(Br, b_mn, b_shift, b_off, N, K).double()
    # zero-pad views that run off the end of the array (the kernel reads zero rows / features there)
    Avp = torch.zeros(128, K, dtype=torch.float64)
    Avp[:Av.shape[0], :Av.shape[1]] = Av
    Bvp = torch.zeros(N, K, dtype=torch.float64)
    Bvp[:Bv.shape[0], :Bv.shape[1]] = Bv
    ref = (Avp @ Bvp.t()) * passes

    dA, dB = A.cuda(), B.cuda()
    dD = torch.full((128, N), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = L.nbss_umma_selftest(
        _lib.ptr(dA), ash[0], ash[1], _lib.ptr(dB), bsh[0], bsh[1], _lib.ptr(dD), N, K, a_mn, b_mn, FMT[fmt],
        a_shift, b_shift, a_off, b_off, passes, tmem_col, FMT[fmt_b], _lib.ptr(err), _lib.stream_ptr())
    _lib.check(st, "nbss_umma_selftest")
    torch.cuda.synchronize()
    assert int(err.item()) == 0, f"device error flag {int(err.item()):#x} (mbarrier timeout?)"
    out = dD.cpu().double()
    tol = 5e-3 if fmt == "tf32" else 2e-5
    rel = (out - ref).norm() / ref.norm()
    assert torch.isfinite(out).all(), f"{name}: non-finite output"
    assert rel < tol, f"{name}: rel-L2 {rel:.3e} (tol {tol}); max abs {float((out-ref).abs().max()):.3e}"
