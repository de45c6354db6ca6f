"""Stage-by-stage check of one NBC2 block against the oracle for a few shapes (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nbss_b200 import ops
from oracle import nbc2_oracle as N2, spatialnet_oracle as O

for (B, F, T) in [(1, 100, 250), (1, 160, 100), (1, 257, 250), (2, 150, 64)]:
    cfg = dict(N2.NBC2_SMALL, n_layers=1, num_freqs=F)
    P = N2.synth_params(cfg, 11)
    Pd = {k: v.cuda() for k, v in P.items()}
    pre = "sa_layers.0."
    x = torch.randn(B, F, T, 96, generator=torch.Generator().manual_seed(F + T))
    xr = x.reshape(B * F, T, 96).double()
    Pdd = {k: v.double() for k, v in P.items()}
    with torch.no_grad():
        y1 = xr + N2.mhsa(N2.layer_norm(xr, Pdd[pre + "norm1.weight"], Pdd[pre + "norm1.bias"]), Pdd, pre + "self_attn.", 2)
        y2 = N2.block(xr, Pdd, pre, cfg)
    img = ops.nbc2_pack_block(Pd, pre)
    xd = x.cuda()
    ya = torch.empty_like(xd)
    part = torch.zeros(B * F * T, 2, device="cuda")
    err = ops.device_err_flag(xd.device)
    st = ops._K("nbss_mhsa_fwd_nh")(ops.ptr(xd), ops.ptr(ya), B * F, T, ops.ptr(Pd[pre + "norm1.weight"]), ops.ptr(Pd[pre + "norm1.bias"]),
                                    ops.ptr(Pd[pre + "self_attn.in_proj_bias"]), ops.ptr(Pd[pre + "self_attn.out_proj.bias"]), ops.ptr(img),
                                    ops.ptr(None), ops.ptr(None), ops.ptr(None), ops.ptr(None), ops.ptr(part), 2, ops.FMT_F16, ops.ptr(err), ops.stream_ptr())
    torch.cuda.synchronize()
    e1 = O.rel_l2(ya.cpu().reshape(B * F, T, 96) - xr.float(), (y1 - xr))
    yb = ops.nbc2_block_fwd(xd.clone(), Pd, pre, img)
    torch.cuda.synchronize()
    e2 = O.rel_l2(yb.cpu().reshape(B * F, T, 96), y2)
    nan_slabs = torch.isnan(yb).reshape(B * F, -1).any(1).nonzero().flatten().tolist()
    nan_a = torch.isnan(ya).reshape(B * F, -1).any(1).nonzero().flatten().tolist()
    print(f"shape {(B, F, T)}: mhsa branch rel-L2 {e1:.2e} (nan slabs {nan_a[:8]}), block rel-L2 {e2:.2e} (nan slabs {nan_slabs[:8]} of {len(nan_slabs)}), flag {int(err.item())}")
