import sys, torch
sys.path.insert(0, '.')
from nbss_b200 import ops
from oracle import spatialnet_oracle as O
torch.manual_seed(0)
B, F, T = 1, 129, 2
cfg = dict(O.SMALL_CFG, num_freqs=F)
P = O.synth_params(cfg, 5); Pd = {k: v.cuda() for k, v in P.items()}
pre = "layers.1.fconv1"
x = torch.randn(B, F, T, 96); dy = torch.randn(B, F, T, 96)
# oracle pieces
h = O.layer_norm(x, P[pre+".0.weight"], P[pre+".0.bias"])            # [B,F,T,96]
hh = h.permute(0, 2, 3, 1).reshape(B*T, 96, F)
W = P[pre+".1.weight"]
c = torch.nn.functional.conv1d(hh, W, P[pre+".1.bias"], padding=2, groups=8)
dyy = dy.permute(0, 2, 3, 1).reshape(B*T, 96, F)
dc = dyy * torch.where(c >= 0, torch.ones_like(c), P[pre+".2.weight"][None, :, None])
def dW_shift(sh):
    out = torch.zeros(96, 12, 5)
    hp = torch.nn.functional.pad(hh, (8, 8))
    for tap in range(5):
        hs = hp[:, :, 8 + tap - 2 + sh: 8 + tap - 2 + sh + F]          # h[f + tap - 2 + sh]
        full = torch.einsum('nof,nif->oi', dc, hs)                    # [96,96]
        for co in range(96):
            g = co // 12
            out[co, :, tap] = full[co, 12*g:12*g+12]
    return out
img = ops.fconv_pack(Pd[pre + ".1.weight"])
G = {k: torch.zeros_like(v) for k, v in Pd.items()}
dx, e = ops.fconv_tc_bwd(x.cuda(), dy.cuda(), Pd, pre, img, G)
torch.cuda.synchronize()
got = G[pre+".1.weight"].cpu()
for sh in (-2, -1, 0, 1, 2):
    ref = dW_shift(sh)
    print("shift", sh, "rel", O.rel_l2(got, ref), " per-tap", [round(O.rel_l2(got[:, :, t], ref[:, :, t]), 3) for t in range(5)])
ref = dW_shift(0)
print("got norm", float(got.norm()), "ref norm", float(ref.norm()))
print("per co-group rel", [round(O.rel_l2(got[12*g:12*g+12], ref[12*g:12*g+12]), 3) for g in range(8)])
print("got[0,:,2]", got[0, :, 2]); print("ref[0,:,2]", ref[0, :, 2])
print("got[13,:,2]", got[13, :, 2]); print("ref[13,:,2]", ref[13, :, 2])
