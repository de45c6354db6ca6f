"""Times nbss_mhsa_fwd alone at the bench shape (batch 32: 4128 slabs, T=250); NBSS_MHSA_PTMEM selects the P-in-TMEM path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbss_b200 import ops  # noqa: E402
from oracle import spatialnet_oracle as O  # noqa: E402  (parameters only)

P = {k: v.cuda() for k, v in O.synth_params(O.SMALL_CFG, 5).items()}
pre = "layers.2."
x = torch.randn(32, 129, 250, 96, device="cuda")
img = ops.pack_layer_weights(P, pre)
for _ in range(3):
    ops.mhsa_fwd(x, P, pre, img, save=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.mhsa_fwd(x, P, pre, img, save=True)
e1.record()
torch.cuda.synchronize()
print(f"NBSS_MHSA_PTMEM={os.environ.get('NBSS_MHSA_PTMEM', '0')}: mhsa_fwd {e0.elapsed_time(e1) / 10:.4f} ms/launch")
