#!/usr/bin/env python
"""Forward / gradient accuracy of whichever build NBSS_LIB selects, at the bench configuration (8 layers, F=129, T=250).
Used to A/B numerical variants (e.g. `make exact`: the two-MUFU sigmoid) on the GPU box:
    for L in "" _exact; do NBSS_LIB=nbss_b200/lib/libnbss_b200$L.so python tools/ab_forward.py; done
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nbss_b200.spatialnet import SpatialNet  # noqa: E402
from oracle import spatialnet_oracle as O  # noqa: E402

cfg = O.SMALL_CFG
res = []
for seed in (21, 22):
    P = O.synth_params(cfg, seed)
    net = SpatialNet(dim_input=12, dim_output=4, dim_squeeze=8, num_layers=8, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4).cuda().eval()
    net.load_state_dict({k: v.clone() for k, v in P.items()})
    x = torch.randn(1, 129, 250, 12, generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        y = net(x.cuda()).cpu()
        ref = O.spatialnet_forward({k: v.double() for k, v in P.items()}, x.double(), cfg)
    res.append(O.rel_l2(y, ref))
print(f"{os.environ.get('NBSS_LIB', 'default lib')}: forward rel-L2 vs fp64 oracle, 8 layers T=250: " + ", ".join(f"{e:.3e}" for e in res))
