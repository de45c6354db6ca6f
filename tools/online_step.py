"""A few streaming steps of OnlineSpatialNet for an ncu launch list:  ncu --metrics gpu__time_duration.sum --csv ... python tools/online_step.py B"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nbss_b200.online import OnlineSpatialNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(2)
net = OnlineSpatialNet(dim_input=12, dim_output=4, num_layers=8, dim_squeeze=8, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4,
                       attention="mhsa(251)").to(dev).eval()
state = net.init_state(B)
x = torch.randn(B, 129, 12, device=dev)
for _ in range(3):
    net.step(x, state)
state.pos.fill_(4 * state.scope)  # full ring: steady state of a long stream
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
net.step(x, state)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
