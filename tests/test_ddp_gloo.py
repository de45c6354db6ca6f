"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path — one flat gradient buffer whose
views are the parameter gradients, a single all-reduce + 1/world scaling, and the r::world utterance split."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nbss_b200.spatialnet import SpatialNet


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(2)
    net = SpatialNet(dim_input=12, dim_output=4, dim_squeeze=8, num_layers=2, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4)
    flat, G, views = net.make_flat_grads("cpu")
    # every state-dict key has a gradient view; shared full.* keys alias ONE view; views tile the flat buffer exactly
    assert set(G.keys()) == set(net.state_dict().keys()) - {k for k in net.state_dict() if k.endswith("window")}
    assert G["layers.0.full.weight"].data_ptr() == G["layers.1.full.weight"].data_ptr()
    assert sum(v.numel() for v in views) == flat.numel() == sum(p.numel() for p in net.parameters())
    # pretend each rank computed gradient = rank+1 everywhere (written through the per-parameter views)
    for v in views:
        v.fill_(float(rank + 1))
    dist.all_reduce(flat)
    flat.mul_(1.0 / world)
    expect = sum(range(1, world + 1)) / world
    ok = bool(torch.all(flat == expect)) and all(bool(torch.all(v == expect)) for v in views)
    # utterance split of a global batch of 32: rank r takes r::world (my_distributed_sampler.py:78)
    mine = list(range(32))[rank::world]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    cover = sorted(sum(gathered, []))
    q.put((rank, ok, cover == list(range(32))))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok and cov for _, ok, cov in res), res
