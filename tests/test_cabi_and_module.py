"""CPU: the C-ABI library loads and exports every symbol include/nbss_b200.h declares; the drop-in module exposes the
reference's parameter names / shapes; the product path refuses to run without CUDA (no fallback)."""
import os
import re
import sys

import pytest
import torch

from nbss_b200 import _lib
from nbss_b200.spatialnet import SpatialNet
from oracle import spatialnet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nbss_b200.h")).read()
    names = set(re.findall(r"\b(nbss_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    L = _lib.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_state_dict_contract_matches_reference_names_and_shapes():
    net = SpatialNet(dim_input=12, dim_output=4, dim_squeeze=8, num_layers=8, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4)
    sd = net.state_dict()
    shapes = O.param_shapes(O.SMALL_CFG)
    assert set(sd.keys()) == set(shapes.keys())
    for k, shp in shapes.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    # the full-band linear is one tensor shared by all layers (SpatialNet.py:192-195)
    assert sd["layers.0.full.weight"].data_ptr() == sd["layers.7.full.weight"].data_ptr()
    assert sum(p.numel() for p in net.parameters()) == 1_191_092
    # loading oracle-style parameters round-trips
    P = O.synth_params(O.SMALL_CFG, 1)
    net.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)


def test_no_cpu_fallback():
    net = SpatialNet(dim_input=12, dim_output=4, dim_squeeze=8, num_layers=1, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4)
    with pytest.raises(Exception):
        with torch.no_grad():
            net(torch.zeros(1, 129, 8, 12))


def test_unsupported_configuration_is_rejected():
    with pytest.raises(NotImplementedError):
        SpatialNet(dim_input=12, dim_output=4, dim_squeeze=16, num_layers=12, num_freqs=129, dim_hidden=192, dim_ffn=384, num_heads=4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="live reference only in the build container")
def test_same_init_as_reference_under_same_seed():
    sys.path.insert(0, "/root/reference")
    from models.arch.SpatialNet import SpatialNet as RefNet
    kw = dict(dim_input=12, dim_output=4, dim_squeeze=8, num_layers=2, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4)
    torch.manual_seed(2)
    ref = RefNet(**kw)
    torch.manual_seed(2)
    mine = SpatialNet(**kw)
    rsd, msd = ref.state_dict(), mine.state_dict()
    assert list(rsd.keys()) == list(msd.keys())
    for k in rsd:
        assert torch.equal(rsd[k], msd[k]), k
