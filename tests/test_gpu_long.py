"""GPU: inference on utterances longer than 256 frames (the reference validates / tests on whole utterances,
SharedTrainer.py:134-189): the chunked long-sequence kernels (mhsa_fwd.cu LONG = 1 / 2, ffn_fwd.cu MODE 3 / 4) against the oracle."""
import pytest
import torch

from nbss_b200 import ops
from nbss_b200.spatialnet import SpatialNet
from oracle import spatialnet_oracle as O

CFG = O.SMALL_CFG
# 257: one frame beyond a tile; 508 / 509 / 510: last chunk of the FFN passes has 4 / 5 / 6 and 0 / 1 / 2 frames beyond the previous
# chunk's outputs; 512: T a multiple of the key block; 700: three chunks, partial last key block; 1100: five
LONG_T = [257, 300, 508, 509, 510, 512, 700, 1100]


def _params(seed=5):
    P = O.synth_params(CFG, seed)
    return P, {k: v.cuda() for k, v in P.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("T", LONG_T)
def test_mhsa_fwd_long(T):
    P, Pd = _params()
    pre = "layers.2."
    x = torch.randn(2, 3, T, 96, generator=torch.Generator().manual_seed(100 + T))
    with torch.no_grad():
        ref = O.mhsa(x.double(), {k: v.double() for k, v in P.items()}, pre, 4)
    img = ops.pack_layer_weights(Pd, pre)
    y, err = ops.mhsa_fwd(x.cuda(), Pd, pre, img)
    torch.cuda.synchronize()
    ops.check_err_flag(err, "mhsa_fwd_long")
    assert torch.isfinite(y).all()
    e = O.rel_l2(y.cpu().double() - x.double(), ref)
    assert e < 1e-3, f"branch rel-L2 {e:.3e}"
    # in place, several work items per CTA (more slabs than SMs), stale tiles from previous items
    xb = torch.randn(4, 129, T if T < 600 else 300, 96, generator=torch.Generator().manual_seed(T))
    with torch.no_grad():
        refb = O.mhsa(xb[:, :2], P, pre, 4)
    h = xb.cuda()
    y2, err = ops.mhsa_fwd(h, Pd, pre, img, out=h)
    torch.cuda.synchronize()
    ops.check_err_flag(err, "mhsa_fwd_long")
    assert O.rel_l2(y2[:, :2].cpu() - xb[:, :2], refb) < 1e-3
    assert torch.isfinite(y2).all()


@pytest.mark.gpu
@pytest.mark.parametrize("T", LONG_T)
def test_ffn_fwd_long(T):
    P, Pd = _params()
    pre = "layers.3."
    x = torch.randn(2, 5, T, 96, generator=torch.Generator().manual_seed(T))
    with torch.no_grad():
        ref = O.tconvffn(x.double(), {k: v.double() for k, v in P.items()}, pre, 8)
    img = ops.pack_layer_weights(Pd, pre)
    y, err = ops.ffn_fwd(x.cuda(), Pd, pre, img)
    torch.cuda.synchronize()
    ops.check_err_flag(err, "ffn_fwd_long")
    assert torch.isfinite(y).all()
    e = O.rel_l2(y.cpu().double() - x.double(), ref)
    assert e < 1e-3, f"branch rel-L2 {e:.3e}"
    xb = torch.randn(3, 129, T if T < 600 else 300, 96, generator=torch.Generator().manual_seed(T + 1))
    with torch.no_grad():
        refb = O.tconvffn(xb[:, 127:], P, pre, 8)
    h = xb.cuda()
    y2, err = ops.ffn_fwd(h, Pd, pre, img, out=h)
    torch.cuda.synchronize()
    ops.check_err_flag(err, "ffn_fwd_long")
    assert O.rel_l2(y2[:, 127:].cpu() - xb[:, 127:], refb) < 1e-3
    assert torch.isfinite(y2).all()


@pytest.mark.gpu
def test_long_path_error_level_matches_one_slab_kernels():
    """The long path on T = 300 agrees with the oracle about as closely as the one-slab kernel does on T = 250 (same weights)."""
    P, Pd = _params()
    pre = "layers.1."
    img = ops.pack_layer_weights(Pd, pre)
    errs = {}
    for T in (250, 300):
        x = torch.randn(1, 4, T, 96, generator=torch.Generator().manual_seed(7))
        with torch.no_grad():
            ref = O.tconvffn(x.double(), {k: v.double() for k, v in P.items()}, pre, 8)
        y, _ = ops.ffn_fwd(x.cuda(), Pd, pre, img)
        errs[T] = O.rel_l2(y.cpu().double() - x.double(), ref)
    assert errs[300] < 1e-3 and errs[300] < 3 * errs[250] + 1e-4, errs


@pytest.mark.gpu
@pytest.mark.parametrize("T", [300, 777, 2001])
def test_network_forward_long_and_training_refuses(T):
    cfg = dict(CFG, num_layers=2)
    P = O.synth_params(cfg, 23)
    net = SpatialNet(dim_input=12, dim_output=4, dim_squeeze=8, num_layers=2, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4).cuda()
    net.load_state_dict({k: v.clone() for k, v in P.items()})
    x = torch.randn(1, 129, T, 12, generator=torch.Generator().manual_seed(T))
    with torch.no_grad():
        ref = O.spatialnet_forward(P, x, cfg)
        y = net(x.cuda())
    torch.cuda.synchronize()
    e = O.rel_l2(y.cpu(), ref)
    print(f"2-layer network, T = {T}: rel-L2 {e:.2e}")
    assert e < 1e-3
    with pytest.raises(NotImplementedError):
        net(x.cuda())  # grad mode on: the training path saves for backward, and the backward kernels hold T <= 256


@pytest.mark.gpu
def test_pipeline_wave_to_wave_long():
    """STFT -> norm -> network -> inverse norm -> iSTFT on a 6.4 s recording (T = 401 frames) under no_grad, against the oracle."""
    from nbss_b200.io import SeparationPipeline

    cfg = dict(CFG, num_layers=2)
    P = O.synth_params(cfg, 29)
    net = SpatialNet(dim_input=12, dim_output=4, dim_squeeze=8, num_layers=2, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4).cuda()
    net.load_state_dict({k: v.clone() for k, v in P.items()})
    pipe = SeparationPipeline(net, 256, 128, channels=None, ref_channel=0)
    x = 0.1 * torch.randn(1, 6, 128 * 400, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = O.io_forward(P, x, cfg, 256, 128, 0)
        y = pipe(x.cuda())
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert O.rel_l2(y.cpu(), ref) < 1e-3
