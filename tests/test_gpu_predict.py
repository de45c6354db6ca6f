"""GPU: the predict-time callers (nbss_b200/predict.py) against the reference's own formulas: torch.linalg.lstsq for
recover_scale (models/utils/metrics.py:207-211, run here in fp64 on the CPU), SharedTrainer.predict_step's order of operations."""
import pytest
import torch

from nbss_b200.predict import load_reference_checkpoint, predict_step, recover_scale
from oracle import spatialnet_oracle as O


def _ref_recover(preds, mixture, norm):
    a = torch.linalg.lstsq(preds.double().transpose(-1, -2), mixture.double().unsqueeze(-1)).solution
    out = preds.double() * a
    if norm:
        mx = out.abs().amax(-1)
        out = out / torch.where(mx > 1, mx, torch.ones_like(mx)).unsqueeze(-1)
    return out, a[..., 0]


@pytest.mark.gpu
@pytest.mark.parametrize("B,S,Ts,norm", [(3, 2, 32000, True), (1, 2, 777, False), (2, 3, 5000, True), (4, 1, 4097, True)])
def test_recover_scale(B, S, Ts, norm):
    g = torch.Generator().manual_seed(B * 100 + S)
    src = torch.randn(B, S, Ts, generator=g)
    gains = 0.5 + 3.0 * torch.rand(B, S, 1, generator=g)                       # some rows exceed 1 after scaling
    mixture = (src * gains).sum(1) + 0.05 * torch.randn(B, Ts, generator=g)
    preds = src * (0.1 + torch.rand(B, S, 1, generator=g)) + 0.02 * torch.randn(B, S, Ts, generator=g)  # scale lost, as after SI-SDR training
    ref, a_ref = _ref_recover(preds, mixture, norm)
    out, a = recover_scale(preds.cuda(), mixture.cuda(), norm_if_exceed_1=norm, return_scales=True)
    torch.cuda.synchronize()
    assert O.rel_l2(a.cpu(), a_ref) < 1e-5
    assert O.rel_l2(out.cpu(), ref) < 1e-5
    if norm:
        assert out.abs().amax().item() <= 1.0 + 1e-6


@pytest.mark.gpu
def test_predict_step_order_of_operations():
    g = torch.Generator().manual_seed(4)
    B, C, Ts = 2, 6, 4096
    src = 0.3 * torch.randn(B, 2, Ts, generator=g)
    x = src.sum(1, keepdim=True).repeat(1, C, 1) + 0.01 * torch.randn(B, C, Ts, generator=g)
    fake_est = torch.stack([src[:, 1] * 0.01, src[:, 0] * 50.0], 1)  # swapped speakers, wrong scales

    # the estimates are exact scaled copies of the targets: |alpha t - p|^2 cancels to ~0 (or just below) in the Gram-sum form
    # of the loss kernel — the permutation must still come out right (regression: NaN compare kept the identity)
    out = predict_step(lambda w: fake_est.cuda(), x.cuda(), yr=src.cuda(), ref_channel=0, norm_if_exceed_1=True)
    ref, _ = _ref_recover(fake_est, x[:, 0], False)
    ref = ref[:, [1, 0]]                                             # PIT against the targets undoes the swap
    mx = ref.abs().amax(-1, keepdim=True)
    ref = ref / torch.where(mx > 1, mx, torch.ones_like(mx))
    assert O.rel_l2(out.cpu(), ref) < 1e-5


def test_load_reference_checkpoint_prefix_and_compiled_keys(tmp_path):
    """CPU: Lightning checkpoints carry the network under 'arch.' (compiled ones under 'arch._orig_mod.'); ensemble averages."""
    from nbss_b200.spatialnet import SpatialNet

    cfg = dict(O.SMALL_CFG, num_layers=1)
    Pa, Pb = O.synth_params(cfg, 1), O.synth_params(cfg, 2)
    paths = []
    for i, (P, pre) in enumerate(((Pa, "arch._orig_mod."), (Pb, "arch._orig_mod."), (Pa, "arch."))):
        sd = {pre + k: v for k, v in P.items()}
        sd["stft.window"] = torch.hann_window(256)
        path = tmp_path / f"epoch{i}_neg_si_sdr.ckpt"
        torch.save({"state_dict": sd}, path)
        paths.append(path)
    net = SpatialNet(dim_input=12, dim_output=4, dim_squeeze=8, num_layers=1, num_freqs=129, dim_hidden=96, dim_ffn=192, num_heads=4)
    load_reference_checkpoint(net, paths[1])
    assert torch.equal(net.state_dict()["decoder.weight"], Pb["decoder.weight"])
    load_reference_checkpoint(net, paths[2])  # plain (not compiled) checkpoint
    assert torch.equal(net.state_dict()["decoder.weight"], Pa["decoder.weight"])
    load_reference_checkpoint(net, paths[1], ensemble=[paths[0]])
    assert torch.allclose(net.state_dict()["encoder.weight"], 0.5 * (Pa["encoder.weight"] + Pb["encoder.weight"]), atol=1e-7)
