"""GPU: the parity holes VERDICT r1 named, on the benchmarked configuration itself.

* 8 layers, T in {250, 251}, wave -> wave + SI-SDR/PIT loss + backward: every parameter-gradient tensor against the
  oracle's autograd (fp32), tolerance stated per tensor family (contract SURVEY.md §8c: 1e-3 forward; gradients are held
  to the measured 16-bit-operand budget, and the part that exceeds it is shown to be the PReLU slope decision).
* the PReLU-sign claim of DESIGN §3 demonstrated: with the upstream gradient masked where |c| < tau the F-conv
  backward agrees to 2e-3.
* 20 optimisation steps: loss curve of the CUDA path (FlatClipAdam) against the fp32 oracle (clip_grad_norm_ + Adam).
* fp16 range guard: all weights x 8 -> finite and still on the oracle.
* stale weight images after a torch optimizer step / load_state_dict; two forwards + one backward;
  gradient accumulation under FlatClipAdam; deferred device-error check.
"""
import copy

import pytest
import torch

from nbss_b200 import ops
from nbss_b200.io import SeparationPipeline
from nbss_b200.loss import neg_si_sdr_pit
from nbss_b200.optim import FlatClipAdam
from nbss_b200.spatialnet import SpatialNet
from oracle import spatialnet_oracle as O


def _net(cfg, P):
    net = SpatialNet(dim_input=cfg["dim_input"], dim_output=cfg["dim_output"], dim_squeeze=8, num_layers=cfg["num_layers"],
                     num_freqs=cfg["num_freqs"], dim_hidden=96, dim_ffn=192, num_heads=4).cuda()
    net.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    return net


def _leaf(P, dtype=torch.float64):
    """Oracle parameters as autograd leaves.  fp64 by default: the fp32 oracle's OWN gradients are only good to ~3e-3 on some
    tensors (measured: fp32 vs fp64 autograd of the same restatement, layers.0.fconv1.0.bias 2.6e-3), so gradient parity is
    judged against fp64 autograd ("fp64 autograd as tie-breaker")."""
    seen, out = {}, {}
    for k, v in P.items():
        if id(v) not in seen:
            seen[id(v)] = v.to(dtype).clone().requires_grad_(True)
        out[k] = seen[id(v)]
    return out


def _family(name: str) -> str:
    if "fconv" in name:
        return "fconv"
    if "full" in name or "squeeze" in name:
        return "full"
    if "mhsa" in name:
        return "mhsa"
    if "tconvffn" in name:
        return "ffn"
    return name.split(".")[0]


# Gradient tolerances per parameter tensor (rel-L2 vs FP64 autograd of the oracle), 8 layers, fp16 operands.
# Measured (r02a): decoder 7e-4; everything upstream of the last F-conv 1e-2 .. 4.8e-2.  What sets that level is NOT the
# 16-bit rounding of the contractions (5e-3, second test) but the PReLU slope decision of the 16 F-conv sub-blocks: the
# kernels evaluate PReLU'(c) on c computed from fp16 operands, so the ~0.4 % of pre-activations with |c| < ~2e-3 may take the
# other slope, each F-conv adds ~5e-3 .. 1e-2 of relative error to the stream gradient that flows through it
# (test_fconv_prelu_sign_claim), and 16 of them add up to ~3e-2 on every parameter upstream.  The kernels' gradient is the
# exact gradient of the function the kernels compute; test_training_20_steps_loss_curve shows training is unaffected.
GRAD_TOL_8L = {"decoder": 2e-3, "ffn": 5e-2, "mhsa": 5e-2, "full": 5e-2, "fconv": 7e-2, "encoder": 5e-2}
GRAD_TOL_8L_SLOPE1 = {"decoder": 2e-3, "ffn": 8e-3, "mhsa": 8e-3, "full": 8e-3, "fconv": 8e-3, "encoder": 8e-3}


def _bench_config_errors(T, B, seed, slope_one):
    cfg = O.SMALL_CFG
    P = O.synth_params(cfg, seed)
    if slope_one:  # PReLU slope 1 = identity: the F-conv sub-block has no sign decision left
        for k in P:
            if k.endswith(".2.weight") and "fconv" in k:
                P[k] = torch.ones_like(P[k])
    Pl = _leaf(P)
    net = _net(cfg, P)
    pipe = SeparationPipeline(net, 256, 128, channels=None, ref_channel=0)
    g = torch.Generator().manual_seed(1000 + T)
    Ts = 128 * (T - 1)
    wave = 0.1 * torch.randn(B, 6, Ts, generator=g)
    tgt = 0.1 * torch.randn(B, 2, Ts, generator=g)
    est = pipe(wave.cuda())
    loss = neg_si_sdr_pit(est, tgt.cuda())[0]
    loss.backward()
    torch.cuda.synchronize()
    net.check_device_errors()
    est_ref = O.io_forward(Pl, wave.double(), cfg, 256, 128, 0)
    loss_ref = O.neg_si_sdr_pit(est_ref, tgt.double())[0]
    loss_ref.backward()
    e_fwd = O.rel_l2(est.detach().cpu(), est_ref.detach())
    errs = {n: O.rel_l2(p.grad.cpu().reshape(-1), Pl[n].grad.reshape(-1)) for n, p in net.named_parameters()}
    worst = {}
    for n, e in errs.items():
        f = _family(n)
        if f not in worst or e > worst[f][1]:
            worst[f] = (n, e)
    print(f"[T={T} slope_one={slope_one}] forward {e_fwd:.2e}; worst gradient rel-L2 per family: "
          + ", ".join(f"{f}: {n} {e:.2e}" for f, (n, e) in sorted(worst.items())))
    return e_fwd, loss.item(), loss_ref.item(), errs


@pytest.mark.gpu
@pytest.mark.parametrize("T", [250, 251])
def test_bench_config_fwd_bwd(T):
    """BASELINE configs[1]/[2] per utterance: 8 layers, 6ch, F=129, T=250 / 4 s (T=251), wave -> wave, SI-SDR + PIT loss,
    backward through iSTFT, the network and nothing else (the input needs no gradient)."""
    e_fwd, loss, loss_ref, errs = _bench_config_errors(T, 2, 77, slope_one=False)
    assert e_fwd < 1e-3, f"forward rel-L2 {e_fwd:.3e}"
    assert abs(loss - loss_ref) < 2e-3 * max(1.0, abs(loss_ref)), (loss, loss_ref)
    bad = {n: f"{e:.2e}" for n, e in errs.items() if not e < GRAD_TOL_8L[_family(n)]}
    assert not bad, f"gradient rel-L2 over the family tolerance {GRAD_TOL_8L}: {bad}"


@pytest.mark.gpu
def test_bench_config_fwd_bwd_without_sign_decisions():
    """The same 8-layer wave -> wave + loss + backward with every PReLU slope set to 1 (identity), i.e. with the only
    discontinuous derivative of the network removed: what is left is the 16-bit operand rounding, and every parameter
    gradient of the whole network is within 8e-3 of fp64 autograd — the proof that the 3e-2 of the test above is the
    slope decision and not an error of the backward kernels."""
    e_fwd, loss, loss_ref, errs = _bench_config_errors(250, 1, 78, slope_one=True)
    assert e_fwd < 1e-3, f"forward rel-L2 {e_fwd:.3e}"
    bad = {n: f"{e:.2e}" for n, e in errs.items() if not e < GRAD_TOL_8L_SLOPE1[_family(n)]}
    assert not bad, f"gradient rel-L2 over the family tolerance {GRAD_TOL_8L_SLOPE1}: {bad}"


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 129, 7), (1, 129, 64)])
def test_fconv_prelu_sign_claim(shape):
    """DESIGN §3 claims the F-conv backward's error above the 16-bit budget comes from PReLU'(c) being evaluated on the
    fp16-operand recomputation of c: elements with |c| within ~1e-3 of zero may take the other slope.  Proof by masking: zero
    the upstream gradient wherever the ORACLE pre-activation has |c| < tau, on both sides; what remains must agree to 2e-3
    (the unmasked comparison of the same tensors is printed for contrast)."""
    import torch.nn.functional as F_

    B, F, T = shape
    cfg = dict(O.SMALL_CFG, num_freqs=F)
    P = O.synth_params(cfg, 5)
    Pd = {k: v.cuda() for k, v in P.items()}
    pre = "layers.1.fconv1"
    g = torch.Generator().manual_seed(F * T + 3)
    x0 = torch.randn(B, F, T, 96, generator=g)
    dy0 = torch.randn(B, F, T, 96, generator=g)
    with torch.no_grad():  # the oracle's pre-activation c (oracle.fconv up to the conv), as [B,F,T,H]
        h = O.layer_norm(x0.double(), P[pre + ".0.weight"].double(), P[pre + ".0.bias"].double()).permute(0, 2, 3, 1).reshape(B * T, 96, F)
        c = F_.conv1d(h, P[pre + ".1.weight"].double(), P[pre + ".1.bias"].double(), padding=2, groups=8).reshape(B, T, 96, F).permute(0, 3, 1, 2)
    tau = 5e-3
    keep = (c.abs() >= tau).float()
    out = {}
    for tag, dy in (("unmasked", dy0), ("masked", dy0 * keep)):
        Pl = _leaf(P)
        x = x0.double().requires_grad_(True)
        (O.fconv(x, Pl, pre, 8)).backward(dy.double())  # branch only: the residual path carries dy unchanged on both sides
        img = ops.fconv_pack(Pd[pre + ".1.weight"])
        G = {k: torch.zeros_like(v) for k, v in Pd.items()}
        dx, e2 = ops.fconv_tc_bwd(x0.cuda(), dy.cuda(), Pd, pre, img, G)
        torch.cuda.synchronize()
        ops.check_err_flag(e2, "fconv_tc_bwd")
        errs = {"dx": O.rel_l2(dx.cpu() - dy, x.grad)}
        for k in (".0.weight", ".0.bias", ".1.weight", ".1.bias", ".2.weight"):
            errs[k] = O.rel_l2(G[pre + k].cpu().reshape(-1), Pl[pre + k].grad.reshape(-1))
        out[tag] = errs
    print(f"masked fraction {1 - keep.mean().item():.4f}; unmasked { {k: f'{v:.1e}' for k, v in out['unmasked'].items()} }; "
          f"masked { {k: f'{v:.1e}' for k, v in out['masked'].items()} }")
    bad = {k: f"{v:.2e}" for k, v in out["masked"].items() if not v < 2e-3}
    assert not bad, f"with the sign-ambiguous elements masked the F-conv backward must agree to 2e-3: {bad}"


@pytest.mark.gpu
def test_training_20_steps_loss_curve():
    """Training is unaffected by the 16-bit operands: 20 steps of clip(5) + Adam(1e-3) on one synthetic batch, the CUDA path
    (SeparationPipeline + CUDA loss + FlatClipAdam) against the fp32 oracle (autograd + clip_grad_norm_ + torch.optim.Adam)
    from identical initial weights.  Loss curves within 1 % (+0.02 dB)."""
    cfg = dict(O.SMALL_CFG, num_layers=4)
    P = O.synth_params(cfg, 31)
    net = _net(cfg, P)
    pipe = SeparationPipeline(net, 256, 128, channels=None, ref_channel=0)
    opt = FlatClipAdam(net, lr=1e-3, max_norm=5.0)
    Pl = _leaf(P, torch.float32)  # the reference trains in fp32
    leaves = list({id(v): v for v in Pl.values()}.values())
    opt_ref = torch.optim.Adam(leaves, lr=1e-3)
    g = torch.Generator().manual_seed(8)
    Ts = 128 * 63
    src = 0.1 * torch.randn(2, 2, Ts, generator=g)
    mix = src.sum(1, keepdim=True) * (1.0 + 0.3 * torch.randn(2, 6, 1, generator=g)) + 0.02 * torch.randn(2, 6, Ts, generator=g)
    xd, yd = mix.cuda(), src.cuda()
    curve, curve_ref = [], []
    for it in range(20):
        opt.zero_grad(set_to_none=True)
        loss = neg_si_sdr_pit(pipe(xd), yd)[0]
        loss.backward()
        opt.step()
        curve.append(loss.item())
        opt_ref.zero_grad(set_to_none=True)
        l_ref = O.neg_si_sdr_pit(O.io_forward(Pl, mix, cfg, 256, 128, 0), src)[0]
        l_ref.backward()
        torch.nn.utils.clip_grad_norm_(leaves, 5.0)
        opt_ref.step()
        curve_ref.append(l_ref.item())
    net.check_device_errors()
    print("loss curve  gpu:", " ".join(f"{v:.3f}" for v in curve))
    print("loss curve  ref:", " ".join(f"{v:.3f}" for v in curve_ref))
    assert curve_ref[-1] < curve_ref[0] - 0.5, "the reference run itself must make progress for the comparison to mean anything"
    for it, (a, b) in enumerate(zip(curve, curve_ref)):
        assert abs(a - b) <= 0.01 * abs(b) + 0.02, f"step {it}: gpu {a:.4f} vs oracle {b:.4f}"


@pytest.mark.gpu
def test_fp16_range_guard_scaled_weights():
    """fp16 operands have a 65504 range.  Every 16-bit operand of the path is either the output of a LayerNorm / GroupNorm
    (bounded by the affine gain) or at most three contractions away from one: the widest un-normalised chain is
    a1 -> c1 -> c2 inside the T-ConvFFN.  Scaling every T-ConvFFN and F-conv weight matrix by 8 (a gain of 8^3 = 512 on c2,
    64 on the sub-block outputs that are added to the fp32 stream) must stay finite and on the oracle.  (The attention
    projections are left alone: multiplying q and k by 8 turns the softmax into an arg-max, an ill-conditioned function
    that says nothing about range.)"""
    cfg = dict(O.SMALL_CFG, num_layers=3)
    P = O.synth_params(cfg, 61)
    P8 = {k: (v * 8.0 if (v.dim() >= 2 and ("tconvffn" in k or "fconv" in k)) else v.clone()) for k, v in P.items()}
    for i in range(cfg["num_layers"]):  # keep the shared full-band tensors shared
        for leaf in ("full.weight", "full.bias"):
            P8[f"layers.{i}.{leaf}"] = P8[f"layers.0.{leaf}"]
    net = _net(cfg, P8).eval()
    x = torch.randn(1, 129, 250, 12, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        y = net(x.cuda())
        net.check_device_errors()
        ref = O.spatialnet_forward({k: v.double() for k, v in P8.items()}, x.double(), cfg)
    assert torch.isfinite(y).all()
    e = O.rel_l2(y.cpu(), ref)
    print(f"T-ConvFFN / F-conv weights x 8: forward rel-L2 {e:.2e}, |y| max {ref.abs().max().item():.1f}")
    assert e < 3e-3, e


# ------------------------------------------------------------------------------------------------ ADVICE r1 regressions
@pytest.mark.gpu
def test_weight_images_follow_torch_optimizer_and_load_state_dict():
    """The cached UMMA weight images are keyed on the parameters' version counters: after torch.optim.Adam.step(),
    load_state_dict() or p.copy_() the next forward must use the new weights (== a forced rebuild)."""
    cfg = dict(O.SMALL_CFG, num_layers=2)
    net = _net(cfg, O.synth_params(cfg, 12))
    x = torch.randn(1, 129, 40, 12, generator=torch.Generator().manual_seed(1)).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    y0 = net(x)
    y0.square().mean().backward()
    opt.step()
    with torch.no_grad():
        y1 = net(x)
        net.engine.invalidate_images()
        y1f = net(x)
    assert not torch.equal(y0.detach(), y1), "the step changed the weights"
    assert torch.equal(y1, y1f), "stale weight images after torch.optim.Adam.step()"
    net.load_state_dict({k: v.clone() for k, v in O.synth_params(cfg, 13).items()})
    with torch.no_grad():
        y2 = net(x)
        net.engine.invalidate_images()
        y2f = net(x)
    assert torch.equal(y2, y2f) and not torch.equal(y2, y1), "stale weight images after load_state_dict()"
    with torch.no_grad():
        net.layers[0].mhsa.in_proj_weight.mul_(0.5)
        y3 = net(x)
        net.engine.invalidate_images()
        assert torch.equal(y3, net(x)), "stale weight images after an in-place update"


@pytest.mark.gpu
def test_two_forwards_one_backward_and_retain_graph():
    cfg = dict(O.SMALL_CFG, num_layers=2)
    net = _net(cfg, O.synth_params(cfg, 14))
    g = torch.Generator().manual_seed(2)
    xa, xb = torch.randn(1, 129, 24, 12, generator=g).cuda(), torch.randn(1, 129, 24, 12, generator=g).cuda()
    da, db = torch.randn(1, 129, 24, 4, generator=g).cuda(), torch.randn(1, 129, 24, 4, generator=g).cuda()
    grads = []
    for x, d in ((xa, da), (xb, db)):
        net.zero_grad(set_to_none=True)
        net(x).backward(d)
        grads.append([p.grad.clone() for p in net.parameters()])
    net.zero_grad(set_to_none=True)
    ya, yb = net(xa), net(xb)
    ((ya * da).sum() + (yb * db).sum()).backward()
    torch.cuda.synchronize()
    for p, ga, gb in zip(net.parameters(), *grads):
        # atomics: summation order differs, nothing else; the two terms may cancel (decoder bias: -54.60 + 55.16), so the
        # error is measured against the size of the terms, not of their sum
        assert (p.grad - (ga + gb)).norm().item() <= 1e-5 * (ga.norm() + gb.norm()).item()
    # a second backward through the same graph is refused with a clear message (activations are freed)
    y = net(xa)
    y.backward(da, retain_graph=True)
    with pytest.raises(Exception, match="twice"):
        y.backward(da)


@pytest.mark.gpu
def test_flat_clip_adam_gradient_accumulation():
    """Two backward passes without zero_grad: autograd accumulates into p.grad, the flat buffer of the last backward is
    stale; FlatClipAdam.step() must step on the accumulated gradient (it gathers p.grad when the views do not alias)."""
    cfg = dict(O.SMALL_CFG, num_layers=2)
    net = _net(cfg, O.synth_params(cfg, 15))
    ref = copy.deepcopy(net)
    ref_params = list(ref.parameters())
    opt, opt_ref = FlatClipAdam(net, lr=1e-3, max_norm=5.0), torch.optim.Adam(ref_params, lr=1e-3)
    g = torch.Generator().manual_seed(3)
    opt.zero_grad(set_to_none=True)
    for _ in range(2):
        x = torch.randn(1, 129, 24, 12, generator=g).cuda()
        net(x).backward(1e-3 * torch.randn(1, 129, 24, 4, generator=g).cuda())
    assert not net.grads_alias_flat()
    for p1, p2 in zip(net.parameters(), ref_params):
        p2.grad = p1.grad.detach().clone()
    torch.nn.utils.clip_grad_norm_(ref_params, 5.0)
    opt_ref.step()
    opt.step()
    torch.cuda.synchronize()
    for (n1, p1), p2 in zip(net.named_parameters(), ref_params):
        assert torch.allclose(p1, p2, rtol=0, atol=2e-6), (n1, (p1 - p2).abs().max().item())


@pytest.mark.gpu
def test_deferred_device_error_check_raises_and_resets():
    cfg = dict(O.SMALL_CFG, num_layers=1)
    net = _net(cfg, O.synth_params(cfg, 16)).eval()
    x = torch.randn(1, 129, 16, 12, generator=torch.Generator().manual_seed(5)).cuda()
    with torch.no_grad():
        net(x)
        torch.cuda.synchronize()
        net(x)  # polls the (clean) flag of the first call
        torch.cuda.synchronize()
        ops.device_err_flag(x.device).fill_(0x7001)  # what a timed-out mbarrier wait leaves behind
        net._post_device_error_check(x.device)       # the asynchronous copy every forward / backward ends with
        torch.cuda.synchronize()
        with pytest.raises(Exception, match="error flag"):
            net(x)
        torch.cuda.synchronize()
        assert int(ops.device_err_flag(x.device).item()) == 0, "the flag is reset so that the caller can recover"
        y = net(x)
        assert torch.isfinite(y).all()
