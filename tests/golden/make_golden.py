"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (imported from /root/reference) on
seeded inputs with the deterministic synthetic parameters of oracle.spatialnet_oracle.synth_params.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle_golden.py); they hold no reference source, only numbers.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from models.arch.SpatialNet import SpatialNet  # noqa: E402  (reference)
from models.io.norm import Norm  # noqa: E402
from models.io.stft import STFT  # noqa: E402

from oracle import spatialnet_oracle as O  # noqa: E402

torch.set_num_threads(8)

TINY = dict(dim_input=4, dim_output=4, dim_squeeze=4, num_layers=2, num_freqs=17, encoder_kernel_size=5,
            dim_hidden=32, dim_ffn=64, num_heads=4, kernel_size=(5, 3), conv_groups=(8, 8))
CFG1 = dict(dim_input=4, dim_output=4, dim_squeeze=8, num_layers=8, num_freqs=65, encoder_kernel_size=5,
            dim_hidden=96, dim_ffn=192, num_heads=4, kernel_size=(5, 3), conv_groups=(8, 8))  # BASELINE configs[0]


LARGE = dict(dim_input=12, dim_output=4, dim_squeeze=16, num_layers=2, num_freqs=129, encoder_kernel_size=5,
             dim_hidden=192, dim_ffn=384, num_heads=4, kernel_size=(5, 3), conv_groups=(8, 8))  # the YAML's "large" widths, 2 layers


def build_ref(cfg, P):
    m = SpatialNet(**cfg)
    sd = {k: v.clone() for k, v in P.items()}
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m


def ref_io_forward(m, stft, norm, x, ref_channel=0):
    """SharedTrainer.py:113-131 restated with the reference's own STFT/Norm/arch objects (no arithmetic here)."""
    X, stft_paras = stft.stft(x)
    B, C, F, T = X.shape
    X, (Xr, XrMM) = norm.norm(X, ref_channel=ref_channel)
    X = X.permute(0, 2, 3, 1)
    X = torch.view_as_real(X).reshape(B, F, T, -1)
    out = m(X)
    out = torch.view_as_complex(out.float().reshape(B, F, T, -1, 2))
    out = out.permute(0, 3, 1, 2)
    Yr_hat = norm.inorm(out, (Xr, XrMM))
    return stft.istft(Yr_hat, stft_paras)


def main():
    out = {}
    # 1) tiny config: forward + every parameter gradient + input gradient
    g = torch.Generator().manual_seed(11)
    P = O.synth_params(TINY, seed=101)
    m = build_ref(TINY, P)
    x = torch.randn(2, 17, 20, 4, generator=g, requires_grad=True)
    dy = torch.randn(2, 17, 20, 4, generator=g)
    y = m(x)
    y.backward(dy)
    tiny = {"x": x.detach().numpy(), "dy": dy.numpy(), "y": y.detach().numpy(), "dx": x.grad.numpy()}
    for k, p in m.named_parameters():
        tiny["grad." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "tiny_fwd_bwd.npz"), **tiny)

    # 2) BASELINE configs[0]: small net, 2ch, F=65, T=64, batch 1, forward on CPU + gradient norms
    g = torch.Generator().manual_seed(12)
    P = O.synth_params(CFG1, seed=102)
    m = build_ref(CFG1, P)
    x = torch.randn(1, 65, 64, 4, generator=g)
    dy = torch.randn(1, 65, 64, 4, generator=g)
    y = m(x)
    y.backward(dy)
    c1 = {"x": x.numpy(), "dy": dy.numpy(), "y": y.detach().numpy()}
    for k, p in m.named_parameters():
        c1["gnorm." + k] = np.float64(p.grad.double().norm().item())
    np.savez_compressed(os.path.join(HERE, "cfg1_small_2ch_f65_t64.npz"), **c1)

    # 3) real small config (6ch, F=129), short T: forward only (weights regenerated from the seed, not stored)
    g = torch.Generator().manual_seed(13)
    P = O.synth_params(O.SMALL_CFG, seed=103)
    m = build_ref(O.SMALL_CFG, P)
    x = torch.randn(1, 129, 12, 12, generator=g)
    with torch.no_grad():
        y = m(x)
    np.savez_compressed(os.path.join(HERE, "small_6ch_f129_t12.npz"), x=x.numpy(), y=y.numpy())

    # 5) the large configuration's layer widths (configs/SpatialNet.yaml:16-25 comments), 2 layers, short T:
    #    forward + gradient norms, pins the oracle for the next tile shapes (H=192, Hf=384, dim_squeeze=16, head dim 48)
    g = torch.Generator().manual_seed(15)
    P = O.synth_params(LARGE, seed=105)
    m = build_ref(LARGE, P)
    x = torch.randn(1, 129, 10, 12, generator=g)
    dy = torch.randn(1, 129, 10, 4, generator=g)
    y = m(x)
    y.backward(dy)
    lg = {"x": x.numpy(), "dy": dy.numpy(), "y": y.detach().numpy()}
    for k, p in m.named_parameters():
        lg["gnorm." + k] = np.float64(p.grad.double().norm().item())
    np.savez_compressed(os.path.join(HERE, "large_widths_f129_t10.npz"), **lg)

    # 6) NBC2 (BASELINE configs[3]): small widths, 2 layers, F=17: forward + gradient norms
    from models.arch.NBC2 import NBC2  # noqa: E402  (reference)
    from oracle import nbc2_oracle as N2  # noqa: E402

    ncfg = dict(N2.NBC2_SMALL, n_layers=2, num_freqs=17)
    g = torch.Generator().manual_seed(16)
    Pn = N2.synth_params(ncfg, seed=106)
    mn = NBC2(dim_input=16, dim_output=4, n_layers=2, dim_hidden=96, dim_ffn=192, num_freqs=17,
              block_kwargs={'n_heads': 2, 'dropout': 0, 'conv_kernel_size': 3, 'n_conv_groups': 8, 'norms': ("LN", "GBN", "GBN"),
                            'group_batch_norm_kwargs': {'share_along_sequence_dim': False}})
    missing, unexpected = mn.load_state_dict({k: v.clone() for k, v in Pn.items()}, strict=True)
    assert not missing and not unexpected
    x = torch.randn(2, 17, 12, 16, generator=g)
    dy = torch.randn(2, 17, 12, 4, generator=g)
    y = mn(x)
    y.backward(dy)
    nb = {"x": x.numpy(), "dy": dy.numpy(), "y": y.detach().numpy()}
    for k, p in mn.named_parameters():
        nb["gnorm." + k] = np.float64(p.grad.double().norm().item())
    np.savez_compressed(os.path.join(HERE, "nbc2_small_f17_t12.npz"), **nb)

    # 4) framing: STFT / Norm / iSTFT and the whole wave->wave path with the tiny-ish net (n_fft 32 -> F=17)
    g = torch.Generator().manual_seed(14)
    stft, norm = STFT(n_fft=32, n_hop=16), Norm(mode="frequency")
    wave = 0.1 * torch.randn(2, 2, 16 * 19, generator=g)
    X, _ = stft.stft(wave)
    P = O.synth_params(TINY, seed=101)
    m = build_ref(TINY, P)
    with torch.no_grad():
        yw = ref_io_forward(m, stft, norm, wave.clone())
    io = {"wave": wave.numpy(), "X_re": X.real.numpy(), "X_im": X.imag.numpy(), "y_wave": yw.numpy()}
    # 8 kHz parameters of the YAML (n_fft 256, hop 128), framing only
    stft8 = STFT(n_fft=256, n_hop=128)
    wave8 = 0.1 * torch.randn(1, 2, 128 * 9, generator=g)
    X8, _ = stft8.stft(wave8)
    Xn8, (Xr8, XrMM8) = norm.norm(X8.clone(), ref_channel=0)
    w8 = stft8.istft(X8, wave8.shape[-1])
    io.update({"wave8": wave8.numpy(), "X8_re": X8.real.numpy(), "X8_im": X8.imag.numpy(),
               "Xn8_re": Xn8.real.numpy(), "Xn8_im": Xn8.imag.numpy(), "XrMM8": XrMM8.numpy(), "rt8": w8.numpy()})
    np.savez_compressed(os.path.join(HERE, "framing.npz"), **io)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
