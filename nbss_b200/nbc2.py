"""Drop-in ``NBC2`` for the reference's ``models.arch.NBC2.NBC2`` (models/arch/NBC2.py:241-289) — INFERENCE (BASELINE.json
configs[3]: 8-channel input, F=257, batch 64).

Same constructor signature, same parameter names / shapes / creation order (reference checkpoints load; ``seed_everything``
gives the same initial weights), same ``forward(x[B,F,T,dim_input]) -> [B,F,T,dim_output]``.  The torch submodules are
parameter containers only; all arithmetic runs in libnbss_b200.so:

    encoder          Conv1d(k=5) along T                      nbss_encoder_fwd
    NBC2Block x L    x += MHSA(LN(x))   (2 heads x 48)        nbss_mhsa_fwd_nh (+ GroupBatchNorm partial sums in its store pass)
                     x += linear2(conv(linear1(GBN(x))))      nbss_gbn_reduce, nbss_nbc2_ffn_a, nbss_gbn_reduce, nbss_nbc2_ffn_b
    decoder          Linear                                   nbss_decoder_fwd

GroupBatchNorm (NBC2.py:57-149) takes its statistics over (all F bins of an utterance, channels) per frame, so the
T-ConvFFN is cut at its inner GroupBatchNorm into two slab kernels with a tiny cross-slab reduction between them.
Supported: the reference's NBC2_small (dim_hidden 96, dim_ffn 192, 2 heads, kernel 3, 8 groups, norms (LN, GBN, GBN),
share_along_sequence_dim False, dropout 0), T <= 256 frames, no attention-weight output (the reference discards it too,
NBC2.py:283-285).  There is no CPU path and no backward: parameters are treated as constants.
"""
from __future__ import annotations

from typing import Any, Dict

import torch
import torch.nn as nn
from torch import Tensor

from . import ops


class GroupBatchNorm(nn.Module):
    """Parameter container for models/arch/NBC2.py:57-149 (weight/bias [H] or, transposed, [H,1])."""

    def __init__(self, dim_hidden: int, group_size: int, share_along_sequence_dim: bool = False, transpose: bool = False,
                 affine: bool = True, eps: float = 1e-5) -> None:
        super().__init__()
        if share_along_sequence_dim or not affine or eps != 1e-5:
            raise NotImplementedError("nbss_b200.NBC2: GroupBatchNorm with share_along_sequence_dim / no affine / eps != 1e-5")
        self.dim_hidden, self.group_size, self.transpose = dim_hidden, group_size, transpose
        shape = [dim_hidden, 1] if transpose else [dim_hidden]
        self.weight = nn.Parameter(torch.ones(shape))
        self.bias = nn.Parameter(torch.zeros(shape))


class NBC2Block(nn.Module):
    """Parameters of one block in the reference's registration order (NBC2.py:154-194)."""

    def __init__(self, dim_hidden: int, dim_ffn: int, n_heads: int, dropout: float = 0, conv_kernel_size: int = 3,
                 n_conv_groups: int = 8, norms=("LN", "GBN", "GBN"), group_batch_norm_kwargs: Dict[str, Any] = None) -> None:
        super().__init__()
        gk = dict(group_batch_norm_kwargs or {})
        if tuple(norms) != ("LN", "GBN", "GBN") or dropout != 0 or conv_kernel_size != 3 or n_conv_groups != 8:
            raise NotImplementedError("nbss_b200.NBC2 supports norms (LN, GBN, GBN), kernel 3, 8 groups, dropout 0")
        self.norm1 = nn.LayerNorm(dim_hidden)
        self.self_attn = nn.MultiheadAttention(embed_dim=dim_hidden, num_heads=n_heads, batch_first=True)
        self.dropout1 = nn.Dropout(dropout)
        self.norm2 = GroupBatchNorm(dim_hidden=dim_hidden, transpose=False, **gk)
        self.linear1 = nn.Linear(dim_hidden, dim_ffn)
        self.conv = nn.Sequential(
            nn.SiLU(), nn.Conv1d(dim_ffn, dim_ffn, conv_kernel_size, padding="same", groups=n_conv_groups, bias=True),
            nn.SiLU(), nn.Conv1d(dim_ffn, dim_ffn, conv_kernel_size, padding="same", groups=n_conv_groups, bias=True),
            GroupBatchNorm(dim_hidden=dim_ffn, transpose=True, **gk),
            nn.SiLU(), nn.Conv1d(dim_ffn, dim_ffn, conv_kernel_size, padding="same", groups=n_conv_groups, bias=True),
            nn.SiLU(), nn.Dropout(dropout))
        self.linear2 = nn.Linear(dim_ffn, dim_hidden)
        self.dropout2 = nn.Dropout(dropout)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)
        nn.init.zeros_(self.linear1.bias)
        nn.init.zeros_(self.linear2.bias)


class NBC2(nn.Module):
    def __init__(self, dim_input: int, dim_output: int, n_layers: int, encoder_kernel_size: int = 5, dim_hidden: int = 192,
                 dim_ffn: int = 384, num_freqs: int = 257, block_kwargs: Dict[str, Any] = None):
        super().__init__()
        bk = dict(block_kwargs or {"n_heads": 2, "dropout": 0, "conv_kernel_size": 3, "n_conv_groups": 8, "norms": ("LN", "GBN", "GBN"),
                                   "group_batch_norm_kwargs": {"share_along_sequence_dim": False}})
        gk = dict(bk.get("group_batch_norm_kwargs", {}))
        gk["group_size"] = num_freqs  # NBC2.py:264
        bk["group_batch_norm_kwargs"] = gk
        if (dim_hidden, dim_ffn, bk.get("n_heads")) != (96, 192, 2) or encoder_kernel_size != 5:
            raise NotImplementedError("nbss_b200.NBC2 supports the NBC2_small configuration (dim_hidden 96, dim_ffn 192, 2 heads)")
        self.num_freqs, self.n_heads = num_freqs, bk["n_heads"]
        self.encoder = nn.Conv1d(dim_input, dim_hidden, encoder_kernel_size, stride=1, padding="same")
        self.sa_layers = nn.ModuleList([NBC2Block(dim_hidden=dim_hidden, dim_ffn=dim_ffn, **bk) for _ in range(n_layers)])
        self.decoder = nn.Linear(dim_hidden, dim_output)
        self._imgs = None
        self._img_key = None
        self._ws: dict = {}

    def _images(self, P: Dict[str, Tensor]):
        names = [f"sa_layers.{i}.{n}" for i in range(len(self.sa_layers))
                 for n in ("linear1.weight", "conv.1.weight", "conv.3.weight", "conv.6.weight", "linear2.weight",
                           "self_attn.in_proj_weight", "self_attn.out_proj.weight")]
        key = tuple((P[n].data_ptr(), P[n]._version) for n in names)
        if self._imgs is None or key != self._img_key or torch.cuda.is_current_stream_capturing():
            old = self._imgs
            self._imgs = [ops.nbc2_pack_block(P, f"sa_layers.{i}.", old[i] if old else None) for i in range(len(self.sa_layers))]
            self._img_key = key
        return self._imgs

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        if not x.is_cuda:
            raise ops._lib.NbssError("nbss_b200.NBC2 runs on CUDA tensors only (there is no CPU path)")
        B, F, T, _ = x.shape
        if F != self.num_freqs:
            raise ValueError(f"NBC2 was built for num_freqs={self.num_freqs} (GroupBatchNorm group size), got F={F}")
        P = {n: p.detach() for n, p in self.named_parameters()}
        with ops._lib.device_of(x):
            imgs = self._images(P)
            h = ops.encoder_fwd(x.detach().float().contiguous(), P)
            for i in range(len(self.sa_layers)):
                h = ops.nbc2_block_fwd(h, P, f"sa_layers.{i}.", imgs[i], num_heads=self.n_heads, ws=self._ws)
            return ops.decoder_fwd(h, P)

    def check_device_errors(self) -> None:
        for f in ops._ERR_FLAGS.values():
            ops.check_err_flag(f, "nbss_b200 kernel")
