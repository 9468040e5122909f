"""torch.ops.madrigal_b200.* — the hot-path entry points as PyTorch custom ops (torch.library), so that code written
against `torch.ops` (and torch.compile / export graphs built around the reference model) can call them.  Each op is a
thin shim over the C ABI (ctypes, include/madrigal_b200.h); only a CUDA implementation is registered — calling an op
with CPU tensors raises, there is no fallback.

  torch.ops.madrigal_b200.pair_score(z_rows, z_cols, weight, precision, out_mode, normalize) -> Tensor [L, Nr, Nc] f32
  torch.ops.madrigal_b200.pair_score_gather(z_rows, z_cols, weight, labels, heads, tails, precision, sigmoid, normalize)
      -> Tensor [n] f32
  torch.ops.madrigal_b200.exact_normalized_ranks(scores) -> Tensor [L, N, N] f32
"""
import torch

from . import decoder as _decoder

_LIB = torch.library.Library("madrigal_b200", "DEF")
_LIB.define("pair_score(Tensor z_rows, Tensor z_cols, Tensor weight, str precision, str out_mode, bool normalize) -> Tensor")
_LIB.define("pair_score_gather(Tensor z_rows, Tensor z_cols, Tensor weight, Tensor labels, Tensor heads, Tensor tails, "
            "str precision, bool sigmoid, bool normalize) -> Tensor")
_LIB.define("exact_normalized_ranks(Tensor scores) -> Tensor")


def _pair_score(z_rows, z_cols, weight, precision, out_mode, normalize):
    if out_mode not in ("logit", "sigmoid"):
        raise ValueError("torch.ops.madrigal_b200.pair_score: out_mode must be 'logit' or 'sigmoid' "
                         "(rank / top-k outputs take a RankTable: use madrigal_b200.pair_score / pair_topk)")
    return _decoder.pair_score(z_rows, z_cols, weight, precision=precision, out=out_mode, normalize=normalize)


def _pair_score_gather(z_rows, z_cols, weight, labels, heads, tails, precision, sigmoid, normalize):
    return _decoder.pair_score_gather(z_rows, z_cols, weight, labels, heads, tails, precision=precision,
                                      out="sigmoid" if sigmoid else "logit", normalize=normalize)


def _exact_normalized_ranks(scores):
    from . import normalize
    return normalize.exact_normalized_ranks(scores)


_LIB.impl("pair_score", _pair_score, "CUDA")
_LIB.impl("pair_score_gather", _pair_score_gather, "CUDA")
_LIB.impl("exact_normalized_ranks", _exact_normalized_ranks, "CUDA")
