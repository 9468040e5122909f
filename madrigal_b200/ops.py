"""torch.ops.madrigal_b200.* — the hot-path entry points as PyTorch custom ops (torch.library), so that code written
against `torch.ops` (and torch.compile / export graphs built around the reference model) can call them.  Each op is a
thin shim over the C ABI (ctypes, include/madrigal_b200.h); only a CUDA implementation is registered — calling an op
with CPU tensors raises, there is no fallback.  Shape-only "fake" implementations are registered as well, so that
tracing (torch.export / torch.compile's front end, meta tensors) sees the output shapes without running anything.

  torch.ops.madrigal_b200.pair_score(z_rows, z_cols, weight, precision, out_mode, normalize) -> Tensor [L, Nr, Nc] f32
  torch.ops.madrigal_b200.pair_score_gather(z_rows, z_cols, weight, labels, heads, tails, precision, sigmoid, normalize)
      -> Tensor [n] f32
  torch.ops.madrigal_b200.exact_normalized_ranks(scores) -> Tensor [L, N, N] f32
  torch.ops.madrigal_b200.fusion_encode(tokens [B,T,E], key_mask [B,T], src_mask [T,T]?, pool_key_mask [T]?,
      params (the TransformerFusion state_dict tensors, in state_dict order), cfg [embed_dim, num_tx_bottlenecks,
      num_layers, heads, head_dim, ffn_dim, norm_first], actn, agg, precision) -> Tensor [B, E] f32
"""
import torch

from . import decoder as _decoder

_LIB = torch.library.Library("madrigal_b200", "DEF")
_LIB.define("pair_score(Tensor z_rows, Tensor z_cols, Tensor weight, str precision, str out_mode, bool normalize) -> Tensor")
_LIB.define("pair_score_gather(Tensor z_rows, Tensor z_cols, Tensor weight, Tensor labels, Tensor heads, Tensor tails, "
            "str precision, bool sigmoid, bool normalize) -> Tensor")
_LIB.define("exact_normalized_ranks(Tensor scores) -> Tensor")
_LIB.define("fusion_encode(Tensor tokens, Tensor key_mask, Tensor? src_mask, Tensor? pool_key_mask, Tensor[] params, "
            "int[] cfg, str actn, str agg, str precision) -> Tensor")


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


_fusion_modules = {}


def _fusion_encode(tokens, key_mask, src_mask, pool_key_mask, params, cfg, actn, agg, precision):
    """Functional form of TransformerFusion.forward (models.py:401-455).  The parameter tensors are adopted, not
    copied: a parameter-less module shell is built on the meta device and the tensors are assigned into it; shells
    (with their prepared bf16 operand copies) are cached per parameter set."""
    from .fusion import TransformerFusion
    if len(cfg) != 7:
        raise ValueError("cfg = [embed_dim, num_tx_bottlenecks, num_layers, heads, head_dim, ffn_dim, norm_first]")
    key = (tuple(int(c) for c in cfg), actn, agg, precision, tuple(p.data_ptr() for p in params))
    mod = _fusion_modules.get(key)
    if mod is None:
        with torch.device("meta"):
            mod = TransformerFusion(int(cfg[0]), int(cfg[1]), int(cfg[2]), int(cfg[3]), int(cfg[4]), int(cfg[5]),
                                    transformer_actn=actn, transformer_norm_first=bool(cfg[6]),
                                    transformer_batch_first=False, transformer_agg=agg, precision=precision)
        names = list(mod.state_dict().keys())
        if len(names) != len(params):
            raise ValueError(f"expected {len(names)} parameter tensors in state_dict order, got {len(params)}")
        mod.load_state_dict(dict(zip(names, params)), assign=True)
        mod.eval()
        if agg == "x-attn":  # the constant pooling mask was created on the meta device: rebuild it for real
            m = torch.zeros(mod.x_attn_key_padding_mask.shape, dtype=torch.bool)
            if int(cfg[1]) > 0:
                from .constants import CELL_LINES, NUM_NON_TX_MODALITIES
                m[:, :NUM_NON_TX_MODALITIES] = True
                m[:, -len(CELL_LINES):] = True
            mod.x_attn_key_padding_mask = m
        if len(_fusion_modules) >= 16:
            _fusion_modules.pop(next(iter(_fusion_modules)))
        _fusion_modules[key] = mod
    if pool_key_mask is not None:
        mod.x_attn_key_padding_mask = pool_key_mask.reshape(1, -1).to(torch.bool)
    with torch.no_grad():
        return mod(tokens, key_mask, src_mask)


_LIB.impl("pair_score", _pair_score, "CUDA")
_LIB.impl("pair_score_gather", _pair_score_gather, "CUDA")
_LIB.impl("exact_normalized_ranks", _exact_normalized_ranks, "CUDA")
_LIB.impl("fusion_encode", _fusion_encode, "CUDA")


# ---- shape-only implementations for tracing (no arithmetic, no device work)
@torch.library.register_fake("madrigal_b200::pair_score")
def _pair_score_fake(z_rows, z_cols, weight, precision, out_mode, normalize):
    return z_rows.new_empty((weight.shape[0], z_rows.shape[0], z_cols.shape[0]), dtype=torch.float32)


@torch.library.register_fake("madrigal_b200::pair_score_gather")
def _pair_score_gather_fake(z_rows, z_cols, weight, labels, heads, tails, precision, sigmoid, normalize):
    return z_rows.new_empty((labels.numel(),), dtype=torch.float32)


@torch.library.register_fake("madrigal_b200::exact_normalized_ranks")
def _exact_normalized_ranks_fake(scores):
    return torch.empty_like(scores, dtype=torch.float32)


@torch.library.register_fake("madrigal_b200::fusion_encode")
def _fusion_encode_fake(tokens, key_mask, src_mask, pool_key_mask, params, cfg, actn, agg, precision):
    return tokens.new_empty((tokens.shape[0], tokens.shape[2]), dtype=torch.float32)
