"""CPU oracle for the drug-pair scoring path of biopharmaai/Madrigal.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; nothing under madrigal_b200/ does.  It is a plain numpy restatement of the reference's algorithm for the
path, each function citing the reference file:line it follows (paths relative to the reference root).

Parity status: the reference has NO tests, golden vectors or fixtures for this path, so the oracle
is pinned the other way the task allows: tests/golden/make_golden.py imports the UNMODIFIED reference modules
(oracle/ref_import.py) in the build container, runs them on seeded synthetic weights/inputs and commits the outputs
under tests/golden/; tests/test_oracle_golden.py checks every function here against those fixtures.  The third-party
arithmetic the reference delegates to (torch 1.13.1 nn.TransformerEncoderLayer / nn.MultiheadAttention / F.gelu /
F.layer_norm; numpy argsort) is restated from its published semantics, call sites cited below.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

try:  # erf for exact GELU: scipy if present, else math.erf vectorised
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

F32 = np.float32


# ------------------------------------------------------------------------------------------------------ decoder (a-4)
def symmetric(P: np.ndarray) -> np.ndarray:
    """Symmetric.forward, models.py:522-524: W = triu(P) + triu(P, 1)^T (exactly symmetric)."""
    return np.triu(P) + np.swapaxes(np.triu(P, 1), -1, -2)


def l2_normalize(z: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """F.normalize(z) as used at models.py:947-949 (p=2, dim=1, eps=1e-12)."""
    n = np.sqrt((z.astype(F32) ** 2).sum(axis=1, keepdims=True, dtype=F32))
    return (z / np.maximum(n, F32(eps))).astype(F32)


def bilinear_scores(z1: np.ndarray, z2: np.ndarray, W: np.ndarray,
                    label_range: Optional[Tuple[int, int]] = None, dtype=F32) -> np.ndarray:
    """BilinearDDIScorer.forward/bilinear, models.py:537-547: matmul(matmul(z1, W[l0:l1]), z2.T) -> [L', N1, N2].

    Association order as in the reference ((z1 W) z2^T).  `dtype=np.float64` gives a higher-precision reference for
    error budgeting; float32 mirrors the reference's arithmetic type.
    """
    if label_range is not None:
        assert len(label_range) == 2
        W = W[label_range[0]:label_range[1]]
    z1 = z1.astype(dtype)
    z2 = z2.astype(dtype)
    y = np.matmul(z1[None, :, :], W.astype(dtype))  # [L, N1, D]
    return np.matmul(y, z2.T[None, :, :])  # [L, N1, N2]


def sigmoid(x: np.ndarray) -> np.ndarray:
    """predict.py:358 (numpy) / predict.py:235 (torch.sigmoid)."""
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


# ------------------------------------------------------------------------------------------- rank normalisation (a-5)
def classwise_normalized_rank(tensor: np.ndarray, kind: Optional[str] = None) -> np.ndarray:
    """classwise_normalized_rank_3d_numpy, notebooks/normalize_scores.py:36-60.

    flat rank = argsort(argsort(flat)) + 1 over each class's flattened N1*N2 scores, divided by N1*(N2-1)/2
    (float64).  `kind=None` uses numpy's default sort exactly as the reference (tie order implementation-defined);
    kind='stable' gives the deterministic tie order the CUDA exact-rank kernel implements.
    """
    L = tensor.shape[0]
    flat = tensor.reshape(L, -1)
    order = flat.argsort(axis=1, kind=kind)
    rank = np.empty_like(order)
    rows = np.arange(L)[:, None]
    rank[rows, order] = np.arange(1, flat.shape[1] + 1)[None, :]
    norm = rank / (tensor.shape[1] * (tensor.shape[2] - 1) / 2)
    return norm.reshape(tensor.shape)


def normalize_scores(raw_scores: np.ndarray, kind: Optional[str] = None) -> np.ndarray:
    """run_slice for every outcome, notebooks/normalize_scores.py:33, 62-74.

    Per outcome: entries with col >= row set to 1e7 (np.triu_indices(N, k=0)), ranked with
    classwise_normalized_rank, masked entries zeroed, result + its transpose, stored as float32.
    """
    L, N, N2 = raw_scores.shape
    iu = np.triu_indices(N, k=0, m=N2)
    out = np.empty((L, N, N2), dtype=F32)
    for l in range(L):
        s = raw_scores[l:l + 1].copy()
        s[:, iu[0], iu[1]] = 1e7
        r = classwise_normalized_rank(s, kind=kind)
        r[:, iu[0], iu[1]] = 0
        r = r + r.swapaxes(1, 2)
        out[l:l + 1] = r
    return out


def lower_triangle_values(scores_l: np.ndarray) -> np.ndarray:
    """The M = N(N-1)/2 scores the reference ranks for one outcome: row > col, in row-major order."""
    i, j = np.tril_indices(scores_l.shape[0], k=-1, m=scores_l.shape[1])
    return scores_l[i, j]


def reference_quantiles(scores: np.ndarray, Q: int) -> np.ndarray:
    """Q order statistics of each outcome's strict-lower-triangle scores, at ranks ceil(i*M/Q), i = 1..Q.

    This is the 'per-outcome reference distribution' the fused rank epilogue looks scores up in; with Q == M it is
    the full sorted sample and the lookup equals the reference's exact rank for untied scores.
    """
    L = scores.shape[0]
    out = np.empty((L, Q), dtype=F32)
    for l in range(L):
        v = np.sort(lower_triangle_values(scores[l]).astype(F32), kind="stable")
        M = v.shape[0]
        ranks = -(-(np.arange(1, Q + 1, dtype=np.int64) * M) // Q)  # ceil(i*M/Q), 1-based
        out[l] = v[ranks - 1]
    return out


def quantile_rank(thresholds: np.ndarray, logits: np.ndarray, side: str = "right") -> np.ndarray:
    """Rank of each logit against its outcome's threshold table: np.searchsorted(thresholds[l], x, side) in fp32.

    thresholds [L, Q] ascending fp32, logits [L, ...] fp32 -> uint16 [L, ...].  This is the lookup the fused CUDA
    epilogue must reproduce bit for bit (north-star: 'ranks bit-exact against the reference quantile lookup').
    """
    assert thresholds.dtype == F32 and logits.dtype == F32
    out = np.empty(logits.shape, dtype=np.uint16)
    for l in range(thresholds.shape[0]):
        out[l] = np.searchsorted(thresholds[l], logits[l].reshape(-1), side=side).reshape(logits[l].shape)
    return out


def ensemble_mean_sigmoid(logits_per_ckpt: Sequence[np.ndarray]) -> np.ndarray:
    """predict.py:493, 612: mean over checkpoints of sigmoid(raw scores)."""
    return np.mean(np.stack([sigmoid(x) for x in logits_per_ckpt], axis=0), axis=0).astype(F32)


# ------------------------------------------------------------------------------------------- fusion transformer (a-1)
def gmean_normalized_ranks(members: Sequence[np.ndarray]) -> np.ndarray:
    """Geometric mean of the checkpoints' normalised-rank tensors (notebooks/generate_embeddings.ipynb cell 18:
    `gmean(np.stack([...], axis=-1), axis=-1)` with scipy.stats.mstats.gmean = exp(mean(log a)) in the input's
    float32; zeros (the diagonal) give log = -inf and a result of 0)."""
    a = np.stack([np.asarray(m, F32) for m in members], axis=-1)
    with np.errstate(divide="ignore"):
        log_a = np.log(a)
    return np.exp(log_a.mean(axis=-1)).astype(F32)


def ensemble_normalized_ranks(members: Sequence[np.ndarray], kind: Optional[str] = None) -> np.ndarray:
    """ipynb cells 18 + 20: gmean of the normalised ranks, then the same run_slice re-normalisation."""
    return normalize_scores(gmean_normalized_ranks(members), kind=kind)


def ilog_table(Q: int, scale: int = 2048) -> np.ndarray:
    """Fixed-point log2 table of the quantile-formulation ensemble: round(scale * (log2(max(r, 1/2)) + 1)), r = 0..Q."""
    r = np.arange(Q + 1, dtype=np.float64)
    return np.round(scale * (np.log2(np.maximum(r, 0.5)) + 1.0)).astype(np.uint16)


def ensemble_quantile_ranks(members: Sequence[np.ndarray], ilog: np.ndarray, thresholds: np.ndarray) -> np.ndarray:
    """Ensemble of fused uint16 ranks in the quantile formulation (generate_embeddings.ipynb:434, 634-649 restated on
    quantile ranks): g = sum_k ilog[r_k] (the gmean's log-sum in fixed point), rank = searchsorted(thresholds[l],
    float32(g), 'right'), 0 where every member rank is 0 (the diagonal).  members: K uint16 [L, ...]."""
    g = np.zeros(members[0].shape, dtype=np.int64)
    anyr = np.zeros(members[0].shape, dtype=bool)
    for m in members:
        g += ilog[np.minimum(m.astype(np.int64), len(ilog) - 1)].astype(np.int64)
        anyr |= m != 0
    out = quantile_rank(thresholds.astype(F32), g.astype(F32), "right")
    out[~anyr] = 0
    return out


def gather_triples(scores: np.ndarray, labels: np.ndarray, heads: np.ndarray, tails: np.ndarray) -> np.ndarray:
    """`pred[ddi_labels, head_idx, tail_idx]` (train_ddi_batch.py:286, evaluate.py:195) on a dense [L, Nh, Nt] array."""
    return scores[labels, heads, tails]


def _layer_norm(x, w, b, eps=1e-5):
    """F.layer_norm over the last dim (biased variance, eps 1e-5 = nn.LayerNorm default; models.py:366, 372-373)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def _linear(x, w, b):
    return x @ w.T + b


def _activation(x, actn):
    if actn == "relu":
        return np.maximum(x, 0)
    if actn == "gelu":  # exact erf GELU (F.gelu default, selected by transformer_actn: gelu)
        return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))
    raise NotImplementedError(actn)


def _mha(q_in, kv_in, in_w, in_b, out_w, out_b, num_heads, add_mask):
    """nn.MultiheadAttention forward (packed in_proj, q scaled by 1/sqrt(head_dim), additive -inf mask, softmax
    over keys, out_proj).  q_in [B,Tq,Dl], kv_in [B,Tk,Dl], add_mask broadcastable to [B,1,Tq,Tk] (0 / -inf)."""
    B, Tq, Dl = q_in.shape
    Tk = kv_in.shape[1]
    hd = Dl // num_heads
    q = _linear(q_in, in_w[:Dl], in_b[:Dl]).reshape(B, Tq, num_heads, hd).transpose(0, 2, 1, 3)
    k = _linear(kv_in, in_w[Dl:2 * Dl], in_b[Dl:2 * Dl]).reshape(B, Tk, num_heads, hd).transpose(0, 2, 1, 3)
    v = _linear(kv_in, in_w[2 * Dl:], in_b[2 * Dl:]).reshape(B, Tk, num_heads, hd).transpose(0, 2, 1, 3)
    s = (q * q.dtype.type(1.0 / math.sqrt(hd))) @ k.transpose(0, 1, 3, 2) + add_mask
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p = p / p.sum(axis=-1, keepdims=True)
    o = (p @ v).transpose(0, 2, 1, 3).reshape(B, Tq, Dl)
    return _linear(o, out_w, out_b)


def fusion_forward(sd: Dict[str, np.ndarray], cfg: dict, fusion_sequence: np.ndarray, fusion_mask: np.ndarray,
                   src_mask: Optional[np.ndarray] = None, pool_key_mask: Optional[np.ndarray] = None,
                   dtype=F32) -> np.ndarray:
    """TransformerFusion.forward in eval mode, models.py:401-455 (dropouts = identity).

    sd: the module's state_dict as numpy arrays (keys as in the reference: embed2latent.*, transformer_encoder.
    layers.{i}.*, latent2embed.*, x_attn_*).  cfg keys: num_layers, num_heads, head_dim, ffn_dim, actn, norm_first,
    agg.  fusion_sequence [B,T,E] (already position-encoded), fusion_mask [B,T] bool True = missing,
    src_mask [T,T] bool True = blocked, pool_key_mask [T] bool = the module's constant x_attn_key_padding_mask
    (models.py:382-385).  Encoder layers follow torch 1.13 nn.TransformerEncoderLayer._sa_block/_ff_block with
    norm_first either way, no final norm (models.py:366-367).  The result does not depend on transformer_batch_first
    except for the reference's x-attn indexing bug at models.py:443 (only batch_first=False is meaningful there).
    """
    g = lambda k: sd[k].astype(dtype)
    x = fusion_sequence.astype(dtype)
    B, T, _ = x.shape
    H = cfg["num_heads"]
    neg = np.zeros((B, 1, 1, T), dtype=dtype)
    neg[np.asarray(fusion_mask, bool)[:, None, None, :]] = -np.inf
    if src_mask is not None:
        sm = np.zeros((1, 1, T, T), dtype=dtype)
        sm[np.asarray(src_mask, bool)[None, None]] = -np.inf
        neg = neg + sm
    h = _linear(x, g("embed2latent.weight"), g("embed2latent.bias"))  # models.py:411
    for i in range(cfg["num_layers"]):  # models.py:412
        p = f"transformer_encoder.layers.{i}."
        sa = lambda y: _mha(y, y, g(p + "self_attn.in_proj_weight"), g(p + "self_attn.in_proj_bias"),
                            g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias"), H, neg)
        ff = lambda y: _linear(_activation(_linear(y, g(p + "linear1.weight"), g(p + "linear1.bias")), cfg["actn"]),
                               g(p + "linear2.weight"), g(p + "linear2.bias"))
        n1 = lambda y: _layer_norm(y, g(p + "norm1.weight"), g(p + "norm1.bias"))
        n2 = lambda y: _layer_norm(y, g(p + "norm2.weight"), g(p + "norm2.bias"))
        if cfg["norm_first"]:
            h = h + sa(n1(h))
            h = h + ff(n2(h))
        else:
            h = n1(h + sa(h))
            h = n2(h + ff(h))
    agg = cfg["agg"]
    l2e = lambda y: _linear(y, g("latent2embed.weight"), g("latent2embed.bias"))
    if agg == "cls":  # models.py:415-421
        return l2e(h)[:, 0, :].astype(dtype)
    if agg == "x-attn":  # models.py:422-443
        kv = _layer_norm(h, g("x_attn_kv_norm.weight"), g("x_attn_kv_norm.bias"))
        q = np.broadcast_to(g("x_attn_query")[None, :, :], (B, 1, h.shape[-1]))
        if cfg["norm_first"]:
            q = _layer_norm(q, g("x_attn_query_norm.weight"), g("x_attn_query_norm.bias"))
        pm = np.zeros((1, 1, 1, T), dtype=dtype)
        if pool_key_mask is not None:
            pm[:, :, :, np.asarray(pool_key_mask, bool)] = -np.inf
        out = _mha(q, kv, g("x_attn_mha_layer.in_proj_weight"), g("x_attn_mha_layer.in_proj_bias"),
                   g("x_attn_mha_layer.out_proj.weight"), g("x_attn_mha_layer.out_proj.bias"), H, pm)
        out = out + q
        if not cfg["norm_first"]:
            out = _layer_norm(out, g("x_attn_query_norm.weight"), g("x_attn_query_norm.bias"))
        return l2e(out)[:, 0, :].astype(dtype)
    e = l2e(h)  # models.py:415
    keep = ~np.asarray(fusion_mask, bool)
    if agg == "mean":  # models.py:444-447: scatter_mean over unmasked tokens
        cnt = np.maximum(keep.sum(axis=1, keepdims=True), 1).astype(dtype)
        return ((e * keep[:, :, None]).sum(axis=1) / cnt).astype(dtype)
    if agg == "max":  # models.py:448-451: scatter_max over unmasked tokens (empty -> 0)
        m = np.where(keep[:, :, None], e, -np.inf).max(axis=1)
        return np.where(np.isinf(m), 0, m).astype(dtype)
    raise NotImplementedError(agg)


# ------------------------------------------------------------------------------------------ positional encodings (a-2)
def sinusoidal_pe(d_model: int, max_len: int, seq_len: int) -> np.ndarray:
    """PositionEncodingSinusoidal buffer, models.py:560-577: [1, seq_len, d_model], zero beyond max_len."""
    position = np.arange(max_len, dtype=F32)[:, None]
    div_term = np.exp(np.arange(0, d_model, 2, dtype=F32) * F32(-math.log(10000.0) / d_model)).astype(F32)
    pe = np.zeros((1, seq_len, d_model), dtype=F32)
    pe[0, :max_len, 0::2] = np.sin(position * div_term)
    pe[0, :max_len, 1::2] = np.cos(position * div_term)
    return pe


# ------------------------------------------------------------------------------------------- unimodal MLP bypass (a-3)
def mlp_adaptor(layers: Sequence[dict], x: np.ndarray, dtype=F32) -> np.ndarray:
    """MLPAdaptor.forward (eval), models.py:459-518, as a list of ops in nn.Sequential order.

    Each entry: {'op': 'linear', 'w', 'b'} | {'op': 'ln', 'w', 'b'} | {'op': 'bn', 'w', 'b', 'mean', 'var'} (eval-mode
    nn.BatchNorm1d, norm='bn' at models.py:154,492: (h - running_mean) / sqrt(running_var + 1e-5) * w + b) |
    {'op': 'act', 'actn'}.
    """
    h = x.astype(dtype)
    for L in layers:
        if L["op"] == "linear":
            h = _linear(h, L["w"].astype(dtype), L["b"].astype(dtype))
        elif L["op"] == "ln":
            h = _layer_norm(h, L["w"].astype(dtype), L["b"].astype(dtype))
        elif L["op"] == "bn":
            h = ((h - L["mean"].astype(dtype)) / np.sqrt(L["var"].astype(dtype) + dtype(1e-5)) * L["w"].astype(dtype)
                 + L["b"].astype(dtype)).astype(dtype)
        elif L["op"] == "act":
            h = _activation(h, L["actn"])
        else:
            raise NotImplementedError(L["op"])
    return h.astype(dtype)


# ------------------------------------------------------------------------------ chemCPA transcriptomic token (f-4)
def chemcpa_mlp(sd: dict, prefix: str, x: np.ndarray, dtype=F32) -> np.ndarray:
    """chemCPA `MLP.forward` in eval mode with last_layer_act='linear' (chemcpa/chemCPA/model.py:161-231), driven by
    the module's own state_dict: entries `{prefix}network.<name>.*` in registration order.  A 2-D `weight` is a
    Linear; `running_mean` marks a BatchNorm1d (eval: (x - mean) / sqrt(var + 1e-5) * weight + bias); a ReLU follows
    every Linear (after its BatchNorm1d when present) except the last Linear of the chain (model.py:178-187,
    199-218)."""
    names = []
    for k in sd:
        if k.startswith(prefix + "network."):
            n = k[len(prefix + "network."):].rsplit(".", 1)[0]
            if n not in names:
                names.append(n)
    get = lambda n, f: np.asarray(sd[f"{prefix}network.{n}.{f}"]).astype(dtype)
    linears = [n for n in names if np.asarray(sd[f"{prefix}network.{n}.weight"]).ndim == 2]
    h = x.astype(dtype)
    for pos, n in enumerate(names):
        if n in linears:
            h = _linear(h, get(n, "weight"), get(n, "bias"))
            nxt = names[pos + 1] if pos + 1 < len(names) else None
            has_bn = nxt is not None and f"{prefix}network.{nxt}.running_mean" in sd
            if not has_bn and n != linears[-1]:
                h = np.maximum(h, 0)
        else:  # BatchNorm1d, always followed by a ReLU
            h = (h - get(n, "running_mean")) / np.sqrt(get(n, "running_var") + dtype(1e-5)) * get(n, "weight") \
                + get(n, "bias")
            h = np.maximum(h, 0)
    return h.astype(dtype)


def chemcpa_tx_latents(sd: dict, genes: np.ndarray, cov_idx: Sequence[np.ndarray], *, use_drugs: bool,
                       doser_type: Optional[str] = None, drug_table: Optional[np.ndarray] = None,
                       drugs_idx: Optional[np.ndarray] = None, dosages: Optional[np.ndarray] = None, dtype=F32):
    """`TxAdaptingComPert.predict` restricted to its latents (model.py:678-697) -> (latent_basal, latent_treated).
    Dose scale per compute_drug_embeddings_ (:601-653) with the dosers of :259-271 ('sigm'/'logsigm'), :622-627
    ('amortized'), :609-621 (per-drug 'mlp') or the dosage itself (nonlin None)."""
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    basal = chemcpa_mlp(sd, "encoder.", genes, dtype)
    treated = basal
    if use_drugs:
        emb = drug_table.astype(dtype)[drugs_idx]
        d = dosages.astype(dtype)
        if doser_type in ("sigm", "logsigm"):
            beta = np.asarray(sd["dosers.beta"]).astype(dtype)[0][drugs_idx]
            bias = np.asarray(sd["dosers.bias"]).astype(dtype)[0][drugs_idx]
            xx = np.log1p(d) if doser_type == "logsigm" else d
            scale = sig(xx * beta + bias) - sig(bias)
        elif doser_type == "amortized":
            scale = chemcpa_mlp(sd, "dosers.", np.concatenate([emb, d[:, None]], axis=1), dtype).reshape(-1)
        elif doser_type == "mlp":  # :609-621 — the drug's own MLP on the scalar dosage, then a sigmoid
            scale = np.asarray([sig(chemcpa_mlp(sd, f"dosers.{int(i)}.", np.asarray([[x]], dtype), dtype)[0, 0])
                                for i, x in zip(drugs_idx, d)], dtype)
        elif doser_type is None:
            scale = d
        else:
            raise NotImplementedError(doser_type)
        lat = chemcpa_mlp(sd, "drug_embedding_encoder.", emb, dtype)
        treated = treated + scale.astype(dtype)[:, None] * lat
    for c, idx in enumerate(cov_idx):
        treated = treated + np.asarray(sd[f"covariates_embeddings.{c}.weight"]).astype(dtype)[idx]
    return basal.astype(dtype), treated.astype(dtype)


# ----------------------------------------------------------------------------------------------- token assembly (a-2)
def assemble_fusion_inputs(all_embeds: np.ndarray, batch_masks: np.ndarray, *, n_non_tx: int, num_tx_bottlenecks: int,
                           agg: str, tx_bottleneck_tokens: Optional[np.ndarray] = None,
                           cls: Optional[np.ndarray] = None, pe: Optional[np.ndarray] = None,
                           pos_emb_type: str = "sinusoidal", pos_max_len: Optional[int] = None,
                           normalize: bool = False):
    """The fusion section of NovelDDIEncoder.encode for fusion='transformer', models.py:793-853.

    all_embeds [B, M, E] in the order [non-TX..., TX...] (models.py:772), batch_masks [B, M] bool True = missing.
    Returns (pos-encoded sequence [B,T,E], fusion mask [B,T], src_mask [T,T] or None).
    """
    seq = all_embeds.astype(F32)
    masks = np.asarray(batch_masks, bool)
    B, M, E = seq.shape
    n_tx = M - n_non_tx
    src_mask = None
    nb = num_tx_bottlenecks
    if nb > 0:  # models.py:799-816
        seq = np.concatenate([seq[:, :n_non_tx], np.broadcast_to(tx_bottleneck_tokens[None], (B, nb, E)),
                              seq[:, n_non_tx:]], axis=1)
        masks = np.concatenate([masks[:, :n_non_tx], np.zeros((B, nb), bool), masks[:, n_non_tx:]], axis=1)
        T = seq.shape[1]
        src_mask = np.zeros((T, T), bool)
        src_mask[:n_non_tx, T - n_tx:] = True
        src_mask[T - n_tx:, :n_non_tx] = True
    if agg == "cls":  # models.py:818-842
        seq = np.concatenate([np.broadcast_to(cls[None], (B, 1, E)), seq], axis=1)
        masks = np.concatenate([np.zeros((B, 1), bool), masks], axis=1)
        if src_mask is not None:
            T = src_mask.shape[0]
            sm = np.zeros((T + 1, T + 1), bool)
            sm[1:, 1:] = src_mask
            src_mask = sm
    if normalize:  # models.py:849-850, F.normalize(p=2, dim=-1, eps=1e-12)
        n = np.sqrt((seq ** 2).sum(axis=-1, keepdims=True))
        seq = seq / np.maximum(n, F32(1e-12))
    seq = seq.copy()
    if pe is not None:  # models.py:852
        if pos_emb_type == "sinusoidal":
            seq = seq + pe  # buffer already has the sequence length (models.py:571-579, 586)
        else:
            seq[:, :pos_max_len, :] += pe  # models.py:602
    return seq.astype(F32), masks, src_mask


def split_unimodal(batch_masks: np.ndarray):
    """fusion='transformer_uni_proj' routing, models.py:781-790: rows with exactly one visible modality bypass the
    transformer.  Returns (multimodal bool [B], index of the single visible modality for the unimodal rows)."""
    vis = ~np.asarray(batch_masks, bool)
    assert (vis.sum(axis=1) > 0).all()  # models.py:783
    multi = vis.sum(axis=1) > 1
    uni_idx = np.where(vis[~multi])[1]
    return multi, uni_idx
