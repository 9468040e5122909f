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
