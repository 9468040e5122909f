"""Score normalisation (reference: notebooks/normalize_scores.py:33-74).

Two formulations:
  * the fused quantile-table rank (`RankTable` + `pair_score(out='rank')`) — what the hot path uses;
  * the reference's exact in-sample rank (`exact_normalized_ranks`, mdg_exact_rank) for parity at sizes where the
    [L, N, N] logits are materialised.

`build_reference_quantiles` produces the per-outcome table rows: Q order statistics (ranks ceil(i*M/Q)) of the strict
lower triangle of each outcome's score matrix over a reference panel of drugs — the distribution the reference ranks
against (normalize_scores.py:67: entries with col >= row are excluded).  It is SETUP, run once per model/catalogue,
outside any timed region; it materialises one outcome chunk of fp32 logits at a time and takes the order statistics
with mdg_lower_triangle_quantiles (gather + CUB radix sort + pick).
"""
from typing import Optional

import torch

from . import _lib
from .decoder import RankTable, _require_cuda_f32, _stream_ptr, _workspace, ensemble_reduce, pair_score


def exact_normalized_ranks(scores: torch.Tensor) -> torch.Tensor:
    """Drop-in for `run_slice` over all outcomes (normalize_scores.py:62-74): [L, N, N] fp32 raw scores -> [L, N, N]
    fp32 symmetric in-sample normalised ranks with zero diagonal (mdg_exact_rank)."""
    s = _require_cuda_f32(scores, "scores")
    if s.dim() != 3 or s.shape[1] != s.shape[2]:
        raise ValueError("scores must be [L, N, N]")
    L, N, _ = s.shape
    out = torch.empty_like(s)
    fn = _lib.lib()
    ws = _workspace(s.device, fn.mdg_exact_rank_workspace_bytes(N))
    with torch.cuda.device(s.device):
        _lib.check(fn.mdg_exact_rank(s.data_ptr(), L, N, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                     _stream_ptr(s.device)), "mdg_exact_rank")
    return out


def classwise_normalized_rank_3d(tensor: torch.Tensor) -> torch.Tensor:
    """Alias with the reference function's name for scripts ported from notebooks/normalize_scores.py; applies the
    run_slice masking (col >= row excluded) because that is the only way the reference calls it."""
    return exact_normalized_ranks(tensor)


def lower_triangle_quantiles(scores: torch.Tensor, Q: int) -> torch.Tensor:
    """[L, N, N] fp32 -> [L, Q] ascending order statistics of each outcome's strict lower triangle."""
    s = _require_cuda_f32(scores, "scores")
    L, N, _ = s.shape
    out = torch.empty((L, Q), dtype=torch.float32, device=s.device)
    fn = _lib.lib()
    ws = _workspace(s.device, fn.mdg_exact_rank_workspace_bytes(N))
    with torch.cuda.device(s.device):
        _lib.check(fn.mdg_lower_triangle_quantiles(s.data_ptr(), L, N, Q, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                   _stream_ptr(s.device)), "mdg_lower_triangle_quantiles")
    return out


def build_reference_quantiles(z: torch.Tensor, weight: torch.Tensor, Q: int, *, panel: Optional[int] = None,
                              precision: str = "bf16", chunk: int = 8, normalize: bool = False) -> torch.Tensor:
    """[L, Q] ascending fp32 quantiles of each outcome's strict-lower-triangle logits over the first `panel` drugs
    (all drugs if None).  `precision` should match the mode the ranks will be computed in."""
    n = z.shape[0] if panel is None else min(panel, z.shape[0])
    zp = z[:n].contiguous()
    L = weight.shape[0]
    M = n * (n - 1) // 2
    if M < 1:
        raise ValueError("need at least 2 drugs for a reference distribution")
    Q = min(Q, M)
    out = torch.empty((L, Q), dtype=torch.float32, device=z.device)
    for l0 in range(0, L, chunk):
        l1 = min(l0 + chunk, L)
        logits = pair_score(zp, zp, weight[l0:l1], precision=precision, out="logit", normalize=normalize)
        out[l0:l1] = lower_triangle_quantiles(logits, Q)
    return out


def build_rank_table(z: torch.Tensor, weight: torch.Tensor, Q: int = 16384, kind: str = "lut", **kw) -> RankTable:
    return RankTable(build_reference_quantiles(z, weight, Q, **kw), kind=kind)


def ranks_to_normalized(ranks_u16: torch.Tensor, Q: int) -> torch.Tensor:
    """uint16 quantile ranks -> the reference's (0, 1] normalised-rank scale (|error| <= 1/Q + snapping)."""
    return ranks_u16.to(torch.float32) / float(Q)


def gmean_normalized_ranks(members, Q: Optional[int] = None) -> torch.Tensor:
    """Geometric mean over checkpoints of their normalised-rank tensors (generate_embeddings.ipynb cell 18:
    scipy.stats.mstats.gmean over the stacked float32 tensors).  members: K fp32 [L, N, N] tensors, or K uint16
    fused-rank tensors with their table size Q (ranks are scaled by 1/Q first)."""
    members = list(members)
    if members[0].dtype == torch.uint16:
        if not Q:
            raise ValueError("uint16 rank members need Q")
        return ensemble_reduce(members, "gmean_rank", 1.0 / float(Q))
    return ensemble_reduce(members, "gmean")


def ensemble_normalized_ranks(members, Q: Optional[int] = None) -> torch.Tensor:
    """The reference's final ensemble tensor (ipynb cells 18 + 20): gmean of the checkpoints' normalised ranks,
    re-normalised with the same in-sample rank (`run_slice` again)."""
    return exact_normalized_ranks(gmean_normalized_ranks(members, Q))


# ------------------------------------------------------------------------------------------------------------------
# Ensemble in the quantile-table formulation (reference: generate_embeddings.ipynb:434, 634-649 — scipy gmean of the
# checkpoints' normalised ranks, then run_slice again).  Members are FUSED uint16 ranks; the geometric mean's log-sum is
# taken in fixed point through a table (which DEFINES the logarithm, so the CPU oracle reproduces every bit) and the
# re-normalisation is a second quantile lookup against an ensemble table built from a reference panel.
# ------------------------------------------------------------------------------------------------------------------
ILOG_SCALE = 2048


def ilog_table(Q: int, scale: int = ILOG_SCALE):
    """uint16 [Q + 1]: round(scale * (log2(max(r, 1/2)) + 1)) — rank 0 (below every threshold) counts as half a rank."""
    import numpy as np
    r = np.arange(Q + 1, dtype=np.float64)
    t = np.round(scale * (np.log2(np.maximum(r, 0.5)) + 1.0))
    if t.max() > 65535:
        raise ValueError("ilog table does not fit uint16: lower `scale`")
    return t.astype(np.uint16)


class EnsembleRankTable:
    """ilog table of the member ranks + the rank table the log-sum is looked up in (`mdg_ensemble_rank_u16`)."""

    def __init__(self, member_Q: int, quantiles: torch.Tensor, scale: int = ILOG_SCALE):
        self.member_Q, self.scale = member_Q, scale
        self.ilog_host = ilog_table(member_Q, scale)
        self.ilog = torch.from_numpy(self.ilog_host.view("int16")).to(quantiles.device)
        self.table = RankTable(quantiles)  # exact LUT over float32(g)
        self.Q = self.table.Q


def _member_ptrs(members):
    import ctypes
    members = list(members)
    first = members[0]
    if not 1 <= len(members) <= _lib.MDG_MAX_ENSEMBLE:
        raise ValueError(f"need 1..{_lib.MDG_MAX_ENSEMBLE} members")
    ms = []
    for m in members:
        if not m.is_cuda or m.dtype != torch.uint16 or m.shape != first.shape or m.device != first.device:
            raise RuntimeError("madrigal_b200: ensemble members must be same-shape uint16 CUDA tensors on one device")
        ms.append(m.contiguous())
    return ms, (ctypes.c_void_p * len(ms))(*[m.data_ptr() for m in ms])


def ensemble_logsum(members, member_Q: int, scale: int = ILOG_SCALE) -> torch.Tensor:
    """float32(sum_k ilog[r_k]) of K uint16 rank tensors [L, ...] (builder mode of mdg_ensemble_rank_u16)."""
    ms, ptrs = _member_ptrs(members)
    first = ms[0]
    L = first.shape[0]
    n = first[0].numel() if L > 0 else 0
    ilog = torch.from_numpy(ilog_table(member_Q, scale).view("int16")).to(first.device)
    out = torch.empty(first.shape, dtype=torch.float32, device=first.device)
    with torch.cuda.device(first.device):
        _lib.check(_lib.lib().mdg_ensemble_rank_u16(ptrs, len(ms), L, n, ilog.data_ptr(), member_Q, None, None,
                                                    out.data_ptr(), _stream_ptr(first.device)), "mdg_ensemble_rank_u16")
    return out


def build_ensemble_rank_table(panel_members, member_Q: int, Q: int = 16384, scale: int = ILOG_SCALE) -> EnsembleRankTable:
    """Ensemble table from the members' fused ranks over a reference panel: K uint16 [L, n, n] tensors in the
    normaliser layout -> Q order statistics of the strict-lower-triangle log-sums per outcome."""
    g = ensemble_logsum(panel_members, member_Q, scale)
    n = g.shape[1]
    Q = min(Q, n * (n - 1) // 2)
    return EnsembleRankTable(member_Q, lower_triangle_quantiles(g, Q), scale)


def ensemble_fused_ranks(members, table: EnsembleRankTable, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint16 ensemble ranks of K fused-rank tensors [L, ...]: searchsorted(table thresholds, float32(log-sum), 'right'),
    0 where every member is 0 (the diagonal).  `table.table.thresholds` is what the CPU oracle is run against."""
    import ctypes
    ms, ptrs = _member_ptrs(members)
    first = ms[0]
    L = first.shape[0]
    if L > table.table.L:
        raise ValueError("more outcomes than the ensemble table holds")
    n = first[0].numel() if L > 0 else 0
    if out is None:
        out = torch.empty(first.shape, dtype=torch.uint16, device=first.device)
    st = table.table.struct(0, L)
    with torch.cuda.device(first.device):
        _lib.check(_lib.lib().mdg_ensemble_rank_u16(ptrs, len(ms), L, n, table.ilog.data_ptr(), table.member_Q,
                                                    ctypes.byref(st), out.data_ptr(), None, _stream_ptr(first.device)),
                   "mdg_ensemble_rank_u16")
    return out
