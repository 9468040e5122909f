"""Bilinear decoder drop-ins (reference: madrigal/models/models.py:521-547) over the C ABI.

`BilinearDDIScorer` keeps the reference's constructor, parameter names (`weight [L, D, D]`, unused `bias [L]`) and
`forward(input1, input2, label_range=None) -> [L', N1, N2]` contract, so a reference `state_dict` loads unchanged and
`register_parametrization(decoder, 'weight', Symmetric())` works as in models.py:922.  The arithmetic runs in
`mdg_pair_score` (tcgen05 GEMMs); there is no PyTorch fallback.
"""
import ctypes
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import MdgRankTable

_PRECISION = {"bf16": _lib.MDG_PREC_BF16, "fp32": _lib.MDG_PREC_FP32}
_OUT = {"logit": (_lib.MDG_OUT_LOGIT_F32, torch.float32), "sigmoid": (_lib.MDG_OUT_SIGMOID_F32, torch.float32),
        "rank": (_lib.MDG_OUT_RANK_U16, torch.uint16)}

_workspaces = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer (operand copies in bf16, the intermediate z.W_l, the tile scheduler's counter), one per
    (device, CUDA stream): calls enqueued on different streams of a device never share operands or task counters, and
    a buffer is only ever used on the stream it was allocated on, so the caching allocator's stream-ordered reuse of
    a replaced (smaller) block is safe."""
    stream = torch.cuda.current_stream(device)
    key = (device.type, device.index, stream.cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _workspaces.pop(key, None)
        with torch.cuda.device(device):
            ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def l2_normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """F.normalize(x, p=2, dim=-1) (eps 1e-12) on the last dimension through mdg_l2_normalize_rows."""
    x2 = _require_cuda_f32(x, "x")
    out = torch.empty_like(x2)
    rows = x2.numel() // max(x2.shape[-1], 1)
    with torch.cuda.device(x2.device):
        _lib.check(_lib.lib().mdg_l2_normalize_rows(x2.data_ptr(), rows, x2.shape[-1], out.data_ptr(),
                                                    _stream_ptr(x2.device)), "mdg_l2_normalize_rows")
    return out


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"madrigal_b200: `{name}` must be a CUDA tensor (no CPU path exists)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"madrigal_b200: `{name}` must be float32, got {t.dtype}")
    return t.contiguous()


class RankTable:
    """Per-outcome reference-quantile table prepared for the fused rank epilogue (mdg_rank_table_build).

    `thresholds[l]` is what `np.searchsorted(..., side='right')` must be run against to reproduce the ranks.
    """

    def __init__(self, quantiles: torch.Tensor, kind: str = "lut"):
        """kind='lut': exact bucket LUT (thresholds move by <= ~3 grid cells).  kind='pwl': 256-bin histogram CDF with
        linear interpolation, conflict-free in shared memory (faster epilogue); `max_rank_deviation[l]` then reports how
        many ranks the table is away from the supplied quantiles at worst."""
        q = _require_cuda_f32(quantiles, "quantiles")
        if q.dim() != 2:
            raise ValueError("quantiles must be [L, Q]")
        if kind not in _lib.MDG_RANK_KIND:
            raise ValueError(f"kind={kind!r} (supported: {sorted(_lib.MDG_RANK_KIND)})")
        L, Q = q.shape
        if Q > _lib.MDG_RANK_MAX_Q:
            raise ValueError(f"Q={Q} exceeds {_lib.MDG_RANK_MAX_Q}")
        self.L, self.Q, self.kind = L, Q, kind
        self.thresholds = torch.empty_like(q)
        self.lut = torch.empty((L, _lib.MDG_RANK_LUT_ENTRIES), dtype=torch.int32, device=q.device)
        self.affine = torch.empty((L, 2), dtype=torch.float32, device=q.device)
        self.max_rank_deviation = None
        with torch.cuda.device(q.device):
            if kind == "pwl":
                self.max_rank_deviation = torch.empty((L,), dtype=torch.float32, device=q.device)
                _lib.check(_lib.lib().mdg_rank_table_build_pwl(q.data_ptr(), L, Q, self.thresholds.data_ptr(),
                                                               self.lut.data_ptr(), self.affine.data_ptr(),
                                                               self.max_rank_deviation.data_ptr(),
                                                               _stream_ptr(q.device)), "mdg_rank_table_build_pwl")
            else:
                _lib.check(_lib.lib().mdg_rank_table_build(q.data_ptr(), L, Q, self.thresholds.data_ptr(),
                                                           self.lut.data_ptr(), self.affine.data_ptr(),
                                                           _stream_ptr(q.device)), "mdg_rank_table_build")

    def struct(self, l0: int = 0, l1: Optional[int] = None) -> MdgRankTable:
        l1 = self.L if l1 is None else l1
        return MdgRankTable(self.thresholds[l0:l1].data_ptr(), self.lut[l0:l1].data_ptr(),
                            self.affine[l0:l1].data_ptr(), l1 - l0, self.Q, _lib.MDG_RANK_KIND[self.kind])

    def lookup(self, logits: torch.Tensor) -> torch.Tensor:
        """ranks[l, ...] = searchsorted(thresholds[l], logits[l, ...], 'right') for materialised logits."""
        x = _require_cuda_f32(logits, "logits")
        if x.shape[0] != self.L:
            raise ValueError("logits.shape[0] must equal the number of outcomes in the table")
        out = torch.empty(x.shape, dtype=torch.uint16, device=x.device)
        st = self.struct()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mdg_rank_lookup(x.data_ptr(), x[0].numel(), ctypes.byref(st), out.data_ptr(),
                                                  _stream_ptr(x.device)), "mdg_rank_lookup")
        return out


def packed_tiles_per_outcome(N: int) -> int:
    """T = nb (nb + 1) / 2 with nb = ceil(N / 32): 32 x 32 tiles of the lower triangle incl. the diagonal tiles."""
    nb = (N + 31) // 32
    return nb * (nb + 1) // 2


def unpack_packed_tiles(packed, N: int):
    """Packed lower-triangular rank tiles [L, T, 32, 32] (MDG_PAIRS_PACKED_TILES) -> the normaliser's layout [L, N, N]
    (each rank at [i, j] and [j, i], zero diagonal; normalize_scores.py:67-70).  Works on a torch tensor (any device)
    or a numpy array; pure data movement (scatter of tiles + transpose), the host-side mirror of the packed copy."""
    import numpy as np
    is_np = isinstance(packed, np.ndarray)
    x = torch.from_numpy(packed.view(np.int16)) if is_np else packed.view(torch.int16)
    L, T = x.shape[0], x.shape[1]
    nb = (N + 31) // 32
    if T != nb * (nb + 1) // 2:
        raise ValueError("tile count does not match N")
    full = torch.zeros((L, nb * 32, nb * 32), dtype=torch.int16, device=x.device)
    blocks = full.view(L, nb, 32, nb, 32)
    t = 0
    for bi in range(nb):  # tile row bi holds tiles (bi, 0..bi) contiguously
        row = x[:, t:t + bi + 1]                      # [L, bi+1, 32, 32]
        blocks[:, bi, :, :bi + 1, :] = row.permute(0, 2, 1, 3)
        t += bi + 1
    full = torch.tril(full, -1)                       # diagonal tiles: keep col < row only
    full = (full + full.transpose(1, 2))[:, :N, :N].contiguous()
    if is_np:
        return full.numpy().view(np.uint16)
    return full.view(torch.uint16)


def mirror_packed_tiles_host(packed: torch.Tensor, N: int, out: Optional[torch.Tensor] = None,
                             threads: int = 0) -> torch.Tensor:
    """Host half of a packed device-to-host transfer (mdg_host_mirror_tiles): packed rank tiles that already sit in
    HOST memory ([L, T, 32, 32] uint16) -> the normaliser's layout [L, N, N] (rank at [i, j] and [j, i], zero diagonal;
    normalize_scores.py:67-70) on `threads` host threads (0: all).  Pure data movement — the ranks were computed on the
    GPU; this only spares PCIe the mirror image.  Same result as `unpack_packed_tiles`, which is the slow pure-torch
    statement of the layout."""
    if packed.is_cuda or packed.dtype != torch.uint16 or packed.dim() != 4 or tuple(packed.shape[2:]) != (32, 32):
        raise ValueError("packed must be a host uint16 tensor [L, T, 32, 32]")
    L, T = packed.shape[0], packed.shape[1]
    if T != packed_tiles_per_outcome(N):
        raise ValueError("tile count does not match N")
    packed = packed.contiguous()
    if out is None:
        out = torch.empty((L, N, N), dtype=torch.uint16)
    if out.is_cuda or out.dtype != torch.uint16 or tuple(out.shape) != (L, N, N) or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous host uint16 tensor {(L, N, N)}")
    _lib.check(_lib.lib().mdg_host_mirror_tiles(packed.data_ptr(), L, N, out.data_ptr(), int(threads)),
               "mdg_host_mirror_tiles")
    return out


class PreparedDecoder:
    """The decoder weights [L, D, D] converted ONCE to the tensor-core operand form (mdg_pair_prepare).  The reference
    re-reads `decoder.weight` through the Symmetric parametrisation on every call (models.py:537-547, 922); a scoring
    driver that loops outcome chunks or steps over fixed weights passes `weight=PreparedDecoder(W, precision)` (or
    `BilinearDDIScorer.prepared()`) and skips the per-call conversion.  `pd[l0:l1]` selects a `label_range`."""

    def __init__(self, weight: torch.Tensor, precision: str = "bf16", _view=None):
        if _view is not None:
            self.buf, self.precision, self.D, self.L_total, self.l0, self.l1 = _view
            return
        W = _require_cuda_f32(weight, "weight")
        if W.dim() != 3 or W.shape[1] != W.shape[2]:
            raise ValueError("weight must be [L, D, D]")
        self.precision, self.D, self.L_total = precision, W.shape[1], W.shape[0]
        self.l0, self.l1 = 0, W.shape[0]
        prec = _PRECISION[precision]
        fn = _lib.lib()
        nbytes = fn.mdg_pair_prepared_bytes(self.D, self.L_total, prec)
        self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=W.device)
        with torch.cuda.device(W.device):
            _lib.check(fn.mdg_pair_prepare(W.data_ptr(), self.D, self.L_total, prec, self.buf.data_ptr(),
                                           self.buf.numel(), _stream_ptr(W.device)), "mdg_pair_prepare")

    @property
    def shape(self):
        return (self.l1 - self.l0, self.D, self.D)

    @property
    def device(self):
        return self.buf.device

    def __getitem__(self, idx):
        if not isinstance(idx, slice) or idx.step not in (None, 1):
            raise TypeError("PreparedDecoder supports contiguous outcome slices only")
        a, b, _ = idx.indices(self.l1 - self.l0)
        return PreparedDecoder(None, _view=(self.buf, self.precision, self.D, self.L_total, self.l0 + a,
                                            self.l0 + max(a, b)))


def pair_score(z_rows: torch.Tensor, z_cols: torch.Tensor, weight, *, precision: str = "fp32",
               out: str = "logit", table: Optional[RankTable] = None, table_offset: int = 0,
               normalize: bool = False, out_tensor: Optional[torch.Tensor] = None,
               symmetric: bool = False, packed: bool = False) -> torch.Tensor:
    """All-pairs bilinear scores  S[l,i,j] = z_rows[i] . W[l] . z_cols[j]  with a fused epilogue.

    out='logit' | 'sigmoid' -> float32 [L, Nr, Nc];  out='rank' -> uint16 quantile ranks against `table`
    (rows table_offset .. table_offset+L of the table).  symmetric=True (out='rank', z_rows is z_cols): compute only
    row > col and write each rank at [i,j] and [j,i] with a zero diagonal — the reference normaliser's layout
    (normalize_scores.py:67-70) at half the MMAs and look-ups.  packed=True (implies symmetric): the same ranks without
    the mirror image, as uint16 [L, T, 32, 32] lower-triangular tiles (MDG_PAIRS_PACKED_TILES; `unpack_packed_tiles`
    rebuilds [L, N, N]) — half the bytes to write and to copy to the host.  `weight`: fp32 [L, D, D] or a
    `PreparedDecoder` (then `precision` is the prepared one).
    """
    symmetric = symmetric or packed
    zr = _require_cuda_f32(z_rows, "z_rows")
    zc = zr if z_cols is z_rows else _require_cuda_f32(z_cols, "z_cols")
    prepared = weight if isinstance(weight, PreparedDecoder) else None
    if prepared is not None:
        precision = prepared.precision
        if prepared.device != zr.device:
            raise RuntimeError("madrigal_b200: prepared decoder and embeddings are on different devices")
        W = None
        wshape = prepared.shape
    else:
        W = _require_cuda_f32(weight, "weight")
        if W.dim() != 3:
            raise ValueError("expected z_rows [Nr,D], z_cols [Nc,D], weight [L,D,D]")
        wshape = tuple(W.shape)
    if symmetric and (out != "rank" or zc.data_ptr() != zr.data_ptr() or zc.shape != zr.shape):
        raise ValueError("symmetric=True needs out='rank' and z_cols is z_rows")
    if zr.dim() != 2 or zc.dim() != 2:
        raise ValueError("expected z_rows [Nr,D], z_cols [Nc,D], weight [L,D,D]")
    Nr, D = zr.shape
    Nc = zc.shape[0]
    L = wshape[0]
    if zc.shape[1] != D or wshape[1] != D or wshape[2] != D:
        raise ValueError(f"shape mismatch: z_rows {tuple(zr.shape)}, z_cols {tuple(zc.shape)}, weight {wshape}")
    mode, dtype = _OUT[out]
    prec = _PRECISION[precision]
    out_shape = (L, packed_tiles_per_outcome(Nr), 32, 32) if packed else (L, Nr, Nc)
    if out_tensor is None:
        out_tensor = torch.empty(out_shape, dtype=dtype, device=zr.device)
    else:
        if tuple(out_tensor.shape) != out_shape or out_tensor.dtype != dtype or not out_tensor.is_contiguous():
            raise ValueError(f"out_tensor must be a contiguous {dtype} tensor of shape {out_shape}")
    tbl = None
    if out == "rank":
        if table is None:
            raise ValueError("out='rank' needs a RankTable")
        tbl = table.struct(table_offset, table_offset + L)
    fn = _lib.lib()
    nbytes = fn.mdg_pair_score_workspace_bytes(Nr, Nc, D, L, prec)
    ws = _workspace(zr.device, nbytes)
    pairs = _lib.MDG_PAIRS_PACKED_TILES if packed else (_lib.MDG_PAIRS_SYMMETRIC if symmetric else _lib.MDG_PAIRS_FULL)
    tbl_ref = ctypes.byref(tbl) if tbl is not None else None
    with torch.cuda.device(zr.device):
        if prepared is not None:
            _lib.check(fn.mdg_pair_score_prepared(zr.data_ptr(), zc.data_ptr(), prepared.buf.data_ptr(), prepared.l0,
                                                  Nr, Nc, D, L, prec, mode, pairs, int(bool(normalize)), tbl_ref,
                                                  out_tensor.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  _stream_ptr(zr.device)), "mdg_pair_score_prepared")
        else:
            _lib.check(fn.mdg_pair_score(zr.data_ptr(), zc.data_ptr(), W.data_ptr(), Nr, Nc, D, L, prec, mode, pairs,
                                         int(bool(normalize)), tbl_ref, out_tensor.data_ptr(), ws.data_ptr(),
                                         ws.numel(), _stream_ptr(zr.device)), "mdg_pair_score")
    return out_tensor


def pair_score_gather(z_rows: torch.Tensor, z_cols: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor,
                      heads: torch.Tensor, tails: torch.Tensor, *, precision: str = "fp32", out: str = "logit",
                      normalize: bool = False) -> torch.Tensor:
    """Scores of the listed (label, head, tail) triples only:  out[t] = z_rows[heads[t]] . W[labels[t]] . z_cols[tails[t]].

    What the reference computes as `model(...)[ddi_labels, head_idx, tail_idx]` (train_ddi_batch.py:285-286,
    evaluate.py:191-195) after materialising the whole [L, Nh, Nt] tensor; here the N^2 GEMM never runs
    (mdg_pair_score_gather).  out='logit' | 'sigmoid'."""
    zr = _require_cuda_f32(z_rows, "z_rows")
    zc = zr if z_cols is z_rows else _require_cuda_f32(z_cols, "z_cols")
    W = _require_cuda_f32(weight, "weight")
    if out not in ("logit", "sigmoid"):
        raise ValueError("out must be 'logit' or 'sigmoid'")
    idx = []
    for name, t in (("labels", labels), ("heads", heads), ("tails", tails)):
        if not t.is_cuda:
            raise RuntimeError(f"madrigal_b200: `{name}` must be a CUDA tensor (no CPU path exists)")
        idx.append(t.to(torch.int32).contiguous().reshape(-1))
    n = idx[0].numel()
    if idx[1].numel() != n or idx[2].numel() != n:
        raise ValueError("labels, heads and tails must have the same length")
    Nr, D = zr.shape
    Nc, L = zc.shape[0], W.shape[0]
    res = torch.empty((n,), dtype=torch.float32, device=zr.device)
    prec = _PRECISION[precision]
    fn = _lib.lib()
    ws = _workspace(zr.device, fn.mdg_pair_score_workspace_bytes(Nr, Nc, D, L, prec))
    with torch.cuda.device(zr.device):
        _lib.check(fn.mdg_pair_score_gather(zr.data_ptr(), zc.data_ptr(), W.data_ptr(), Nr, Nc, D, L, prec,
                                            int(bool(normalize)), idx[0].data_ptr(), idx[1].data_ptr(),
                                            idx[2].data_ptr(), n, _OUT[out][0], res.data_ptr(), ws.data_ptr(),
                                            ws.numel(), _stream_ptr(zr.device)), "mdg_pair_score_gather")
    return res


def ensemble_reduce(members, mode: str = "mean", rank_scale: float = 1.0) -> torch.Tensor:
    """mean / geometric mean over K same-shaped CUDA tensors (mdg_ensemble_reduce).  mode: 'mean' (fp32),
    'gmean' (fp32, scipy.stats.mstats.gmean semantics), 'gmean_rank' (uint16 quantile ranks * rank_scale)."""
    modes = {"mean": (_lib.MDG_ENS_MEAN_F32, torch.float32), "gmean": (_lib.MDG_ENS_GMEAN_F32, torch.float32),
             "gmean_rank": (_lib.MDG_ENS_GMEAN_RANK_U16, torch.uint16)}
    if mode not in modes:
        raise ValueError(f"mode={mode!r}")
    code, dtype = modes[mode]
    members = list(members)
    if not 1 <= len(members) <= _lib.MDG_MAX_ENSEMBLE:
        raise ValueError(f"need 1..{_lib.MDG_MAX_ENSEMBLE} members")
    first = members[0]
    ms = []
    for m in members:
        if not m.is_cuda or m.dtype != dtype or m.shape != first.shape or m.device != first.device:
            raise RuntimeError(f"madrigal_b200: ensemble members must be same-shape {dtype} CUDA tensors on one device")
        ms.append(m.contiguous())
    out = torch.empty(first.shape, dtype=torch.float32, device=first.device)
    ptrs = (ctypes.c_void_p * len(ms))(*[m.data_ptr() for m in ms])
    with torch.cuda.device(first.device):
        _lib.check(_lib.lib().mdg_ensemble_reduce(ptrs, len(ms), first.numel(), code, float(rank_scale),
                                                  out.data_ptr(), _stream_ptr(first.device)), "mdg_ensemble_reduce")
    return out


class Symmetric(nn.Module):
    """Parametrisation making each W_l exactly symmetric from its upper triangle (models.py:522-524)."""

    def forward(self, W: torch.Tensor) -> torch.Tensor:
        upper = torch.triu(W)
        return upper + torch.triu(W, diagonal=1).transpose(-1, -2)


class BilinearDDIScorer(nn.Bilinear):
    """Drop-in for the reference decoder (models.py:526-547): same ctor, same parameters, same forward contract.

    `precision` ('fp32' default = bf16x3 split, 1e-3 bar; 'bf16' = single bf16 term, 1e-2 bar) is an attribute so
    the reference's call sites (`model.decoder(z, z, (start, end))`, predict.py:428, 544) run unmodified.
    """

    def __init__(self, input_dim1: int, input_dim2: int, output_dim: int, precision: str = "fp32"):
        if input_dim1 != input_dim2:
            raise NotImplementedError("madrigal_b200 decoder needs input_dim1 == input_dim2 (as every Madrigal config)")
        super().__init__(in1_features=input_dim1, in2_features=input_dim2, out_features=output_dim)
        self.precision = precision

    def bilinear(self, input1: torch.Tensor, input2: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        return pair_score(input1, input2, weight, precision=self.precision, out="logit")

    def prepared(self) -> PreparedDecoder:
        """The (parametrised) weight in operand form, cached until the underlying parameter changes."""
        raw = self.parametrizations.weight.original if hasattr(self, "parametrizations") else self.weight
        key = (self.precision, raw.data_ptr(), raw._version, str(raw.device))
        cache = getattr(self, "_mdg_prepared", None)
        if cache is None or cache[0] != key:
            with torch.no_grad():
                cache = (key, PreparedDecoder(self.weight.detach(), self.precision))
            self._mdg_prepared = cache
        return cache[1]

    def forward(self, input1: torch.Tensor, input2: torch.Tensor, label_range: Optional[Tuple[int, int]] = None):
        weight = self.weight  # through the parametrisation, if one is registered
        if label_range is not None:
            assert len(label_range) == 2
            weight = weight[label_range[0]:label_range[1], :, :]
        return self.bilinear(input1, input2, weight)


def pair_topk(z_rows: torch.Tensor, z_cols: torch.Tensor, weight: torch.Tensor, thresholds: torch.Tensor, k: int, *,
              cap: Optional[int] = None, symmetric: bool = True, precision: str = "bf16", normalize: bool = False):
    """Per-outcome top-k scoring pairs without any dense [L, N, N] output (mdg_pair_topk).

    thresholds [L]: candidates are the scores >= thresholds[l] (e.g. `RankTable.thresholds[:, -1]`, the 1 - 1/Q
    quantile).  symmetric=True keeps unordered pairs of one catalogue (row > col) and skips tiles above the diagonal.
    Returns (scores [L,k] fp32 descending, rows [L,k] int32, cols [L,k] int32, status [L] int32: 0 ok / 1 fewer than k
    candidates / 2 candidate list overflowed `cap`).
    """
    zr = _require_cuda_f32(z_rows, "z_rows")
    zc = _require_cuda_f32(z_cols, "z_cols")
    W = _require_cuda_f32(weight, "weight")
    thr = _require_cuda_f32(thresholds, "thresholds")
    Nr, D = zr.shape
    Nc, L = zc.shape[0], W.shape[0]
    if thr.numel() != L:
        raise ValueError("thresholds must have one entry per outcome")
    cap = int(cap) if cap is not None else max(4 * k, 4096)
    dev = zr.device
    scores = torch.empty((L, k), dtype=torch.float32, device=dev)
    rows = torch.empty((L, k), dtype=torch.int32, device=dev)
    cols = torch.empty((L, k), dtype=torch.int32, device=dev)
    status = torch.empty((L,), dtype=torch.int32, device=dev)
    prec = _PRECISION[precision]
    fn = _lib.lib()
    ws = _workspace(dev, fn.mdg_pair_topk_workspace_bytes(Nr, Nc, D, L, prec, cap))
    with torch.cuda.device(dev):
        _lib.check(fn.mdg_pair_topk(zr.data_ptr(), zc.data_ptr(), W.data_ptr(), Nr, Nc, D, L, prec,
                                    _lib.MDG_PAIRS_SYMMETRIC if symmetric else _lib.MDG_PAIRS_FULL, int(bool(normalize)),
                                    thr.data_ptr(), k, cap, scores.data_ptr(), rows.data_ptr(), cols.data_ptr(),
                                    status.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "mdg_pair_topk")
    return scores, rows, cols, status
