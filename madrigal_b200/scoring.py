"""All-pairs scoring drivers (reference: madrigal/evaluate/predict.py:381-463, 502-579 and the catalogue loop in
notebooks/generate_embeddings.ipynb cell 10) and the multi-GPU partition of SURVEY.md §8e.

The reference loops outcomes in chunks of 10 (`label_range`), runs the decoder on the GPU, copies every chunk to the
host and writes it into an np.memmap (predict.py:420-429).  `score_all_pairs_to_host` keeps that call pattern —
chunked outcomes, host destination buffer — but the chunk is produced by one fused kernel (logits never exist in HBM
when ranks are requested) and the device->host copy of chunk c overlaps the compute of chunk c+1 on a second stream.
File I/O itself (memmap/.npy naming) is out of scope.
"""
from typing import Optional, Tuple

import torch

from .decoder import RankTable, pair_score, pair_topk

_OUT_DTYPE = {"logit": torch.float32, "sigmoid": torch.float32, "rank": torch.uint16}


def outcome_shard(L: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of L outcomes over `world_size` ranks (sizes differ by at most 1)."""
    base, rem = divmod(L, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def row_shard(N: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous partition of N drugs over ranks for the encoder stage."""
    return outcome_shard(N, rank, world_size)


def encoder_is_replicated(n_drugs: int, tokens_per_drug: int, world_size: int, device=None) -> bool:
    """Multi-GPU plan of the encoder stage.  The fused encoder works on 128-token-row tiles, one per SM:
    a catalogue that fits ONE wave (tiles <= SMs) takes one tile latency whether a rank encodes all of it or only its row
    shard, so every rank encodes the whole catalogue and the exchange of z disappears from the step; larger catalogues
    are row-sharded and replicated by the one exchange (`PeerAllGather` / `all_gather_embeddings`)."""
    if world_size <= 1:
        return True
    sms = 148
    if device is not None and torch.device(device).type == "cuda":
        sms = torch.cuda.get_device_properties(device).multi_processor_count
    drugs_per_tile = max(1, 128 // max(1, tokens_per_drug))
    return (n_drugs + drugs_per_tile - 1) // drugs_per_tile <= sms


def all_gather_embeddings(z_shard: torch.Tensor, N: int, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The path's ONLY collective: replicate the fused-embedding table [N, D] from per-rank row shards.

    Shards are the `row_shard` partition.  When N divides evenly the shards land straight in the final [N, D] buffer
    (`out`, reusable across steps) with one `all_gather_into_tensor` — no padding, no concatenation; otherwise sizes
    differ by one row and the shards are padded to equal length first (NCCL over NVLink on GPUs; gloo in the CPU tests).
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    D = z_shard.shape[1]
    if N % world == 0:
        if z_shard.shape[0] != N // world:
            raise ValueError("z_shard is not this rank's row_shard of N")
        if out is None:
            out = torch.empty((N, D), dtype=z_shard.dtype, device=z_shard.device)
        dist.all_gather_into_tensor(out, z_shard.contiguous(), group=group)
        return out
    per = -(-N // world)
    padded = torch.zeros((per, D), dtype=z_shard.dtype, device=z_shard.device)
    padded[: z_shard.shape[0]] = z_shard
    gathered = torch.empty((world * per, D), dtype=z_shard.dtype, device=z_shard.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = []
    for r in range(world):
        s, e = row_shard(N, r, world)
        parts.append(gathered[r * per: r * per + (e - s)])
    return torch.cat(parts, dim=0)


class PeerAllGather:
    """Replicates the fused-embedding table with the library's own NVLink push kernel (mdg_peer_allgather) instead of
    a collective library call: each rank stores its row shard straight into every rank's copy of the table through
    peer-mapped pointers and the ranks exchange one epoch flag, all in ONE kernel launch on the caller's stream.

    torch's symmetric-memory allocator is used for the plumbing only (allocate + exchange peer mappings of one buffer
    holding two alternating tables and the flag words).  Where peer mappings are unavailable (CPU / gloo tests, no P2P
    between the devices) `mode` is 'collective' and `gather` falls back to `all_gather_embeddings` (NCCL / gloo).
    """

    def __init__(self, N: int, D: int, device: torch.device, group=None):
        import ctypes
        import torch.distributed as dist
        self.N, self.D, self.group = N, D, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.r0, self.r1 = row_shard(N, self.rank, self.world)
        self.epoch = 0
        self.mode, self.reason = "collective", ""
        self._fallback = None
        if device.type != "cuda" or self.world > 8 or D % 4 != 0:
            self.reason = "needs CUDA devices, <= 8 ranks and D % 4 == 0"
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem
            table = N * D                                    # floats per table
            total = 2 * table + 64                           # two alternating tables + 64 flag words
            self._buf = symm_mem.empty(total, dtype=torch.float32, device=device)
            self._buf.zero_()
            pg = group if group is not None else dist.group.WORLD
            self._hdl = symm_mem.rendezvous(self._buf, pg)
            ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            if len(ptrs) != self.world or ptrs[self.rank] != self._buf.data_ptr():
                raise RuntimeError("unexpected peer pointer table")
            self._tables = [(ctypes.c_void_p * self.world)(*[p + 4 * k * table for p in ptrs]) for k in (0, 1)]
            self._flags = (ctypes.c_void_p * self.world)(*[p + 4 * 2 * table for p in ptrs])
            self._local = [self._buf[k * table:(k + 1) * table].view(N, D) for k in (0, 1)]
            self._hdl.barrier()                               # every rank's flags are zero before anyone signals
            self.mode = "peer"
        except Exception as e:  # no symmetric memory / no P2P: use the collective
            self.reason = f"{type(e).__name__}: {e}"

    def gather(self, z_shard: torch.Tensor) -> torch.Tensor:
        """[rows of this rank, D] -> the full table [N, D] (valid until the call after next)."""
        if self.mode != "peer":
            if self._fallback is None and self.N % self.world == 0:
                self._fallback = torch.empty((self.N, self.D), dtype=z_shard.dtype, device=z_shard.device)
            return all_gather_embeddings(z_shard, self.N, self.group, out=self._fallback)
        from . import _lib
        from .decoder import _require_cuda_f32, _stream_ptr
        z = _require_cuda_f32(z_shard, "z_shard")
        if tuple(z.shape) != (self.r1 - self.r0, self.D):
            raise ValueError("z_shard is not this rank's row_shard of the table")
        self.epoch += 1
        k = self.epoch & 1
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().mdg_peer_allgather(z.data_ptr(), z.shape[0], self.r0, self.D, self._tables[k],
                                                     self._flags, self.world, self.rank, self.epoch,
                                                     _stream_ptr(z.device)), "mdg_peer_allgather")
        return self._local[k]


def score_all_pairs(z: torch.Tensor, weight: torch.Tensor, *, out: str = "rank", table: Optional[RankTable] = None,
                    precision: str = "bf16", label_range: Optional[Tuple[int, int]] = None,
                    normalize: bool = False, out_tensor: Optional[torch.Tensor] = None,
                    symmetric: bool = False) -> torch.Tensor:
    """Device-resident all-pairs scores for outcomes `label_range` of `weight` (the decoder call of predict.py:428)."""
    l0, l1 = (0, weight.shape[0]) if label_range is None else label_range
    return pair_score(z, z, weight[l0:l1], precision=precision, out=out, table=table, table_offset=l0,
                      normalize=normalize, out_tensor=out_tensor, symmetric=symmetric)


def score_row_block(z: torch.Tensor, weight: torch.Tensor, rank: int, world_size: int, *, out: str = "rank",
                    table: Optional[RankTable] = None, precision: str = "bf16", normalize: bool = False) -> torch.Tensor:
    """The other axis of the SURVEY 8e partition, for L < number of GPUs or for load balance: this rank scores its block
    of head-drug ROWS (`row_shard`) against the whole catalogue for every outcome -> [L, rows_of_rank, N].  Concatenating
    the ranks' blocks along dim 1 gives the single-GPU tensor bit for bit (full N x N layout; the normaliser layout
    needs whole outcomes and shards by outcome instead)."""
    r0, r1 = row_shard(z.shape[0], rank, world_size)
    return pair_score(z[r0:r1].contiguous(), z, weight, precision=precision, out=out, table=table, normalize=normalize)


def score_all_pairs_to_host(z: torch.Tensor, weight: torch.Tensor, out_host: torch.Tensor, *, out: str = "rank",
                            table: Optional[RankTable] = None, precision: str = "bf16", chunk: int = 10,
                            normalize: bool = False, symmetric: bool = False, packed: bool = False,
                            host_mirror: bool = False, mirror_threads: int = 0) -> torch.Tensor:
    """predict.py:420-429 call pattern: outcomes in chunks of `chunk`, each chunk scored on the GPU and copied into
    `out_host` ([L, N, N], pinned for overlap).  Double-buffered: copy of chunk c overlaps compute of chunk c+1.
    packed=True (rank output): the device writes and the host receives the packed lower-triangular tiles
    ([L, T, 32, 32], half the PCIe volume); `decoder.unpack_packed_tiles` rebuilds [L, N, N] on the host when needed.
    host_mirror=True (symmetric rank output): `out_host` is still the drop-in [L, N, N] array, but only the packed tiles
    cross PCIe (into a cached pinned staging buffer) and `mdg_host_mirror_tiles` writes the mirror image on
    `mirror_threads` host threads while the next chunk is computed and copied — worth it where the host's memory
    system is faster than its PCIe link."""
    if host_mirror:
        if out != "rank" or not symmetric or packed:
            raise ValueError("host_mirror=True needs out='rank', symmetric=True, packed=False")
        return _score_all_pairs_to_host_mirrored(z, weight, out_host, table=table, precision=precision, chunk=chunk,
                                                 normalize=normalize, threads=mirror_threads)
    L, N = weight.shape[0], z.shape[0]
    dtype = _OUT_DTYPE[out]
    from .decoder import packed_tiles_per_outcome
    item_shape = (packed_tiles_per_outcome(N), 32, 32) if packed else (N, N)
    if tuple(out_host.shape) != (L,) + item_shape or out_host.dtype != dtype:
        raise ValueError(f"out_host must be {dtype} {(L,) + item_shape}")
    dev = z.device
    compute = torch.cuda.current_stream(dev)
    copy = _copy_stream(dev)
    bufs = [torch.empty((min(chunk, L),) + item_shape, dtype=dtype, device=dev) for _ in range(2)]
    done_copy = [None, None]
    for ci, l0 in enumerate(range(0, L, chunk)):
        l1 = min(l0 + chunk, L)
        b = ci & 1
        if done_copy[b] is not None:
            compute.wait_event(done_copy[b])  # buffer b is free once its previous copy finished
        dst = bufs[b][: l1 - l0]
        pair_score(z, z, weight[l0:l1], precision=precision, out=out, table=table, table_offset=l0,
                   normalize=normalize, out_tensor=dst, symmetric=symmetric, packed=packed)
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(copy):
            copy.wait_event(ready)
            out_host[l0:l1].copy_(dst, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
            done_copy[b] = ev
    compute.wait_stream(copy)
    # The reference writes each chunk into its memmap synchronously (predict.py:428-429): when this function returns the
    # host buffer must hold every chunk, so wait for the last device-to-host copies here, not in the caller.
    for ev in done_copy:
        if ev is not None:
            ev.synchronize()
    return out_host


_pinned_staging = {}


def _pinned(shape, dtype, slot: int) -> torch.Tensor:
    """Cached pinned staging buffers (page-locking hundreds of MB costs tens of ms: never inside a scoring call twice)."""
    key = (tuple(shape), dtype, slot)
    if key not in _pinned_staging:
        _pinned_staging[key] = torch.empty(shape, dtype=dtype).pin_memory()
    return _pinned_staging[key]


def _score_all_pairs_to_host_mirrored(z, weight, out_host, *, table, precision, chunk, normalize, threads):
    """score_all_pairs_to_host(host_mirror=True): kernel c+1 and the packed copy of chunk c+1 run on the GPU while the
    host threads mirror chunk c into `out_host`."""
    from .decoder import mirror_packed_tiles_host, packed_tiles_per_outcome
    L, N = weight.shape[0], z.shape[0]
    if tuple(out_host.shape) != (L, N, N) or out_host.dtype != torch.uint16 or out_host.is_cuda \
            or not out_host.is_contiguous():
        raise ValueError(f"out_host must be a contiguous host uint16 tensor {(L, N, N)}")
    dev = z.device
    T = packed_tiles_per_outcome(N)
    n = min(chunk, L) if L > 0 else 1
    compute = torch.cuda.current_stream(dev)
    copy = _copy_stream(dev)
    dbufs = [torch.empty((n, T, 32, 32), dtype=torch.uint16, device=dev) for _ in range(2)]
    hbufs = [_pinned((n, T, 32, 32), torch.uint16, b) for b in range(2)]
    pending = [None, None]  # (copy-done event, l0, l1) per buffer pair

    def drain(b):
        if pending[b] is not None:
            ev, a0, a1 = pending[b]
            ev.synchronize()
            mirror_packed_tiles_host(hbufs[b][: a1 - a0], N, out=out_host[a0:a1], threads=threads)
            pending[b] = None

    for ci, l0 in enumerate(range(0, L, chunk)):
        l1 = min(l0 + chunk, L)
        b = ci & 1
        drain(b)  # chunk ci - 2 is on the host and mirrored: both buffers of pair b are free (host-synchronous)
        dst = dbufs[b][: l1 - l0]
        pair_score(z, z, weight[l0:l1], precision=precision, out="rank", table=table, table_offset=l0,
                   normalize=normalize, out_tensor=dst, symmetric=True, packed=True)
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(copy):
            copy.wait_event(ready)
            hbufs[b][: l1 - l0].copy_(dst, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        pending[b] = (ev, l0, l1)
    nch = (L + chunk - 1) // chunk
    drain(nch & 1)        # the older of the two chunks still in flight first
    drain((nch & 1) ^ 1)
    compute.wait_stream(copy)
    return out_host


def raw_scores_path(checkpoint_dir: str, eval_type: str, drug_group_str: str, epoch, all_outcomes: bool = True) -> str:
    """The `.npy` file name the reference's drivers write and its notebooks read back (predict.py:412-436, 460-462)."""
    which = "all" if all_outcomes else "selected"
    return f"{checkpoint_dir}/{eval_type}_{which}_outcomes_{drug_group_str}_drugs_raw_scores_{epoch}.npy"


def score_all_pairs_to_npy(z: torch.Tensor, weight: torch.Tensor, path: str, *, out: str = "logit",
                           outcome_inds=None, table: Optional[RankTable] = None, precision: str = "fp32",
                           chunk: int = 10, normalize: bool = False, symmetric: bool = False):
    """predict.py:412-436 / 439-462: all-pairs scores of the selected drugs written to a `.npy` memmap, outcomes in chunks
    of `chunk` (or only `outcome_inds`, predict.py:447-452), returned as `np.load(path, mmap_mode='r')`.

    The reference fills a `.raw` memmap and then copies it into a `.npy` with `np.save`; here the `.npy` is created
    directly (`np.lib.format.open_memmap`) and each chunk goes GPU -> pinned staging buffer -> file, the device-to-host
    copy of chunk c overlapping the kernel of chunk c+1."""
    import numpy as np
    W = weight if outcome_inds is None else weight[torch.as_tensor(outcome_inds, device=weight.device)].contiguous()
    L, N = W.shape[0], z.shape[0]
    np_dtype = {"logit": np.float32, "sigmoid": np.float32, "rank": np.uint16}[out]
    fp = np.lib.format.open_memmap(path, mode="w+", dtype=np_dtype, shape=(L, N, N))
    dev = z.device
    compute = torch.cuda.current_stream(dev)
    copy = _copy_stream(dev)
    n = min(chunk, L) if L > 0 else 1
    dbufs = [torch.empty((n, N, N), dtype=_OUT_DTYPE[out], device=dev) for _ in range(2)]
    hbufs = [torch.empty((n, N, N), dtype=_OUT_DTYPE[out]).pin_memory() for _ in range(2)]
    pending = [None, None]  # (event, l0, l1) of the copy in flight per buffer

    def drain(b):
        if pending[b] is not None:
            ev, a0, a1 = pending[b]
            ev.synchronize()
            fp[a0:a1] = hbufs[b][: a1 - a0].numpy()
            pending[b] = None

    for ci, l0 in enumerate(range(0, L, chunk)):
        l1 = min(l0 + chunk, L)
        b = ci & 1
        drain(b)  # the previous use of this buffer pair has reached the file
        dst = dbufs[b][: l1 - l0]
        pair_score(z, z, W[l0:l1], precision=precision, out=out, table=table, table_offset=l0 if outcome_inds is None else 0,
                   normalize=normalize, out_tensor=dst, symmetric=symmetric)
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(copy):
            copy.wait_event(ready)
            hbufs[b][: l1 - l0].copy_(dst, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        pending[b] = (ev, l0, l1)
    drain(0)
    drain(1)
    compute.wait_stream(copy)
    fp.flush()
    del fp
    return np.load(path, mmap_mode="r")


_copy_streams = {}


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.type, dev.index)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=dev)
    return _copy_streams[key]


def ensemble_mean_sigmoid(z_list, weight_list, *, precision: str = "bf16",
                          label_range: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """predict.py:493, 612: mean over checkpoints of sigmoid(raw scores), per checkpoint's (z, W)."""
    from .decoder import ensemble_reduce
    members = [score_all_pairs(z, W, out="sigmoid", precision=precision, label_range=label_range)
               for z, W in zip(z_list, weight_list)]
    return ensemble_reduce(members, "mean")


def ensemble_normalized_ranks_chunks(z_list, weight_list, *, precision: str = "fp32", chunk: int = 2,
                                     normalize: bool = False):
    """The reference's whole ensemble normalisation (generate_embeddings.ipynb cells 13-20) streamed by outcome chunk,
    nothing but one chunk ever materialised: for each checkpoint k, raw scores (models.py:537-547) -> in-sample
    normalised ranks (normalize_scores.py:62-74); geometric mean over checkpoints (cell 18); `run_slice` again on the
    gmean (cell 20).  Yields (l0, l1, ranks[l1-l0, N, N] fp32); the yielded tensor is reused by the next iteration."""
    from .normalize import exact_normalized_ranks, gmean_normalized_ranks
    L = weight_list[0].shape[0]
    for l0 in range(0, L, chunk):
        l1 = min(l0 + chunk, L)
        members = []
        for z, W in zip(z_list, weight_list):
            raw = pair_score(z, z, W[l0:l1], precision=precision, out="logit", normalize=normalize)
            members.append(exact_normalized_ranks(raw))
            del raw
        g = gmean_normalized_ranks(members)
        del members
        yield l0, l1, exact_normalized_ranks(g)


def ensemble_fused_ranks_chunks(z_list, weight_list, member_tables, ens_table, *, precision: str = "bf16",
                                chunk: int = 16, normalize: bool = False, packed: bool = False):
    """The ensemble normalisation in the quantile formulation, streamed by outcome chunk: per checkpoint k the fused
    uint16 ranks against its own table (one kernel, logits never materialised), then `mdg_ensemble_rank_u16` (fixed-point
    log-sum + ensemble-table lookup).  Yields (l0, l1, ranks uint16 [l1-l0, N, N] in the normaliser layout); the yielded
    tensor is reused by the next iteration.  `weight_list` entries may be PreparedDecoder handles.  Build `ens_table`
    with `normalize.build_ensemble_rank_table` from the members' ranks over a reference panel.
    packed=True: members and result stay in the packed lower-triangular tile layout ([.., T, 32, 32], no mirror image):
    half the look-ups' traffic in every kernel; `decoder.unpack_packed_tiles` gives the mirrored array."""
    from .decoder import packed_tiles_per_outcome
    from .normalize import ensemble_fused_ranks
    L = weight_list[0].shape[0]
    N = z_list[0].shape[0]
    K = len(z_list)
    dev = z_list[0].device
    item = (packed_tiles_per_outcome(N), 32, 32) if packed else (N, N)
    bufs = [torch.empty((min(chunk, L),) + item, dtype=torch.uint16, device=dev) for _ in range(K)]
    out = torch.empty((min(chunk, L),) + item, dtype=torch.uint16, device=dev)
    for l0 in range(0, L, chunk):
        l1 = min(l0 + chunk, L)
        members = []
        for k, (z, W, tbl) in enumerate(zip(z_list, weight_list, member_tables)):
            members.append(pair_score(z, z, W[l0:l1], precision=precision, out="rank", table=tbl, table_offset=l0,
                                      normalize=normalize, out_tensor=bufs[k][: l1 - l0], symmetric=True, packed=packed))
        sub = _EnsembleView(ens_table, l0, l1)
        yield l0, l1, ensemble_fused_ranks(members, sub, out=out[: l1 - l0])


class _EnsembleView:
    """Outcome slice [l0, l1) of an EnsembleRankTable (same ilog table, offset rank table)."""

    def __init__(self, ens, l0, l1):
        self.ilog, self.member_Q = ens.ilog, ens.member_Q
        self.table = _TableView(ens.table, l0, l1)


class _TableView:
    def __init__(self, table, l0, l1):
        self._t, self._l0, self.L = table, l0, l1 - l0

    def struct(self, a=0, b=None):
        b = self.L if b is None else b
        return self._t.struct(self._l0 + a, self._l0 + b)


def top_pairs_per_outcome(z: torch.Tensor, weight: torch.Tensor, k: int, table: RankTable, *, precision: str = "bf16",
                          cap: int = 65536, max_rounds: int = 20, normalize: bool = False):
    """Per-outcome top-k unordered pairs of one catalogue (BASELINE "top-1000 per outcome") with the candidate
    threshold taken from the rank table and adjusted per outcome until every list is complete.

    The fused top-k kernel keeps the scores >= a per-outcome threshold (mdg_pair_topk).  The starting threshold is the
    table quantile that leaves ~3k candidates if the table's panel represents the catalogue.  An outcome that came
    back short (status 1) needs a lower quantile, one that overflowed `cap` (status 2) a higher one: the quantile
    index is moved in growing steps until the outcome is bracketed, then bisected; only the failed outcomes are
    re-run each round.  An overflow at the table's last quantile grows the candidate list 4x instead.  Any threshold
    with status 0 yields the exact top-k.  Returns (scores [L,k], rows, cols, status, rounds); status is all-zero
    unless the table cannot bracket an outcome within max_rounds."""
    N, L, Q = z.shape[0], weight.shape[0], table.Q
    M = N * (N - 1) // 2
    if k > M:
        raise ValueError("k exceeds the number of unordered pairs")
    dev = z.device
    q0 = min(Q - 2, max(0, int(Q * (1.0 - 3.0 * k / M)) - 1))
    qidx = torch.full((L,), q0, dtype=torch.int64, device=dev)
    lo = torch.full((L,), -1, dtype=torch.int64, device=dev)   # highest index known to overflow
    hi = torch.full((L,), Q, dtype=torch.int64, device=dev)    # lowest index known to come back short
    thr = table.thresholds.gather(1, qidx[:, None]).reshape(-1).contiguous()
    scores, rows, cols, status = pair_topk(z, z, weight, thr, k, cap=cap, symmetric=True, precision=precision,
                                           normalize=normalize)
    rounds, step, caps = 1, 1, cap
    while rounds < max_rounds:
        bad = torch.nonzero(status != 0).reshape(-1)
        if bad.numel() == 0:
            break
        st, q = status[bad], qidx[bad]
        lo[bad] = torch.where(st == 2, torch.maximum(lo[bad], q), lo[bad])
        hi[bad] = torch.where(st == 1, torch.minimum(hi[bad], q), hi[bad])
        if bool(((st == 2) & (q >= Q - 1)).any()):
            caps *= 4
        step *= 4
        bracketed = (lo[bad] >= 0) & (hi[bad] < Q)
        walk = torch.where(st == 1, q - step, q + step)
        new_q = torch.where(bracketed, (lo[bad] + hi[bad]) // 2, walk).clamp_(0, Q - 1)
        qidx[bad] = new_q
        thr_b = table.thresholds[bad].gather(1, new_q[:, None]).reshape(-1).contiguous()
        s_b, r_b, c_b, st_b = pair_topk(z, z, weight[bad].contiguous(), thr_b, k, cap=caps, symmetric=True,
                                        precision=precision, normalize=normalize)
        scores[bad], rows[bad], cols[bad], status[bad] = s_b, r_b, c_b, st_b
        rounds += 1
    return scores, rows, cols, status, rounds


def bind_host_thread_to_gpu(device_index: int):
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (intersected with the cores
    the process is allowed to use), so that pinned host buffers allocated afterwards are placed on the GPU's NUMA node
    and the D2H stream of the rank tensor does not cross the inter-socket link.  One process per GPU, call it before
    allocating pinned memory.  Returns (previous cpu set, new cpu set) or None when NVML / affinity information is
    unavailable or the intersection is empty (nothing is changed then)."""
    import os
    if not hasattr(os, "sched_getaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(ncpu, 1024) + 63) // 64)
    except Exception:
        return None
    local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
    before = set(os.sched_getaffinity(0))
    target = before & local
    if not target or target == before:
        return None
    try:
        os.sched_setaffinity(0, target)
    except OSError:
        return None
    return before, target
