#!/usr/bin/env python
"""Benchmark of the drug-pair scoring path (BASELINE.json metric: scored (outcome, drugA, drugB) triples/sec, fused
rank).  Contract: one JSON line on stdout from rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload auto|configs1|configs2]

Workloads (BASELINE.json `configs` indices):
  N = 1 : configs[1] — 4,096 drugs x 86 outcomes, hidden 256, all-pairs fused scoring + uint16 quantile rank on 1 B200.
  N > 1 : configs[2] — 4,096 drugs x 963 outcomes, STRONG scaling: the same 963 outcomes are sharded over the N ranks
          (`scoring.outcome_shard`, the reference's TWOSIDES call pattern predict.py:381-463); drugs are row-sharded for
          the encoder and the only collective is the all-gather of the fused-embedding table.  Untimed, inside the same
          run: every rank checks one of its outcomes bit for bit against the CPU oracle, the all-reduced checksum of all
          ranks is compared with a single-GPU pass over all 963 outcomes on rank 0 (which also gives the N = 1 time of
          the SAME workload), and with >= 6 ranks a second timed leg runs configs[3] (20,000 drugs x 953 outcomes, uint16
          ranks and per-outcome top-1000) and is reported under `config3_20k_x_953`.

A "step" = one pass of the hot path over the whole batch: fusion encoder on this rank's drugs -> [all-gather z] ->
operand prep -> GEMM 1 (z.W_l, prepared decoder weights) -> GEMM 2 + fused rank epilogue writing uint16 ranks for every
(outcome, drugA, drugB).
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

HIDDEN = 256
N_TOKENS = 4          # BASELINE: "up to four modality tokens per drug (structure, KG, transcriptomic, cell-viability)"
ENC = dict(embed_dim=HIDDEN, num_layers=2, num_heads=8, head_dim=32, ffn_dim=512, actn="gelu", norm_first=True,
           agg="x-attn", nb=0)  # latent 256 = 8 x 32, FFN 2 x latent, pre-LN + GELU + x-attn pooling as shipped configs
Q_TABLE = 16384
PANEL = 2048
TOPK = 1000
METRIC = "scored (outcome, drugA, drugB) triples/sec, fused rank"
UNIT = "triples/s"
WORKLOADS = {
    "configs1": dict(name="BASELINE configs[1]", drugs=4096, outcomes=86),
    "configs2": dict(name="BASELINE configs[2]", drugs=4096, outcomes=963),
    "configs3": dict(name="BASELINE configs[3]", drugs=20000, outcomes=953),
}
CPU_WORKERS = 16      # fixed worker count of the CPU arm (when the box has that many cores): BENCH and SCALE boxes differ


def pick_workload(args):
    if args.workload != "auto":
        return args.workload
    return "configs1" if args.gpus == 1 else "configs2"


def workload_config(wl_key, n_gpus):
    wl = WORKLOADS[wl_key]
    n, L = wl["drugs"], wl["outcomes"]
    strong = wl_key != "configs1"
    return {
        "workload": f"{wl['name']}: {n} drugs x {L} outcomes" + (f" sharded by outcome over {n_gpus} GPU(s)" if strong else "")
                    + f", hidden {HIDDEN}: fusion encoder ({N_TOKENS} modality tokens/drug, random missing-modality "
                    f"masks, 2 layers, 8 heads, latent 256, FFN 512, x-attn pooling) -> all-pairs bf16-input/fp32-accumulate "
                    f"bilinear scoring -> fused uint16 rank (Q={Q_TABLE} reference quantiles/outcome from a {PANEL}-drug panel)",
        "drugs": n, "outcomes_total": L, "outcomes_per_gpu": -(-L // n_gpus), "hidden": HIDDEN,
        "pairs": "one catalogue scored against itself in the reference normaliser's layout (notebooks/normalize_scores.py:"
                 "67-70): each unordered pair (row > col) is scored and ranked once and its rank written at [l,i,j] and "
                 "[l,j,i], diagonal 0; `value` counts the L*N*N uint16 entries written (ordered triples)",
        "parallelism": (f"outcomes sharded over {n_gpus} GPUs for the decoder (strong scaling: total work fixed); encoder: "
                        f"one wave of 128-row tiles covers this catalogue, so every rank encodes all of it and the step "
                        f"has NO exchange (scoring.encoder_is_replicated); larger catalogues (the configs[3] leg) are "
                        f"row-sharded and replicated by one peer all-gather of z per step") if n_gpus > 1 else "1 GPU",
        "decoder_weights": "prepared once (mdg_pair_prepare): constant between checkpoint loads, outside the timed region",
        "l2": "no explicit flush: every step streams >= 2.9 GB of output per GPU through the 126 MB L2, evicting the inputs",
    }


def outcome_weights(l0, l1, D=HIDDEN):
    """Symmetric decoder weights of outcomes [l0, l1): outcome l is seeded by l alone, so that any sharding of the
    outcomes sees the same W[l] (nn.Bilinear init U(+-1/sqrt(D)) through the Symmetric parametrisation, models.py:522-524)."""
    W = np.empty((l1 - l0, D, D), dtype=np.float32)
    b = np.float32(1.0 / np.sqrt(D))
    for i, l in enumerate(range(l0, l1)):
        P = np.random.default_rng(100_000 + l).uniform(-b, b, size=(D, D)).astype(np.float32)
        W[i] = np.triu(P) + np.triu(P, 1).T
    return W


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own code for this path on the host cores when it was staged into oracle/_ref (git-ignored,
# travels with the snapshot; oracle/make_ref.py), else the numpy port in oracle/oracle.py.
#   encoder   = TransformerFusion.forward                       (models.py:401-455, torch fp32, all host threads)
#   decoder   = BilinearDDIScorer.forward, outcomes in chunks of 10 into a raw-score memmap   (predict.py:420-429)
#   normalise = run_slice per outcome in multiprocessing.Pool   (notebooks/normalize_scores.py:62-85)
# ----------------------------------------------------------------------------------------------------------------
_REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_CPU_STATE = {}


def _reference_staged():
    return os.path.isdir(os.path.join(_REF_DIR, "madrigal"))


def _cpu_cores():
    n = os.cpu_count() or 1
    return min(CPU_WORKERS, n)


def _ref_run_slice(sl):
    _CPU_STATE["run_slice"](sl)           # the reference function, writing into the shared memmap
    return sl[0]


def _port_one_outcome(l):
    from oracle import oracle
    raw, out = _CPU_STATE["raw"], _CPU_STATE["out"]
    out[l:l + 1] = oracle.normalize_scores(np.array(raw[l:l + 1]))
    return l


def cpu_reference_pass(n_drugs, n_outcomes, cores):
    """Encode + score + rank-normalise `n_outcomes` outcomes of the n_drugs-drug workload on the host; (seconds, kind)."""
    import multiprocessing as mp
    import tempfile
    import synth
    tok, msk = synth.fusion_inputs(n_drugs, N_TOKENS, HIDDEN, seed=0)
    sd = synth.fusion_state_dict(ENC, seed=7)
    W = outcome_weights(0, n_outcomes)
    import shutil
    need = 2 * n_outcomes * n_drugs * n_drugs * 4
    shm = None   # RAM-backed scratch for the two memmaps when it is large enough, else the default temp directory
    try:
        if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 1.25 * need:
            shm = "/dev/shm"
    except OSError:
        shm = None
    kind = "reference" if _reference_staged() else "port"
    with tempfile.TemporaryDirectory(dir=shm) as tmp:
        raw = np.lib.format.open_memmap(os.path.join(tmp, "raw.npy"), mode="w+", dtype=np.float32,
                                        shape=(n_outcomes, n_drugs, n_drugs))
        out = np.lib.format.open_memmap(os.path.join(tmp, "norm.npy"), mode="w+", dtype=np.float32,
                                        shape=(n_outcomes, n_drugs, n_drugs))
        if kind == "reference":
            import torch
            os.environ["MADRIGAL_REFERENCE_ROOT"] = _REF_DIR
            from oracle import ref_import
            ref_import.REFERENCE_ROOT = _REF_DIR
            models = ref_import.load_reference_models()
            _, make_run_slice = ref_import.load_reference_normalizer()
            torch.set_num_threads(cores)
            enc = models.TransformerFusion(HIDDEN, 0, ENC["num_layers"], ENC["num_heads"], ENC["head_dim"], ENC["ffn_dim"],
                                           transformer_dropout=0.0, transformer_actn="gelu", transformer_norm_first=True,
                                           transformer_batch_first=False, transformer_agg="x-attn").eval()
            enc.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
            enc.x_attn_key_padding_mask = torch.zeros(1, N_TOKENS, dtype=torch.bool)   # nb = 0: every token is a key
            dec = models.BilinearDDIScorer(HIDDEN, HIDDEN, n_outcomes).eval()
            with torch.no_grad():
                dec.weight.copy_(torch.from_numpy(W))
            t0 = time.perf_counter()
            with torch.no_grad():
                z = enc(torch.from_numpy(tok), torch.from_numpy(msk))
                for l0 in range(0, n_outcomes, 10):                                     # predict.py:420-429
                    l1 = min(l0 + 10, n_outcomes)
                    raw[l0:l1] = dec(z, z, (l0, l1)).numpy()
            _CPU_STATE["run_slice"] = make_run_slice(raw, out)
            ctx = mp.get_context("fork")
            with ctx.Pool(processes=cores) as pool:                                      # normalize_scores.py:78-85
                pool.map(_ref_run_slice, [(l, l + 1) for l in range(n_outcomes)])
            dt = time.perf_counter() - t0
        else:
            from oracle import oracle
            t0 = time.perf_counter()
            z = oracle.fusion_forward(sd, ENC, tok, msk, None, np.zeros(N_TOKENS, bool))
            for l0 in range(0, n_outcomes, 10):
                l1 = min(l0 + 10, n_outcomes)
                raw[l0:l1] = oracle.bilinear_scores(z, z, W, (l0, l1))
            _CPU_STATE["raw"], _CPU_STATE["out"] = raw, out
            ctx = mp.get_context("fork")
            with ctx.Pool(processes=cores) as pool:
                pool.map(_port_one_outcome, range(n_outcomes))
            dt = time.perf_counter() - t0
        del raw, out
        _CPU_STATE.clear()
    return dt, kind


def _cpu_sample_text(n_out, wl, cores, kind, dt):
    what = ("the unmodified reference staged in oracle/_ref (madrigal/models/models.py TransformerFusion + "
            "BilinearDDIScorer, notebooks/normalize_scores.py run_slice)") if kind == "reference" else \
        "the numpy port oracle/oracle.py (oracle/_ref not staged on this box)"
    return (f"{n_out} of {wl['outcomes']} outcomes x {wl['drugs']}^2 pairs: fp32 fusion encoder for all {wl['drugs']} drugs "
            f"+ fp32 decoder in chunks of 10 outcomes + exact argsort rank normaliser, one outcome per task in "
            f"Pool({cores}); {what}; {dt:.1f} s")


def cpu_baseline(wl_key):
    wl = WORKLOADS[wl_key]
    cores = _cpu_cores()
    n_out = max(1, min(4 * cores, wl["outcomes"]))  # four outcomes per worker: ~12 s wall on the 16-core box
    dt, kind = cpu_reference_pass(wl["drugs"], n_out, cores)
    v = n_out * wl["drugs"] * wl["drugs"] / dt
    return {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "value_per_core": v / cores,
            "sample": _cpu_sample_text(n_out, wl, cores, kind, dt)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl_key = pick_workload(args)
    wl = WORKLOADS[wl_key]
    cores = _cpu_cores()
    n_out = max(1, min(2 * cores, wl["outcomes"]))
    kind = "port"
    for _ in range(min(args.warmup, 1)):
        cpu_reference_pass(wl["drugs"], 1, 1)
    times = []
    for _ in range(args.steps):
        dt, kind = cpu_reference_pass(wl["drugs"], n_out, cores)
        times.append(dt)
    dt = float(np.mean(times))
    value = n_out * wl["drugs"] * wl["drugs"] / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak" if wl_key == "configs1" else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(wl_key, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "value_per_core": value / cores,
                         "sample": "each step = " + _cpu_sample_text(n_out, wl, cores, kind, dt)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML from a background thread DURING the timed region."""

    def __init__(self, gpu_index, period_s=0.002):
        import threading
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None
            return
        self._period = period_s
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self._period)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": float(np.median(sm[len(sm) // 2:])) if sm else None,  # upper half = samples under load
                "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(sm)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class Job:
    """One rank's share of a workload: encoder row shard -> all-gather -> decoder over outcomes [l0, l1)."""

    def __init__(self, ctx, wl_key, l0, l1, world, rank, alloc_out=True):
        import torch
        import madrigal_b200 as mb
        from madrigal_b200 import normalize, scoring
        self.ctx, self.wl, self.l0, self.l1, self.world, self.rank = ctx, WORKLOADS[wl_key], l0, l1, world, rank
        dev = ctx["dev"]
        n = self.wl["drugs"]
        tokens, masks = ctx["inputs"](n)
        self.n = n
        # one wave of the fused encoder covers the catalogue -> every rank encodes all of it, no exchange (scoring.py)
        self.replicated = scoring.encoder_is_replicated(n, tokens.shape[1], world, dev)
        r0, r1 = (0, n) if self.replicated else scoring.row_shard(n, rank, world)
        self.tok_shard, self.mask_shard = tokens[r0:r1].contiguous(), masks[r0:r1].contiguous()
        self.W = torch.from_numpy(outcome_weights(l0, l1)).to(dev)
        self.prepared = mb.decoder.PreparedDecoder(self.W, "bf16")
        with torch.no_grad():
            self.z_full = ctx["encoder"](tokens, masks)
        # setup (untimed): per-outcome reference quantiles from a drug panel -> prepared rank table (exact bucket LUT)
        self.quantiles = normalize.build_reference_quantiles(self.z_full, self.W, Q_TABLE, panel=PANEL, precision="bf16")
        self.table = mb.RankTable(self.quantiles)
        self.out = torch.empty((l1 - l0, n, n), dtype=torch.uint16, device=dev) if alloc_out else None
        self.gatherer = ctx["gatherer"](n) if world > 1 and not self.replicated else None
        self.launches = 0
        torch.cuda.synchronize()

    def encode(self):
        from madrigal_b200 import scoring
        z = self.ctx["encoder"](self.tok_shard, self.mask_shard)        # fusion encoder on this rank's drugs
        self.launches = self.ctx["encoder"].last_launch_count
        if self.gatherer is not None:
            z = self.gatherer.gather(z)                                  # the path's only exchange step
            self.launches += 1
        return z

    def step(self, table=None):
        import torch
        import madrigal_b200 as mb
        from madrigal_b200 import _lib
        with torch.no_grad():
            z = self.encode()
            mb.pair_score(z, z, self.prepared, out="rank", table=table or self.table, out_tensor=self.out, symmetric=True)
            self.launches += _lib.lib().mdg_last_launch_count()

    def checksum(self):
        """(sum of all uint16 ranks, position-weighted sum) as Python ints — identical for any sharding of the outcomes."""
        import torch
        n = self.n
        wgt = ((torch.arange(n, device=self.out.device, dtype=torch.int64)[:, None] * 131
                + torch.arange(n, device=self.out.device, dtype=torch.int64)[None, :] * 7) % 251).to(torch.int32) \
            if n <= 8192 else None
        s0, s1 = 0, 0
        for l in range(self.out.shape[0]):
            v = self.out[l].view(torch.int16).to(torch.int32)     # ranks <= 65535 >> 1 here (Q = 16384): no sign issue
            s0 += int(v.sum(dtype=torch.int64).item())
            if wgt is not None:
                s1 += int((v * wgt).sum(dtype=torch.int64).item()) * ((self.l0 + l) % 97 + 1)
        return s0, s1

    def parity_one_outcome(self, j=0, sample_rows=None):
        """Outcome j of this rank, bit for bit against the CPU oracle's searchsorted on the same logits, in the
        normaliser layout (row > col ranked, mirrored, zero diagonal).  sample_rows: check only these rows (large N)."""
        import torch
        import madrigal_b200 as mb
        from oracle import oracle
        n = self.n
        thr = self.table.thresholds[j:j + 1].cpu().numpy()
        if sample_rows is None:
            lg = mb.pair_score(self.z_full, self.z_full, self.W[j:j + 1], precision="bf16", out="logit").cpu().numpy()
            exp = oracle.quantile_rank(thr, lg, "right")[0].astype(np.uint16)
            ref = np.tril(exp, -1)
            ref = ref + ref.T
            got = self.out[j].cpu().numpy()
            return bool(np.array_equal(ref, got))
        rows = torch.as_tensor(sample_rows, device=self.z_full.device)
        lg = mb.pair_score(self.z_full[rows].contiguous(), self.z_full, self.W[j:j + 1], precision="bf16",
                           out="logit").cpu().numpy()                       # [1, R, n]
        exp = oracle.quantile_rank(thr, lg, "right")[0].astype(np.uint16)
        o16 = self.out[j].view(torch.int16)     # torch has no CUDA index kernel for uint16: same bits as int16
        got = o16[rows].cpu().numpy().view(np.uint16)
        got_t = o16[:, rows].cpu().numpy().view(np.uint16).T
        ok = True
        for k, r in enumerate(sample_rows):
            ok &= bool(np.array_equal(exp[k, :r], got[k, :r])) and int(got[k, r]) == 0   # row > col part + diagonal
            ok &= bool(np.array_equal(got[k], got_t[k]))                                  # mirrored
        return ok


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import madrigal_b200 as mb
    from madrigal_b200 import _lib, scoring
    import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    numa = scoring.bind_host_thread_to_gpu(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().mdg_check_device(local_rank), "mdg_check_device")
    wl_key = pick_workload(args)
    wl = WORKLOADS[wl_key]
    N, L_total = wl["drugs"], wl["outcomes"]
    steps, warm = args.steps, max(args.warmup, 3)

    # ---- synthetic inputs (seeded): shared drug catalogue (modality tokens + masks), globally seeded outcomes
    encoder = mb.TransformerFusion(HIDDEN, 0, ENC["num_layers"], ENC["num_heads"], ENC["head_dim"], ENC["ffn_dim"],
                                   transformer_actn=ENC["actn"], transformer_norm_first=True,
                                   transformer_batch_first=False, transformer_agg="x-attn", precision="bf16")
    encoder.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(ENC, seed=7).items()})
    encoder.x_attn_key_padding_mask = torch.zeros(1, N_TOKENS, dtype=torch.bool)  # nb = 0: every token is a key
    encoder = encoder.to(dev).eval()
    host_inputs = {}

    def inputs(n):
        if n not in host_inputs:
            host_inputs[n] = synth.fusion_inputs(n, N_TOKENS, HIDDEN, seed=0)
        t, m = host_inputs[n]
        return torch.from_numpy(t).to(dev), torch.from_numpy(m).to(dev)

    gatherers = {}

    def gatherer(n):   # one peer-mapped table pair per catalogue size (shared by the jobs of this process)
        if n not in gatherers:
            gatherers[n] = scoring.PeerAllGather(n, HIDDEN, dev)
        return gatherers[n]

    ctx = {"dev": dev, "encoder": encoder, "inputs": inputs, "gatherer": gatherer}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(fn, n_steps, n_warm, profile=False):
        """(ms per step: CUDA events between barriers, max over ranks; mean ms of the profiled dominant kernel)."""
        for _ in range(n_warm):
            fn()
        barrier()
        if profile:
            _lib.check(_lib.lib().mdg_profile_enable(min(n_steps, 256)), "mdg_profile_enable")
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(n_steps):
            fn()
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1) / n_steps)
        kern = None
        if profile:
            buf = (ctypes.c_float * 256)()
            n_rec = _lib.lib().mdg_profile_read(buf, 256)
            _lib.lib().mdg_profile_enable(0)
            kern = float(np.mean(buf[:n_rec])) if n_rec > 0 else None
        return ms, kern

    # ================================================================== main timed leg
    l0, l1 = scoring.outcome_shard(L_total, rank, world)
    job = Job(ctx, wl_key, l0, l1, world, rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_per_step, kern_ms = timed(job.step, steps, warm, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = job.launches
    triples_per_step = L_total * N * N
    value = triples_per_step / (ms_per_step * 1e-3)
    per_gpu_triples = (l1 - l0) * N * N

    # ---- parity (untimed): one outcome per rank bit for bit against the CPU oracle; checksum over every rank's output
    parity_ok = all_true(job.parity_one_outcome(0))
    s0, s1 = job.checksum()
    ck = torch.tensor([s0, s1], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ck)
    checksum = [int(ck[0].item()), int(ck[1].item())]

    # ---- context (untimed): what a plain write of the same tensor costs on this GPU
    memset_ms = pwl_ms = pwl_dev = None
    if rank == 0:
        flat = job.out.view(torch.uint8).reshape(-1)
        for _ in range(3):
            flat.zero_()
        ms_list = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            flat.zero_()
            b.record()
            torch.cuda.synchronize()
            ms_list.append(a.elapsed_time(b))
        memset_ms = float(np.median(ms_list))
        job.out_is_stale = True
    # ---- secondary measurement (untimed): the same kernel with the histogram-CDF rank table (MDG_RANK_PWL)
    if rank == 0 and world == 1:
        table_pwl = mb.RankTable(job.quantiles, kind="pwl")
        with torch.no_grad():
            for _ in range(3):
                mb.pair_score(job.z_full, job.z_full, job.prepared, out="rank", table=table_pwl, out_tensor=job.out, symmetric=True)
            torch.cuda.synchronize()
            _lib.check(_lib.lib().mdg_profile_enable(10), "mdg_profile_enable")
            for _ in range(10):
                mb.pair_score(job.z_full, job.z_full, job.prepared, out="rank", table=table_pwl, out_tensor=job.out, symmetric=True)
        torch.cuda.synchronize()
        buf = (ctypes.c_float * 256)()
        n_pwl = _lib.lib().mdg_profile_read(buf, 256)
        _lib.lib().mdg_profile_enable(0)
        pwl_ms = float(np.mean(buf[:n_pwl])) if n_pwl > 0 else None
        pwl_dev = float(table_pwl.max_rank_deviation.max().item())
        del table_pwl
        job.out_is_stale = True
    if getattr(job, "out_is_stale", False):   # restore the exact-table output (rank 0's copy was overwritten above)
        with torch.no_grad():
            mb.pair_score(job.z_full, job.z_full, job.prepared, out="rank", table=job.table, out_tensor=job.out, symmetric=True)
    barrier()

    # ================================================================== e2e: host buffers in, host buffers out
    e2e_steps = max(1, min(steps, 3))
    host_cap = 10 << 30                                   # pinned-memory bound per rank
    Le = min(l1 - l0, max(1, host_cap // (2 * N * N)))    # outcomes of this rank's shard in the e2e sample
    tok_np, mask_np = host_inputs[N]
    replicated = scoring.encoder_is_replicated(N, tok_np.shape[1], world, dev)   # one encoder wave: no exchange step
    r0, r1 = (0, N) if replicated else scoring.row_shard(N, rank, world)
    tok_host = torch.from_numpy(tok_np[r0:r1]).pin_memory()
    mask_host = torch.from_numpy(mask_np[r0:r1]).pin_memory()
    W_host = job.W[:Le].cpu().pin_memory()
    out_host = torch.empty((Le, N, N), dtype=torch.uint16).pin_memory()

    @torch.no_grad()
    def e2e_step():
        zd = encoder(tok_host.to(dev, non_blocking=True), mask_host.to(dev, non_blocking=True))
        Wd = W_host.to(dev, non_blocking=True)
        if not replicated:
            zd = gatherer(N).gather(zd)
        scoring.score_all_pairs_to_host(zd, Wd, out_host, out="rank", table=job.table, precision="bf16", chunk=10,
                                        symmetric=True)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    le_t = torch.tensor([Le], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(le_t)
    e2e_triples = int(le_t.item()) * N * N
    e2e_value = e2e_triples / (e2e_ms * 1e-3)
    e2e_matches = all_true(bool(torch.equal(out_host[:1].view(torch.int16), job.out[:1].cpu().view(torch.int16))))
    h2d = tok_host.numel() * 4 + mask_host.numel() + W_host.numel() * 4
    d2h = out_host.numel() * 2

    # ---- the same drop-in [L, N, N] host array, but only the packed tiles cross PCIe and the library's host threads write
    #      the mirror image (mdg_host_mirror_tiles) while the next chunk is computed and copied
    e2e_mirror_ms, mirror_matches, mirror_threads = None, None, 0
    if world == 1:   # single-GPU line only: a second pinned [L, N, N] buffer per rank is not worth it on a shared host
        mirror_threads = max(1, (os.cpu_count() or 1) // max(1, world))
        out_host2 = torch.empty((Le, N, N), dtype=torch.uint16).pin_memory()

        @torch.no_grad()
        def e2e_mirror_step():
            zd = encoder(tok_host.to(dev, non_blocking=True), mask_host.to(dev, non_blocking=True))
            Wd = W_host.to(dev, non_blocking=True)
            if not replicated:
                zd = gatherer(N).gather(zd)
            scoring.score_all_pairs_to_host(zd, Wd, out_host2, out="rank", table=job.table, precision="bf16", chunk=10,
                                            symmetric=True, host_mirror=True, mirror_threads=mirror_threads)

        e2e_mirror_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_mirror_step()
        barrier()
        e2e_mirror_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
        mirror_matches = all_true(bool(torch.equal(out_host2.view(torch.int16), out_host.view(torch.int16))))
        del out_host2
    del out_host

    # ---- the same call with the reduced-volume output layout (MDG_PAIRS_PACKED_TILES: no mirror image, half the D2H)
    from madrigal_b200.decoder import packed_tiles_per_outcome, unpack_packed_tiles
    packed_host = torch.empty((Le, packed_tiles_per_outcome(N), 32, 32), dtype=torch.uint16).pin_memory()

    @torch.no_grad()
    def e2e_packed_step():
        zd = encoder(tok_host.to(dev, non_blocking=True), mask_host.to(dev, non_blocking=True))
        Wd = W_host.to(dev, non_blocking=True)
        if not replicated:
            zd = gatherer(N).gather(zd)
        scoring.score_all_pairs_to_host(zd, Wd, packed_host, out="rank", table=job.table, precision="bf16", chunk=10,
                                        packed=True)

    e2e_packed_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_packed_step()
    barrier()
    e2e_packed_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    packed_matches = all_true(bool(torch.equal(
        unpack_packed_tiles(packed_host[:1].to(dev), N).view(torch.int16), job.out[:1].view(torch.int16))))
    d2h_packed = packed_host.numel() * 2
    del packed_host, W_host

    # ================================================================== configs[2] on this one GPU (N = 1 line): the anchor
    # the strong-scaled N > 1 lines are compared with (same code path as their `single_gpu_same_workload` leg)
    c2_single = None
    if world == 1 and wl_key == "configs1" and not args.no_encoder_block:
        del job.out
        job.out = None
        torch.cuda.empty_cache()
        wl2 = WORKLOADS["configs2"]
        job2 = Job(ctx, "configs2", 0, wl2["outcomes"], 1, 0)
        ms2, kern2 = timed(job2.step, 5, 2, profile=True)
        ok2 = job2.parity_one_outcome(wl2["outcomes"] - 1)
        c0, c1 = job2.checksum()
        tr2 = wl2["outcomes"] * wl2["drugs"] * wl2["drugs"]
        c2_single = {"workload": f"{wl2['name']}: {wl2['drugs']} drugs x {wl2['outcomes']} outcomes on this ONE GPU "
                                 f"(what `bench.py --gpus N` strong-scales for N > 1), 5 steps after 2 warm-ups",
                     "ms_per_step": ms2, "value": tr2 / (ms2 * 1e-3), "unit": UNIT, "kernel_ms": kern2,
                     "kernel_frac_of_hbm": (2.0 * tr2 / (kern2 * 1e-3) / 1e9 / load_peaks()["hbm_gbs"]) if kern2 else None,
                     "parity_last_outcome_vs_oracle": ok2, "checksum": [c0, c1]}
        del job2
        torch.cuda.empty_cache()

    # ================================================================== the exchange step, verified in every N > 1 run (untimed):
    # row shards of z -> peer all-gather -> must equal the table every rank computed for itself, bit for bit
    exchange_info = None
    if world > 1:
        g0 = gatherer(N)
        with torch.no_grad():
            a0, a1 = scoring.row_shard(N, rank, world)
            z_sh = encoder(job.tok_shard[a0:a1].contiguous(), job.mask_shard[a0:a1].contiguous()) if job.replicated \
                else encoder(job.tok_shard, job.mask_shard)
            same = True
            for _ in range(3):   # three epochs: both alternating tables are exercised
                same &= bool(torch.equal(g0.gather(z_sh), job.z_full))
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(20):
                g0.gather(z_sh)
            ev1.record()
            torch.cuda.synchronize()
        exchange_info = {"mode": g0.mode, "fallback_reason": g0.reason,
                         "in_timed_step": not job.replicated,
                         "gathered_table_equals_locally_encoded_table": all_true(same),
                         "us_per_gather": max_over_ranks(ev0.elapsed_time(ev1) / 20) * 1e3,
                         "note": "peer = mdg_peer_allgather (NVLink push kernel + epoch flags, one launch); collective = NCCL "
                                 "all_gather_into_tensor.  configs[2]'s catalogue is one encoder wave, so its timed step "
                                 "encodes it on every rank and needs no exchange; the configs[3] leg (>= 6 GPUs) has it "
                                 "inside the timed step"}
        barrier()

    # ================================================================== strong-scaling anchor + checksum (N > 1)
    single = None
    if world > 1 and wl_key == "configs2":
        del job.out
        job.out = None
        torch.cuda.empty_cache()
        if rank == 0:
            job1 = Job(ctx, wl_key, 0, L_total, 1, 0)
            for _ in range(2):
                job1.step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                job1.step()
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / 5
            c0, c1 = job1.checksum()
            single = {"workload": f"the same {L_total} outcomes x {N} drugs on ONE GPU (rank 0, other ranks idle), "
                                  f"5 steps after 2 warm-ups", "ms_per_step": ms1,
                      "value": triples_per_step / (ms1 * 1e-3), "checksum": [c0, c1],
                      "checksum_equals_sharded_run": [c0, c1] == checksum}
            del job1
            torch.cuda.empty_cache()
        barrier()

    # ================================================================== configs[3] leg (needs >= 6 GPUs of 180 GB)
    big = None
    dry = os.environ.get("MDG_BENCH_CONFIG3_DRYRUN")   # "drugs,outcomes": exercise the leg at a reduced size (test knob)
    if (world >= 6 and args.workload == "auto") or dry:
        job.out = None
        del job
        torch.cuda.empty_cache()
        if dry:
            WORKLOADS["configs3"] = dict(name="DRY RUN of the configs[3] leg", drugs=int(dry.split(",")[0]),
                                         outcomes=int(dry.split(",")[1]))
        big = run_config3_leg(ctx, world, rank, barrier, max_over_ranks, all_true, timed)

    # ================================================================== encoder stress (configs[4]) on one GPU
    enc_block = None
    if world == 1 and rank == 0 and not args.no_encoder_block:
        enc_block = encoder_block(dev)

    if rank == 0:
        peaks = load_peaks()
        roofline = None
        if kern_ms:
            # 2 B of uint16 output per ordered triple; 2*D flop per SCORED pair, and in this layout only
            # the row > col half is scored (flops halved, bytes not) => the binding roofline is the HBM write of the
            # rank tensor: 2 B/triple / 6556 GB/s = 3.05e-13 s  vs  D flop/triple / 1362.7 TFLOP/s = 1.88e-13 s.
            out_bytes = 2.0 * per_gpu_triples
            flops = 2.0 * HIDDEN * per_gpu_triples / 2
            achieved = out_bytes / (kern_ms * 1e-3) / 1e9
            traffic = None
            prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
            if os.path.exists(prof) and wl_key == "configs1":
                try:
                    traffic = json.load(open(prof)).get("pair_score_kernel", {}).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                        "kernel": "pair_score_kernel<EPI_RANK_U16_MIRROR> (N^2 GEMM + fused rank epilogue, exact LUT table)",
                        "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": out_bytes,
                        "peak_source": peaks["src"] + ", HBM copy bandwidth",
                        "whole_step": {"ms_per_step": ms_per_step, "kernel_share": kern_ms / ms_per_step,
                                       "other_ms": ms_per_step - kern_ms,
                                       "frac": out_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                       "note": "whole step (encoder + all-gather + operand prep + GEMM 1 + GEMM 2) against the "
                                               "same HBM bound (rank 0's shard / max-over-ranks step time)"},
                        "tensor_view": {"achieved_tflops": flops / (kern_ms * 1e-3) / 1e12,
                                        "peak_tflops": peaks["tf_sustained"],
                                        "frac": flops / (kern_ms * 1e-3) / 1e12 / peaks["tf_sustained"]}}
            if memset_ms:
                roofline["write_only_context"] = {
                    "memset_ms": memset_ms, "memset_gbs": out_bytes / (memset_ms * 1e-3) / 1e9,
                    "kernel_over_memset": kern_ms / memset_ms,
                    "note": "cudaMemset of the same uint16 rank tensor: the write-only bandwidth this GPU delivers; "
                            "`peak` above is the copy (read+write) bandwidth the contract prescribes"}
            if pwl_ms:
                roofline["histogram_cdf_table_variant"] = {
                    "kernel_ms": pwl_ms, "achieved": out_bytes / (pwl_ms * 1e-3) / 1e9,
                    "frac": out_bytes / (pwl_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "max_rank_deviation": pwl_dev,
                    "note": "same kernel with RankTable(kind='pwl') (conflict-free lookup; thresholds up to "
                            "max_rank_deviation ranks from the supplied quantiles); not part of `value`"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if wl_key == "configs1" else "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(wl_key, world),
            "clocks": clocks, "roofline": roofline,
            "parity_checked": parity_ok,
            "parity": {"checked": parity_ok, "what": "every rank: all N*N uint16 ranks of its first outcome equal "
                       "oracle.quantile_rank (np.searchsorted, side='right') on the same logits in the normaliser layout "
                       "(row > col, mirrored, zero diagonal)", "checksum_all_ranks": checksum,
                       "e2e_output_equals_device_output": e2e_matches},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms,
                    "sample": (f"{int(le_t.item())} of {L_total} outcomes (pinned host output bounded to 10 GiB per rank)"
                               if int(le_t.item()) != L_total else "all outcomes"),
                    "aggregate_d2h_gbs": 2.0 * e2e_triples / (e2e_ms * 1e-3) / 1e9},
            "e2e_host_mirror": None if e2e_mirror_ms is None else {"value": e2e_triples / (e2e_mirror_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_mirror_ms,
                                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_packed, "host_threads": mirror_threads,
                                "equals_plain_e2e_output": mirror_matches,
                                "note": "the same drop-in [L,N,N] host array as `e2e`, produced by score_all_pairs_to_host("
                                        "host_mirror=True): the packed lower-triangular tiles cross PCIe (half the bytes) and "
                                        "mdg_host_mirror_tiles writes the mirror image on the host threads, overlapped with "
                                        "the next chunk's kernel and copy; data movement only, every rank is computed on the GPU. "
                                        "Opt-in: on this pool's hosts the mirror threads reach ~48 GB/s of output, less than the "
                                        "PCIe link delivers, so the plain copy (`e2e`) stays the default"},
            "e2e_packed_tiles": {"value": e2e_triples / (e2e_packed_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_packed_ms,
                                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_packed,
                                 "unpacked_equals_device_output": packed_matches,
                                 "note": "same public call with packed=True: every unordered pair's rank reaches the host once, "
                                         "as 32x32 lower-triangular tiles (half the PCIe volume); `value` counts the same ordered "
                                         "triples; rebuilding the mirrored [L,N,N] array on the host "
                                         "(decoder.unpack_packed_tiles) is NOT in this time — `e2e` above is the drop-in layout"},
            "gpu_launches": launches_per_step * steps,
        }
        if single is not None:
            line["single_gpu_same_workload"] = single
            line["parity"]["checksum_equals_single_gpu"] = single["checksum_equals_sharded_run"]
            line["strong_scaling_efficiency_vs_single_gpu_same_box"] = value / (world * single["value"])
        if c2_single is not None:
            line["config2_4096_x_963_single_gpu"] = c2_single
        if big is not None:
            line["config3_20k_x_953"] = big
        if enc_block is not None:
            line["encoder"] = enc_block
        if world > 1:
            line["exchange"] = exchange_info
        if numa is not None:
            line["host_affinity"] = {"cores_before": len(numa[0]), "cores_gpu_local": len(numa[1])}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl_key)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_config3_leg(ctx, world, rank, barrier, max_over_ranks, all_true, timed):
    """BASELINE configs[3]: 20,000 drugs x 953 outcomes on >= 6 GPUs, uint16 ranks (normaliser layout) and top-1000 per
    outcome.  Roofline per SURVEY §8d with HALVED flops (only row > col is scored) and full bytes: the bound is the HBM
    write of each rank's [L/G, N, N] uint16 slab."""
    import torch
    import torch.distributed as dist
    from madrigal_b200 import scoring
    wl = WORKLOADS["configs3"]
    N, L_total = wl["drugs"], wl["outcomes"]
    l0, l1 = scoring.outcome_shard(L_total, rank, world)
    job = Job(ctx, "configs3", l0, l1, world, rank)
    ms_rank, kern_ms = timed(job.step, 5, 2, profile=True)
    rng = np.random.default_rng(rank)
    rows = sorted(int(r) for r in rng.choice(np.arange(1, N), size=48, replace=False))
    parity_ok = all_true(job.parity_one_outcome(0, sample_rows=rows))
    chk = job.out[(l1 - l0) - 1]
    sym_ok = True
    for a in range(0, N, 4000):   # symmetric + zero diagonal on the last outcome of every rank, in row blocks
        blk = chk[a:a + 4000].view(torch.int16)
        sym_ok &= bool(torch.equal(blk, chk[:, a:a + 4000].view(torch.int16).T))
    sym_ok &= bool((torch.diagonal(chk.view(torch.int16)) == 0).all())
    sym_ok = all_true(sym_ok)
    s0, _ = job.checksum()
    ck = torch.tensor([s0], dtype=torch.int64, device=ctx["dev"])
    if world > 1:
        dist.all_reduce(ck)
    job.out = None
    torch.cuda.empty_cache()

    def step_topk():
        with torch.no_grad():
            z = job.encode()
            return scoring.top_pairs_per_outcome(z, job.W, TOPK, job.table, precision="bf16", cap=65536)

    ms_topk, _ = timed(step_topk, 3, 1)
    sc, rw, cl, st, rounds = step_topk()
    topk_ok = all_true(bool((st == 0).all()) and bool((rw > cl).all()) and bool((sc[:, 1:] <= sc[:, :-1]).all()))
    peaks = load_peaks()
    triples = float(L_total) * N * N
    per_gpu_bytes = 2.0 * (l1 - l0) * N * N
    res = None
    if rank == 0:
        bound_ms_hbm = per_gpu_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3
        bound_ms_tensor = (HIDDEN * (l1 - l0) * float(N) * N) / (peaks["tf_sustained"] * 1e12) * 1e3   # 2*D flop / 2
        res = {"workload": f"{wl['name']}: {N} drugs x {L_total} outcomes on {world} GPUs (outcome shards of <= {l1 - l0}), "
                           f"encoder row shards -> all-gather of z -> uint16 ranks in the normaliser layout "
                           f"({per_gpu_bytes / 1e9:.1f} GB per GPU)",
               "rank_u16": {"ms_per_step": ms_rank, "value": triples / (ms_rank * 1e-3), "unit": UNIT, "steps": 5, "warmup": 2,
                            "kernel_ms_rank0": kern_ms},
               "roofline": {"bound": "hbm", "bound_ms": max(bound_ms_hbm, bound_ms_tensor),
                            "hbm_bound_ms": bound_ms_hbm, "tensor_bound_ms_halved_flops": bound_ms_tensor,
                            "frac": max(bound_ms_hbm, bound_ms_tensor) / ms_rank,
                            "kernel_frac_rank0": (max(bound_ms_hbm, bound_ms_tensor) / kern_ms) if kern_ms else None,
                            "note": "SURVEY §8d units: 2 B per ordered triple written, D flop per ordered triple "
                                    "(only row > col is scored and mirrored): the slower of the two bounds per GPU; "
                                    "north-star target >= 0.60"},
               "top1000": {"ms_per_step": ms_topk, "value": triples / (ms_topk * 1e-3), "unit": UNIT, "steps": 3,
                           "rounds": int(rounds), "all_lists_complete_sorted_lower_triangle": topk_ok},
               "parity_checked": bool(parity_ok and sym_ok),
               "parity": {"oracle_rows": "every rank: 48 random rows of its first outcome bit for bit against "
                                         "oracle.quantile_rank on the same logits (row > col), zero diagonal, mirrored",
                          "oracle_rows_ok": parity_ok, "symmetric_zero_diagonal_last_outcome": sym_ok,
                          "checksum_all_ranks": int(ck.item())}}
    del job
    torch.cuda.empty_cache()
    barrier()
    return res


def encoder_block(dev):
    """BASELINE configs[4] (fusion-encoder stress, 1M drug views, random missing-modality masks) and the reference's
    shipped production shapes, encoder only, on one GPU: drugs/s and fraction of the sustained bf16 tensor peak with the
    flop count of SURVEY §8d."""
    import torch
    import madrigal_b200 as mb
    import synth
    peaks = load_peaks()

    def flops_per_drug(T, E, Dl, F, layers, agg):
        per_tok = layers * (8 * Dl * Dl + 4 * Dl * F) + 2 * E * Dl
        pool = 4 * Dl * Dl * T + 2 * Dl * Dl + 2 * Dl * E if agg == "x-attn" else 2 * Dl * E * T
        return T * per_tok + pool

    cases = [("configs4_hidden256", 1 << 20, 4, 128, 8, 32, 512, "mean", 0),
             ("configs4_hidden512", 1 << 19, 4, 128, 8, 64, 1024, "mean", 0),
             ("production_drugbank_T23", 1 << 15, 23, 128, 8, 64, 256, "x-attn", 4),
             ("production_twosides_T21", 1 << 15, 21, 128, 2, 256, 512, "x-attn", 2)]
    out = {}
    for name, B, T, E, H, hd, F, agg, nb in cases:
        cfg = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True, agg=agg, nb=nb)
        enc = mb.TransformerFusion(E, nb, 2, H, hd, F, transformer_actn="gelu", transformer_norm_first=True,
                                   transformer_batch_first=False, transformer_agg=agg, precision="bf16")
        enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(cfg, 1).items()})
        if agg == "x-attn" and nb == 0:
            enc.x_attn_key_padding_mask = torch.zeros(1, T, dtype=torch.bool)
        enc = enc.to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(0)
        tokens = torch.randn(B, T, E, device=dev, generator=g)
        mask = torch.rand(B, T, device=dev, generator=g) < 0.5
        mask[:, 0] = False
        with torch.no_grad():
            for _ in range(2):
                enc(tokens, mask)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                enc(tokens, mask)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        tf = flops_per_drug(T, E, H * hd, F, 2, agg) * B / (ms * 1e-3) / 1e12
        out[name] = {"drugs": B, "tokens": T, "latent": H * hd, "heads": H, "ffn": F, "agg": agg, "ms": ms,
                     "drugs_per_s": B / (ms * 1e-3), "tflops": tf, "frac_of_sustained_bf16": tf / peaks["tf_sustained"],
                     "launches": enc.last_launch_count}
        del enc, tokens, mask
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "configs1", "configs2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encoder-block", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
