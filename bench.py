#!/usr/bin/env python
"""Benchmark of the drug-pair scoring path (BASELINE.json metric: scored (outcome, drugA, drugB) triples/sec, fused
rank).  Contract: one JSON line on stdout from rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N = 1): BASELINE.json configs[1] — 4,096 drugs x 86 outcomes, hidden 256, all-pairs fused scoring + uint16
quantile rank on one B200.  N > 1: the same per-GPU workload with DISTINCT outcomes per rank (weak scaling, outcomes
sharded); the only collective is the all-gather of the fused-embedding table.

A "step" = one pass of the hot path over the whole batch: [all-gather z] -> operand prep -> GEMM 1 (z.W_l) -> GEMM 2 +
fused rank epilogue writing uint16 ranks for every (outcome, drugA, drugB).
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

N_DRUGS = 4096
N_OUTCOMES = 86
HIDDEN = 256
N_TOKENS = 4          # BASELINE: "up to four modality tokens per drug (structure, KG, transcriptomic, cell-viability)"
ENC = dict(embed_dim=HIDDEN, num_layers=2, num_heads=8, head_dim=32, ffn_dim=512, actn="gelu", norm_first=True,
           agg="x-attn", nb=0)  # latent 256 = 8 x 32, FFN 2 x latent, pre-LN + GELU + x-attn pooling as shipped configs
Q_TABLE = 16384
PANEL = 2048
METRIC = "scored (outcome, drugA, drugB) triples/sec, fused rank"
UNIT = "triples/s"


def workload_config(n_gpus):
    return {
        "workload": f"BASELINE configs[1]: {N_DRUGS} drugs x {N_OUTCOMES} outcomes per GPU, hidden {HIDDEN}: fusion "
                    f"encoder ({N_TOKENS} modality tokens/drug, random missing-modality masks, 2 layers, 8 heads, "
                    f"latent 256, FFN 512, x-attn pooling) -> all-pairs bf16-input/fp32-accumulate bilinear scoring "
                    f"-> fused uint16 rank (Q={Q_TABLE} reference quantiles/outcome from a {PANEL}-drug panel)",
        "drugs": N_DRUGS, "outcomes_per_gpu": N_OUTCOMES, "outcomes_total": N_OUTCOMES * n_gpus, "hidden": HIDDEN,
        "pairs": "one catalogue scored against itself in the reference normaliser's layout (notebooks/normalize_scores.py:"
                 "67-70): each unordered pair (row > col) is scored and ranked once and its rank written at [l,i,j] and "
                 "[l,j,i], diagonal 0; `value` counts the L*N*N uint16 entries written (ordered triples)",
        "parallelism": (f"drugs sharded over {n_gpus} GPUs for the encoder, one all-gather of z per step, outcomes "
                        f"sharded for the decoder") if n_gpus > 1 else "1 GPU",
        "l2": "no explicit flush: each step streams 2.9 GB of output through the 126 MB L2, evicting the inputs",
    }


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm for this path on the host cores (oracle port; the Python reference cannot
# travel to the GPU box).  decoder = fp32 matmul, normaliser = run_slice per outcome in a multiprocessing.Pool()
# exactly as notebooks/normalize_scores.py:78-85.
# ----------------------------------------------------------------------------------------------------------------
_CPU_STATE = {}


def _cpu_worker_init(z, W):
    _CPU_STATE["z"], _CPU_STATE["W"] = z, W


def _cpu_one_outcome(l):
    from oracle import oracle
    z, W = _CPU_STATE["z"], _CPU_STATE["W"]
    raw = oracle.bilinear_scores(z, z, W, (l, l + 1))          # models.py:537-547
    norm = oracle.normalize_scores(raw)                         # normalize_scores.py:62-74
    return float(norm[0, 1, 0])


def cpu_reference_pass(n_outcomes, cores):
    """Score + rank-normalise `n_outcomes` outcomes of the 4,096-drug workload on `cores` processes; seconds."""
    import multiprocessing as mp
    import synth
    from oracle import oracle
    tokens, masks = synth.fusion_inputs(N_DRUGS, N_TOKENS, HIDDEN, seed=0)
    sd = synth.fusion_state_dict(ENC, seed=7)
    _, W = synth.decoder_inputs(1, HIDDEN, n_outcomes, seed=100)
    t0 = time.perf_counter()
    # fusion encoder for the whole catalogue (TransformerFusion.forward restated, fp32, BLAS threads) ...
    z = oracle.fusion_forward(sd, ENC, tokens, masks, None, np.zeros(N_TOKENS, bool))
    # ... then decoder + exact rank normalisation, one outcome per process
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=cores, initializer=_cpu_worker_init, initargs=(z, W)) as pool:
        pool.map(_cpu_one_outcome, range(n_outcomes))
    return time.perf_counter() - t0


def cpu_baseline(cores=None):
    cores = cores or os.cpu_count() or 1
    n_out = max(1, min(4 * cores, N_OUTCOMES))  # four outcomes per worker: ~10 s wall on the box's 16 cores
    dt = cpu_reference_pass(n_out, cores)
    return {"value": n_out * N_DRUGS * N_DRUGS / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_out} of {N_OUTCOMES} outcomes x {N_DRUGS}^2 pairs: fp32 fusion encoder for all {N_DRUGS} drugs "
                      f"+ fp32 decoder (numpy matmul) + the reference's exact argsort rank normaliser, one outcome per "
                      f"process in Pool({cores}); {dt:.1f} s"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_out = max(1, min(2 * cores, N_OUTCOMES))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_pass(1, 1)
    times = [cpu_reference_pass(n_out, cores) for _ in range(args.steps)]
    dt = float(np.mean(times))
    value = n_out * N_DRUGS * N_DRUGS / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = {n_out} of {N_OUTCOMES} outcomes x {N_DRUGS}^2 pairs (decoder + exact "
                                   f"rank normaliser, Pool({cores})); the reference is Python and cannot be installed "
                                   f"on the GPU box, so this is the oracle port of its algorithm"},
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


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import madrigal_b200 as mb
    from madrigal_b200 import _lib, normalize, scoring
    from synth import decoder_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    # one process per GPU: run on the cores local to this GPU so that pinned buffers land on its NUMA node
    numa = scoring.bind_host_thread_to_gpu(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().mdg_check_device(local_rank), "mdg_check_device")

    # ---- synthetic inputs (seeded): shared drug catalogue (modality tokens + masks), per-rank outcomes
    import synth
    tok_np, mask_np = synth.fusion_inputs(N_DRUGS, N_TOKENS, HIDDEN, seed=0)
    _, W_np = decoder_inputs(1, HIDDEN, N_OUTCOMES, seed=100 + rank)
    encoder = mb.TransformerFusion(HIDDEN, 0, ENC["num_layers"], ENC["num_heads"], ENC["head_dim"], ENC["ffn_dim"],
                                   transformer_actn=ENC["actn"], transformer_norm_first=True,
                                   transformer_batch_first=False, transformer_agg="x-attn", precision="bf16")
    encoder.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(ENC, seed=7).items()})
    encoder.x_attn_key_padding_mask = torch.zeros(1, N_TOKENS, dtype=torch.bool)  # nb = 0: every token is a key
    encoder = encoder.to(dev).eval()
    tokens = torch.from_numpy(tok_np).to(dev)
    masks = torch.from_numpy(mask_np).to(dev)
    W = torch.from_numpy(W_np).to(dev)
    r0, r1 = scoring.row_shard(N_DRUGS, rank, world)
    tok_shard, mask_shard = tokens[r0:r1].contiguous(), masks[r0:r1].contiguous()
    with torch.no_grad():
        z_full = encoder(tokens, masks)
    # setup (untimed): per-outcome reference quantiles from a drug panel -> prepared rank table
    table_quantiles = normalize.build_reference_quantiles(z_full, W, Q_TABLE, panel=PANEL, precision="bf16")
    table = mb.RankTable(table_quantiles)   # exact bucket LUT (the headline configuration)
    out = torch.empty((N_OUTCOMES, N_DRUGS, N_DRUGS), dtype=torch.uint16, device=dev)
    torch.cuda.synchronize()
    launches = {"n": 0}

    @torch.no_grad()
    def step():
        z = encoder(tok_shard, mask_shard)                       # fusion encoder on this rank's drugs
        launches["n"] = encoder.last_launch_count
        if world > 1:
            z = scoring.all_gather_embeddings(z, N_DRUGS)        # the path's only collective
            launches["n"] += 1
        mb.pair_score(z, z, W, precision="bf16", out="rank", table=table, out_tensor=out, symmetric=True)
        launches["n"] += _lib.lib().mdg_last_launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches_per_step = launches["n"]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    _lib.check(_lib.lib().mdg_profile_enable(min(args.steps, 256)), "mdg_profile_enable")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    buf = (ctypes.c_float * 256)()
    n_rec = _lib.lib().mdg_profile_read(buf, 256)
    _lib.lib().mdg_profile_enable(0)
    kern_ms = float(np.mean(buf[:n_rec])) if n_rec > 0 else None
    clocks = sampler.stop() if rank == 0 else None

    # ---- context (untimed region): what a plain write of the same tensor costs on this GPU (the kernel's traffic is
    #      ~all writes; a copy, the contract's HBM denominator, moves half its bytes as reads)
    memset_ms = None
    if rank == 0:
        flat = out.view(torch.uint8).reshape(-1)
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

    # ---- secondary measurement (untimed region): the same kernel with the histogram-CDF rank table (MDG_RANK_PWL)
    pwl_ms = None
    if rank == 0:
        table_pwl = mb.RankTable(table_quantiles, kind="pwl")
        for _ in range(3):
            mb.pair_score(z_full, z_full, W, precision="bf16", out="rank", table=table_pwl, out_tensor=out, symmetric=True)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().mdg_profile_enable(10), "mdg_profile_enable")
        for _ in range(10):
            mb.pair_score(z_full, z_full, W, precision="bf16", out="rank", table=table_pwl, out_tensor=out, symmetric=True)
        torch.cuda.synchronize()
        n_pwl = _lib.lib().mdg_profile_read(buf, 256)
        _lib.lib().mdg_profile_enable(0)
        pwl_ms = float(np.mean(buf[:n_pwl])) if n_pwl > 0 else None
        pwl_dev = float(table_pwl.max_rank_deviation.max().item())
        del table_pwl

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / args.steps
    triples_per_step = world * N_OUTCOMES * N_DRUGS * N_DRUGS
    value = triples_per_step / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, host buffers out, through the public scoring driver
    e2e_steps = max(1, min(args.steps, 3))
    tok_host = torch.from_numpy(tok_np[r0:r1]).pin_memory()
    mask_host = torch.from_numpy(mask_np[r0:r1]).pin_memory()
    W_host = torch.from_numpy(W_np).pin_memory()
    out_host = torch.empty((N_OUTCOMES, N_DRUGS, N_DRUGS), dtype=torch.uint16).pin_memory()

    @torch.no_grad()
    def e2e_step():
        zd = encoder(tok_host.to(dev, non_blocking=True), mask_host.to(dev, non_blocking=True))
        Wd = W_host.to(dev, non_blocking=True)
        if world > 1:
            zd = scoring.all_gather_embeddings(zd, N_DRUGS)
        scoring.score_all_pairs_to_host(zd, Wd, out_host, out="rank", table=table, precision="bf16", chunk=10,
                                        symmetric=True)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = triples_per_step / (te.item() * 1e-3)
    h2d = tok_host.numel() * 4 + mask_host.numel() + W_host.numel() * 4
    d2h = out_host.numel() * 2

    if rank == 0:
        peaks = load_peaks()
        per_gpu_triples = N_OUTCOMES * N_DRUGS * N_DRUGS
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
            if os.path.exists(prof):
                try:
                    traffic = json.load(open(prof)).get("pair_score_kernel", {}).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                        "kernel": "pair_score_kernel<EPI_RANK_U16_MIRROR> (N^2 GEMM + fused rank epilogue)",
                        "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": out_bytes,
                        "peak_source": peaks["src"] + ", HBM copy bandwidth",
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
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks, "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": te.item()},
            "gpu_launches": launches_per_step * args.steps,
        }
        if numa is not None:
            line["host_affinity"] = {"cores_before": len(numa[0]), "cores_gpu_local": len(numa[1])}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
