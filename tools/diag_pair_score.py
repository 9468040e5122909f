"""GPU diagnostic battery for mdg_pair_score (prints a compact report; exits non-zero on any failure)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb  # noqa: E402
from oracle import oracle  # noqa: E402
from synth import decoder_inputs  # noqa: E402

dev = torch.device("cuda:0")
fails = 0


def report(name, got, ref, tol):
    global fails
    ref64 = ref.astype(np.float64)
    err = np.abs(got.astype(np.float64) - ref64)
    rms = np.sqrt(np.mean(ref64 ** 2))
    bound = tol * np.maximum(np.abs(ref64), rms)
    bad = err > bound
    worst = np.unravel_index(np.argmax(err / bound), err.shape)
    ok = not bad.any() and np.isfinite(got).all()
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max|err|={err.max():.3e} rms(ref)={rms:.3e} max(err/bound)={(err / bound).max():.3f} "
          f"worst={worst} got={got[worst]:.6f} ref={ref[worst]:.6f} nbad={int(bad.sum())}/{bad.size}")
    if not ok:
        fails += 1
        # error map per 32x32 block of the first outcome with errors
        l = worst[0]
        e = (err[l] / bound[l])
        n0, n1 = e.shape
        bs = 32
        print("    block map (max err/bound per 32x32 block, outcome %d; '.'<=1, digit=log10):" % l)
        for i in range(0, min(n0, 512), bs):
            row = ""
            for j in range(0, min(n1, 512), bs):
                m = e[i:i + bs, j:j + bs].max()
                row += "." if m <= 1 else str(min(9, int(np.log10(m)) + 1))
            print("    " + row)
    return ok


def run_case(N1, N2, D, L, precision, tol, seed=0, mode="logit"):
    z1, W = decoder_inputs(N1, D, L, seed)
    z2, _ = decoder_inputs(N2, D, 1, seed + 1)
    ref = oracle.bilinear_scores(z1, z2, W, dtype=np.float64)
    t0 = time.time()
    out = mb.pair_score(torch.from_numpy(z1).to(dev), torch.from_numpy(z2).to(dev), torch.from_numpy(W).to(dev),
                        precision=precision, out=mode)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    if mode == "sigmoid":
        ref = 1.0 / (1.0 + np.exp(-ref))
    return report(f"pair_score N1={N1} N2={N2} D={D} L={L} {precision} {mode} ({time.time() - t0:.2f}s)", got, ref, tol)


def run_rank_case(N, D, L, Q, seed=0, precision="bf16"):
    global fails
    z, W = decoder_inputs(N, D, L, seed)
    zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
    logits = mb.pair_score(zt, zt, Wt, precision=precision, out="logit")
    torch.cuda.synchronize()
    lg = logits.cpu().numpy()
    quant = oracle.reference_quantiles(lg, Q)
    table = mb.RankTable(torch.from_numpy(quant).to(dev))
    torch.cuda.synchronize()
    thr = table.thresholds.cpu().numpy()
    # table sanity
    asc = bool((np.diff(thr, axis=1) >= 0).all())
    rng_ = (quant[:, -1] - quant[:, 0])
    snap = np.abs(thr - quant).max(axis=1) / np.maximum(rng_, 1e-30)
    print(f"    table: ascending={asc} max snap error / range = {snap.max():.3e} (cells: {snap.max() * 131072 / 1.02:.2f})")
    # stand-alone lookup vs searchsorted
    r_lookup = table.lookup(logits).cpu().numpy()
    r_ref = oracle.quantile_rank(thr, lg, side="right")
    n_bad_lookup = int((r_lookup != r_ref).sum())
    # fused epilogue vs searchsorted on identical logits
    r_fused = mb.pair_score(zt, zt, Wt, precision=precision, out="rank", table=table)
    torch.cuda.synchronize()
    r_fused = r_fused.cpu().numpy()
    n_bad_fused = int((r_fused != r_ref).sum())
    ok = asc and n_bad_lookup == 0 and n_bad_fused == 0
    print(f"[{'OK ' if ok else 'BAD'}] rank N={N} D={D} L={L} Q={Q} {precision}: lookup mismatches={n_bad_lookup} fused mismatches={n_bad_fused} "
          f"of {r_ref.size}; rank range [{r_ref.min()}, {r_ref.max()}]")
    if not ok:
        fails += 1
        if n_bad_fused:
            idx = np.argwhere(r_fused != r_ref)[:5]
            for (l, i, j) in idx:
                print(f"      l={l} i={i} j={j} logit={lg[l, i, j]!r} fused={r_fused[l, i, j]} ref={r_ref[l, i, j]} lookup={r_lookup[l, i, j]}")
    return ok


if __name__ == "__main__":
    print("device:", torch.cuda.get_device_name(0))
    quick = "--quick" in sys.argv
    os.environ.pop("MDG_FORCE_DIRECT_STORE", None)
    run_case(128, 128, 64, 1, "bf16", 1e-2)
    run_case(128, 128, 64, 1, "fp32", 1e-3)
    run_case(256, 256, 128, 2, "bf16", 1e-2)
    run_case(256, 384, 256, 3, "bf16", 1e-2)
    run_case(256, 384, 256, 3, "fp32", 1e-3)
    run_case(200, 328, 128, 2, "fp32", 1e-3)       # ragged, TMA-store-able (328*4 % 16 == 0)
    run_case(130, 75, 192, 2, "fp32", 1e-3)        # ragged, direct-store path
    run_case(130, 75, 192, 2, "bf16", 1e-2)
    run_case(1, 1, 64, 1, "fp32", 1e-3)
    run_case(300, 520, 256, 2, "fp32", 1e-3, mode="sigmoid")
    os.environ["MDG_FORCE_DIRECT_STORE"] = "1"
    run_case(256, 384, 256, 3, "fp32", 1e-3)
    os.environ.pop("MDG_FORCE_DIRECT_STORE")
    run_rank_case(192, 128, 2, 1024)
    run_rank_case(256, 256, 3, 16384)
    run_rank_case(64, 64, 2, 2016)                  # Q == M: exact in-sample ranks
    if not quick:
        run_case(1024, 1024, 128, 86, "fp32", 1e-3)   # BASELINE config 1
        run_case(1024, 1024, 256, 86, "bf16", 1e-2)
        run_rank_case(1024, 256, 8, 16384, precision="bf16")
    print("FAILS:", fails)
    sys.exit(1 if fails else 0)
