"""Quick device-side timing of mdg_pair_score variants (CUDA events, L2 flushed between iterations)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb  # noqa: E402
from synth import decoder_inputs  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def case(N, D, L, precision, mode, Q=16384, symmetric=False):
    z, W = decoder_inputs(N, D, L, 0)
    zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
    table = None
    if mode == "rank":
        # quantiles from a 1024-drug panel of the same catalogue, outcome by outcome (setup, untimed)
        qs = []
        for l in range(L):
            lg = mb.pair_score(zt[:1024], zt[:1024], Wt[l:l + 1], precision=precision, out="logit")[0]
            i, j = torch.tril_indices(1024, 1024, -1, device=dev)
            v = lg[i, j].sort().values
            M = v.numel()
            idx = (torch.arange(1, Q + 1, device=dev, dtype=torch.int64) * M + Q - 1) // Q - 1
            qs.append(v[idx])
        table = mb.RankTable(torch.stack(qs))
    out = torch.empty((L, N, N), dtype=torch.uint16 if mode == "rank" else torch.float32, device=dev)
    fn = lambda: mb.pair_score(zt, zt, Wt, precision=precision, out=mode, table=table, out_tensor=out,
                               symmetric=symmetric)
    med, best = timeit(fn)
    triples = L * N * N
    flops = 2.0 * D * triples * (3 if precision == "fp32" else 1)
    obytes = triples * (2 if mode == "rank" else 4)
    print(f"N={N} D={D} L={L} {precision} {mode}{' symmetric' if symmetric else ''}: {med:.3f} ms (best {best:.3f}) -> {triples / med / 1e9:.1f} G triples/s, "
          f"{flops / med / 1e9:.0f} TFLOP/s (incl. split terms), out {obytes / med / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    case(4096, 256, 86, "bf16", "rank")
    case(4096, 256, 86, "bf16", "rank", symmetric=True)
    case(4096, 128, 86, "bf16", "rank")
    case(4096, 128, 86, "bf16", "rank", symmetric=True)
    case(8192, 256, 32, "bf16", "rank", symmetric=True)
    case(4096, 256, 86, "bf16", "logit")
    case(4096, 256, 86, "fp32", "logit")
    case(4096, 256, 86, "bf16", "sigmoid")
    case(8192, 256, 32, "bf16", "rank")
