"""Timing of the top-k output mode (GEMM 2 with a compare-only epilogue: shows the tcgen05 pipeline's own speed)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize
from synth import decoder_inputs
dev = torch.device("cuda:0")

def case(N, D, L, symmetric, k=1000, Q=16384):
    z, W = decoder_inputs(N, D, L, 0)
    zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
    quant = normalize.build_reference_quantiles(zt, Wt, Q, panel=2048, precision="bf16")
    M = N * (N - 1) // 2 if symmetric else N * N
    qi = min(Q - 1, max(0, int(Q * (1.0 - 3.0 * k / M)) - 1))    # ~3k candidates per outcome
    thr = (quant[:, qi] + float(os.environ.get('THR_SHIFT', '0'))).contiguous()
    cap = int(os.environ.get('CAP', 8192))
    import ctypes
    from madrigal_b200 import _lib
    fn = lambda: mb.pair_topk(zt, zt, Wt, thr, k, cap=cap, symmetric=symmetric, precision="bf16")
    for _ in range(2): out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ms = float(np.median(ts))
    _lib.lib().mdg_profile_enable(8)
    for _ in range(5): fn()
    buf = (ctypes.c_float * 8)(); n = _lib.lib().mdg_profile_read(buf, 8); _lib.lib().mdg_profile_enable(0)
    kms = float(np.median(buf[:n]))
    st = out[3].cpu().numpy()
    pairs = L * N * N
    flops = 2.0 * D * pairs * (0.5 if symmetric else 1.0)
    print(f"topk N={N} D={D} L={L} symmetric={symmetric}: call {ms:.3f} ms, GEMM2 kernel {kms:.3f} ms -> "
          f"{flops / kms / 1e9:.0f} TFLOP/s (algorithmic flops), {pairs / ms / 1e9:.2f} T ordered triples/s covered by the call, "
          f"status ok={int((st == 0).sum())}/{L} overflow={int((st == 2).sum())}", flush=True)

if __name__ == "__main__":
    case(4096, 256, 86, False)
    case(4096, 256, 86, True)
    case(4096, 128, 86, False)
    case(8192, 256, 64, True)
