"""Large-output runs of the pair kernel's modes (is the ~3.7 TB/s of the config-4 rank slice a DRAM write ceiling?)."""
import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize, _lib
from synth import decoder_inputs
dev = torch.device("cuda:0")
def kms(fn, iters=3):
    fn(); torch.cuda.synchronize()
    _lib.lib().mdg_profile_enable(iters)
    for _ in range(iters): fn()
    torch.cuda.synchronize()
    buf = (ctypes.c_float * 16)(); n = _lib.lib().mdg_profile_read(buf, 16); _lib.lib().mdg_profile_enable(0)
    return float(np.mean(buf[:n]))
N, D, L = 16384, 256, 32
z, W = decoder_inputs(N, D, L, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
out32 = torch.empty((L, N, N), dtype=torch.float32, device=dev)
ms = kms(lambda: mb.pair_score(zt, zt, Wt, precision="bf16", out="logit", out_tensor=out32))
print(f"fp32 logits  {out32.numel()*4/1e9:.1f} GB: {ms:.2f} ms -> {out32.numel()*4/ms/1e6:.0f} GB/s")
x = out32.view(torch.uint8).reshape(-1)
for _ in range(2): x.zero_()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); x.zero_(); e.record(); torch.cuda.synchronize()
print(f"memset       {x.numel()/1e9:.1f} GB: {s.elapsed_time(e):.2f} ms -> {x.numel()/s.elapsed_time(e)/1e6:.0f} GB/s")
del out32, x
out16 = torch.empty((L, N, N), dtype=torch.uint16, device=dev)
for kind in ("lut", "pwl"):
    table = normalize.build_rank_table(zt, Wt, 16384, kind=kind, panel=2048, precision="bf16")
    for sym in (True, False):
        ms = kms(lambda: mb.pair_score(zt, zt, Wt, precision="bf16", out="rank", table=table, out_tensor=out16, symmetric=sym))
        print(f"rank {kind} sym={sym} {out16.numel()*2/1e9:.1f} GB: {ms:.2f} ms -> {out16.numel()*2/ms/1e6:.0f} GB/s")
