"""pair_topk on the config-4 slice (20,000 drugs x 119 outcomes, top-1000, cap 65536) for an ncu launch list."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize
from synth import decoder_inputs
dev = torch.device("cuda:0")
N, L, Q, k = 20000, 119, 16384, 1000
z, W = decoder_inputs(N, 256, L, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
table = normalize.build_rank_table(zt, Wt, Q, panel=2048, precision="bf16")
thr = table.thresholds[:, Q - 2].contiguous()
for _ in range(2):
    out = mb.pair_topk(zt, zt, Wt, thr, k, cap=65536, symmetric=True, precision="bf16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = mb.pair_topk(zt, zt, Wt, thr, k, cap=65536, symmetric=True, precision="bf16"); e1.record()
torch.cuda.synchronize()
print("ok", int((out[3] == 0).sum()), "call ms", e0.elapsed_time(e1))
