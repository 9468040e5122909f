import os, sys, torch, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import madrigal_b200 as mb
from synth import decoder_inputs
dev = torch.device("cuda:0")
N, D, L = 4096, 256, 86
z, W = decoder_inputs(N, D, L, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
pd = mb.PreparedDecoder(Wt, "bf16")
lab = torch.zeros(8, dtype=torch.int32, device=dev)
# the gather entry point runs convert_z + GEMM 1 + a tiny gather kernel: times GEMM 1 in isolation
fn = lambda: mb.pair_score_gather(zt, zt, Wt, lab, lab, lab, precision="bf16")
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): fn()
e1.record(); torch.cuda.synchronize()
print(os.environ.get("MDG_DEBUG_GEMM1_NOSTORE", "stores on"), "convert_w + convert_z + GEMM1 + gather: %.1f us per call" % (e0.elapsed_time(e1) / 50 * 1e3))
