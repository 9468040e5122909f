import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize
from synth import decoder_inputs
dev = torch.device("cuda:0")
N, L, Q, k = 4096, 86, 16384, 1000
z, W = decoder_inputs(N, 256, L, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
quant = normalize.build_reference_quantiles(zt, Wt, Q, panel=2048, precision="bf16")
qi = min(Q - 1, max(0, int(Q * (1.0 - 3.0 * k / (N * N))) - 1))
thr = quant[:, qi].contiguous()
for _ in range(3):
    out = mb.pair_topk(zt, zt, Wt, thr, k, cap=8192, symmetric=False, precision="bf16")
torch.cuda.synchronize()
print("ok", int((out[3] == 0).sum()))
