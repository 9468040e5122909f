import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize
from synth import decoder_inputs
dev = torch.device("cuda:0")
N, D, L = int(os.environ.get("BIG_N", "16384")), 256, int(os.environ.get("BIG_L", "8"))
z, W = decoder_inputs(N, D, L, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
table = normalize.build_rank_table(zt, Wt, 16384, kind=os.environ.get("KIND", "lut"), panel=2048, precision="bf16")
out = torch.empty((L, N, N), dtype=torch.uint16, device=dev)
for _ in range(3):
    mb.pair_score(zt, zt, Wt, precision="bf16", out="rank", table=table, out_tensor=out, symmetric=True)
torch.cuda.synchronize(); print("ok")
