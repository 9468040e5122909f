import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from synth import decoder_inputs
dev = torch.device("cuda:0")
z, W = decoder_inputs(4096, 256, 86, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
thr = torch.full((86,), 0.5, device=dev)
for _ in range(3):
    mb.pair_topk(zt, zt, Wt, thr, 100, cap=4096, symmetric=False, precision="bf16")
torch.cuda.synchronize()
print("ok")
