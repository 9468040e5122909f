"""mdg_exact_rank at the reference's published size (11,607 drugs), 2 outcomes, for an ncu launch list."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from madrigal_b200 import normalize
dev = torch.device("cuda:0")
N, L = 11607, 2
s = torch.randn(L, N, N, device=dev)
for _ in range(2):
    out = normalize.exact_normalized_ranks(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = normalize.exact_normalized_ranks(s); e1.record(); torch.cuda.synchronize()
print("ok ms per outcome", e0.elapsed_time(e1) / L)
