"""The reference's ensemble normalisation at its published size (generate_embeddings.ipynb: 5 checkpoints, 158 outcomes,
11,607 drugs: 5 x 837.66 s per-checkpoint normalisation + 3,015.09 s gmean + 837.66 s re-normalisation), streamed by
outcome chunk through scoring.ensemble_normalized_ranks_chunks on one GPU.  OUTCOMES env limits the outcome count."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import madrigal_b200 as mb
from madrigal_b200 import scoring

dev = torch.device("cuda:0")
N, D, K = 11607, 128, 5
L = int(os.environ.get("OUTCOMES", "158"))
g = torch.Generator(device=dev).manual_seed(0)
zs = [torch.randn(N, D, device=dev, generator=g) / D ** 0.5 for _ in range(K)]
Ws = []
for _ in range(K):
    P = (torch.rand(L, D, D, device=dev, generator=g) * 2 - 1) / D ** 0.5
    Ws.append((torch.triu(P) + torch.triu(P, 1).transpose(1, 2)).contiguous())
for _ in scoring.ensemble_normalized_ranks_chunks(zs, [W[:2] for W in Ws], chunk=2):  # warm-up
    pass
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); e0.record()
chk = 0.0
for l0, l1, r in scoring.ensemble_normalized_ranks_chunks(zs, Ws, chunk=2):
    chk += float(r[0, 1, 0])  # touch the result
e1.record(); torch.cuda.synchronize()
dev_s, wall_s = e0.elapsed_time(e1) / 1e3, time.time() - t0
res = {"what": "ensemble normalisation, 5 checkpoints x %d outcomes x 11607^2" % L, "device_s": dev_s, "wall_s": wall_s,
       "per_outcome_ms": 1e3 * dev_s / L, "reference_s_published_158_outcomes": 5 * 837.6621 + 3015.0882 + 837.6621,
       "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "checksum": chk}
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ensemble_timing.json"), "w"))
