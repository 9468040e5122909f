"""Time the symmetric (normaliser-layout) fused rank kernel; prints a checksum so that variants can be compared."""
import os, sys, ctypes
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize, _lib
from synth import decoder_inputs
dev = torch.device("cuda:0")
N, D, L = int(os.environ.get("N", 4096)), int(os.environ.get("D", 256)), int(os.environ.get("L", 86))
z, W = decoder_inputs(N, D, L, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
kind = os.environ.get("KIND", "lut")
table = normalize.build_rank_table(zt, Wt, 16384, kind=kind, panel=2048, precision="bf16")
if kind == "pwl": print("max rank deviation per outcome: max", table.max_rank_deviation.max().item(), "mean", table.max_rank_deviation.mean().item())
sym = os.environ.get("SYM", "1") == "1"
packed = os.environ.get("PACKED", "0") == "1"
from madrigal_b200.decoder import packed_tiles_per_outcome
out = torch.empty((L, packed_tiles_per_outcome(N), 32, 32) if packed else (L, N, N), dtype=torch.uint16, device=dev)
pd = mb.PreparedDecoder(Wt, "bf16")
fn = lambda: mb.pair_score(zt, zt, pd, out="rank", table=table, out_tensor=out, symmetric=sym, packed=packed)
for _ in range(3): fn()
torch.cuda.synchronize()
_lib.lib().mdg_profile_enable(20)
for _ in range(20): fn()
torch.cuda.synchronize()
buf = (ctypes.c_float * 256)(); n = _lib.lib().mdg_profile_read(buf, 256)
ms = np.array(buf[:n])
chk = sum(int(out[l].view(torch.int16).sum(dtype=torch.int64).item()) for l in range(L))
print(f"epi={os.environ.get('MDG_MIRROR_EPI', 'default')} packed={packed} kind={kind} sym={sym} N={N} D={D} L={L}: kernel {ms.mean():.4f} ms (min {ms.min():.4f}) -> "
      f"{out.numel()*2.0/ms.mean()/1e6:.0f} GB/s out ({L*N*N/ms.mean()/1e9:.3f} T triples/s), checksum {chk}")
