"""Quantile-formulation ensemble at BASELINE configs[1] size: K = 5 checkpoints x 86 outcomes x 4096 drugs, streamed by
outcome chunk (5 fused-rank kernels + mdg_ensemble_rank_u16 per chunk) vs the reference-faithful exact chain
(fp32 logits -> exact in-sample rank per member -> gmean -> exact re-rank) on a few outcomes."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize, scoring
import synth
dev = torch.device("cuda:0")
N, D, L, K, Q, PANEL = 4096, 256, 86, 5, 16384, 2048
zs, Ws, pds, tables = [], [], [], []
for k in range(K):
    z, W = synth.decoder_inputs(N, D, L, seed=10 + k)
    zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
    zs.append(zt); Ws.append(Wt); pds.append(mb.PreparedDecoder(Wt, "bf16"))
    tables.append(normalize.build_rank_table(zt, Wt, Q, panel=PANEL, precision="bf16"))
# ensemble table from the members' ranks over the panel (setup)
ens_quant = []
for l0 in range(0, L, 8):
    l1 = min(l0 + 8, L)
    members = [mb.pair_score(z[:PANEL].contiguous(), z[:PANEL].contiguous(), W[l0:l1], precision="bf16", out="rank",
                             table=t, table_offset=l0, symmetric=True) for z, W, t in zip(zs, Ws, tables)]
    g = normalize.ensemble_logsum(members, Q)
    ens_quant.append(normalize.lower_triangle_quantiles(g, Q))
ens = normalize.EnsembleRankTable(Q, torch.cat(ens_quant))
def run(chunk, packed):
    n = 0
    for l0, l1, r in scoring.ensemble_fused_ranks_chunks(zs, pds, tables, ens, chunk=chunk, packed=packed):
        n += l1 - l0
    return n
res = {}
for packed in (False, True):
  for chunk in (16, 43):
    run(chunk, packed); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(chunk, packed); run(chunk, packed); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    res[f"packed={packed},chunk={chunk}"] = ms
    print(f"quantile-formulation ensemble K={K} L={L} N={N} chunk={chunk} packed={packed}: {ms:.2f} ms per pass "
          f"({K * L * N * N / ms / 1e9:.2f} T member-triples/s)")
# packed result == normaliser-layout result
from madrigal_b200.decoder import unpack_packed_tiles
a = [r.clone() for _, _, r in scoring.ensemble_fused_ranks_chunks(zs, pds, tables, ens, chunk=43)]
b = [unpack_packed_tiles(r, N) for _, _, r in scoring.ensemble_fused_ranks_chunks(zs, pds, tables, ens, chunk=43, packed=True)]
print("packed == mirrored:", all(torch.equal(x.view(torch.int16), y.view(torch.int16)) for x, y in zip(a, b)))
del a, b
# the exact chain on 4 outcomes for comparison
t0 = time.perf_counter()
for l0, l1, r in scoring.ensemble_normalized_ranks_chunks(zs, [W[:4] for W in Ws], precision="bf16", chunk=2):
    pass
torch.cuda.synchronize()
ex = (time.perf_counter() - t0) * 1e3 / 4
print(f"exact chain (logits -> exact rank x{K} -> gmean -> exact re-rank): {ex:.2f} ms per outcome => {ex * L:.0f} ms for {L} outcomes")
json.dump({"config": f"K={K} checkpoints x {L} outcomes x {N} drugs, D={D}", "quantile_formulation_ms": res,
           "exact_chain_ms_per_outcome": ex, "exact_chain_ms_extrapolated": ex * L},
          open(os.path.join(ROOT, "gpurun_out", "ensemble_fused_timing.json"), "w"), indent=1)
