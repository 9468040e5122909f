"""Randomised GPU-vs-oracle fuzz of the decoder's rank outputs (full / normaliser / packed layouts, exact-LUT and
histogram-CDF tables, prepared and plain weights, bf16 and fp32-parity precision) at ragged sizes.  Every case is checked
bit for bit against np.searchsorted on the dense logits of the same call.  FUZZ_CASES / FUZZ_SEED env overrides."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize
from madrigal_b200.decoder import unpack_packed_tiles
from oracle import oracle
import synth
dev = torch.device("cuda:0")
rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "0")))
cases = int(os.environ.get("FUZZ_CASES", "120"))
t0 = time.time(); bad = 0
for it in range(cases):
    N = int(rng.choice([1, 2, 7, 31, 32, 33, 63, 64, 65, 127, 128, 129, 200, 255, 256, 257, 300, 511, 513, 640, 777, 1000]))
    if rng.random() < 0.3:
        N = int(rng.integers(2, 900))
    D = int(rng.choice([64, 128, 192, 256])); L = int(rng.integers(1, 5))
    prec = str(rng.choice(["bf16", "fp32"])); kind = str(rng.choice(["lut", "pwl"]))
    M = max(N * (N - 1) // 2, 1)
    Q = int(min(rng.choice([16, 255, 1024, 4096, 16384]), max(M, 1)))
    z, W = synth.decoder_inputs(N, D, L, seed=1000 + it)
    if rng.random() < 0.2:
        z *= np.float32(rng.choice([1e-3, 30.0]))     # scores far from / all over the table range
    zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
    if N >= 2:
        table = normalize.build_rank_table(zt, Wt, Q, precision=prec, kind=kind)
    else:
        table = mb.RankTable(torch.sort(torch.randn(L, 4, device=dev), dim=1).values.contiguous(), kind=kind)
    weight = mb.PreparedDecoder(Wt, prec) if rng.random() < 0.5 else Wt
    lg = mb.pair_score(zt, zt, Wt, precision=prec, out="logit").cpu().numpy()
    exp = oracle.quantile_rank(table.thresholds.cpu().numpy(), lg, "right").astype(np.uint16)
    full = mb.pair_score(zt, zt, weight, precision=prec, out="rank", table=table).cpu().numpy()
    low = np.tril(exp, -1); ref = low + low.swapaxes(1, 2)
    sym = mb.pair_score(zt, zt, weight, precision=prec, out="rank", table=table, symmetric=True)
    pk = mb.pair_score(zt, zt, weight, precision=prec, out="rank", table=table, packed=True)
    ok = np.array_equal(full, exp) and np.array_equal(sym.cpu().numpy(), ref) and \
        torch.equal(unpack_packed_tiles(pk, N).view(torch.int16), sym.view(torch.int16))
    if not ok:
        bad += 1
        print("MISMATCH", dict(it=it, N=N, D=D, L=L, Q=Q, prec=prec, kind=kind, prepared=not torch.is_tensor(weight)),
              "full", np.array_equal(full, exp), "sym", np.array_equal(sym.cpu().numpy(), ref))
print(f"fuzz: {cases} cases, {bad} mismatches, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
