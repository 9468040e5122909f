"""One small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck) runs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import normalize, scoring
import synth
dev = torch.device("cuda:0")
N, D, L, Q = 203, 128, 2, 1024
z, W = synth.decoder_inputs(N, D, L, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
lg = mb.pair_score(zt, zt, Wt, precision="fp32", out="logit")
sg = mb.pair_score(zt, zt, Wt, precision="bf16", out="sigmoid")
quant = normalize.build_reference_quantiles(zt, Wt, Q, precision="bf16")
for kind in ("lut", "pwl"):
    table = mb.RankTable(quant, kind=kind)
    for sym in (False, True):
        r = mb.pair_score(zt, zt, Wt, precision="bf16", out="rank", table=table, symmetric=sym)
    table.lookup(lg)
thr = table.thresholds[:, Q - 20].contiguous()
mb.pair_topk(zt, zt, Wt, thr, 50, cap=4096, symmetric=True)
idx = torch.arange(300, device=dev)
mb.pair_score_gather(zt, zt, Wt, idx % L, idx % N, (idx * 3) % N, precision="fp32")
ex = normalize.exact_normalized_ranks(lg)
normalize.gmean_normalized_ranks([ex, ex, ex])
for prec, agg, T, dims in (("bf16", "x-attn", 4, (128, 8, 32, 256)), ("bf16", "mean", 5, (64, 4, 16, 100)),
                           ("fp32", "cls", 7, (64, 2, 64, 128)), ("bf16", "max", 23, (128, 8, 64, 256))):
    E, H, hd, F = dims
    cfg = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True, agg=agg, nb=0)
    enc = mb.TransformerFusion(E, 0, 2, H, hd, F, transformer_actn="gelu", transformer_norm_first=True,
                               transformer_batch_first=False, transformer_agg=agg, precision=prec)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(cfg, 1).items()})
    if agg == "x-attn":
        enc.x_attn_key_padding_mask = torch.zeros(1, T, dtype=torch.bool)
    enc = enc.to(dev).eval()
    tok, msk = synth.fusion_inputs(70, T, E, 3)
    with torch.no_grad():
        enc(torch.from_numpy(tok).to(dev), torch.from_numpy(msk).to(dev))
torch.cuda.synchronize()
print("sanitize_small ok")
