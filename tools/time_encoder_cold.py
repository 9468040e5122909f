"""Does the bench-shape encoder (4096 drugs x 4 tokens, one 128-row tile per CTA) pay for cold weights?  Each iteration
first streams 3 GB through L2 (as the rank kernel does in a bench step), then optionally touches the prepared weights
(a 2.7 MB read that pulls them back into L2), then times the encoder alone."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb, synth, bench
dev = torch.device("cuda:0")
ENC = bench.ENC
enc = mb.TransformerFusion(256, 0, 2, 8, 32, 512, transformer_actn="gelu", transformer_norm_first=True,
                           transformer_batch_first=False, transformer_agg="x-attn", precision="bf16")
enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(ENC, seed=7).items()})
enc.x_attn_key_padding_mask = torch.zeros(1, 4, dtype=torch.bool)
enc = enc.to(dev).eval()
tok, msk = synth.fusion_inputs(4096, 4, 256, seed=0)
tok, msk = torch.from_numpy(tok).to(dev), torch.from_numpy(msk).to(dev)
big = torch.empty(3 << 30, dtype=torch.uint8, device=dev)
with torch.no_grad():
    enc(tok, msk)
    prepared = enc._mdg_cache[2]
    for mode in ("cold (L2 flushed)", "weights touched first", "warm (no flush)"):
        ts = []
        for _ in range(12):
            if mode != "warm (no flush)":
                big.zero_()
            if mode == "weights touched first":
                prepared.view(torch.int32).sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); enc(tok, msk); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"{mode:24s}: encoder {np.median(ts[2:]):.1f} us (min {min(ts[2:]):.1f})")
