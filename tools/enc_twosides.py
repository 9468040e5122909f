"""One forward of the production TWOSIDES encoder shape (T=21, 2 heads of 256, FFN 512, x-attn, nb=2): timing + ncu list."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb, synth
dev = torch.device("cuda:0")
B, T, E, H, hd, F = 16384, 21, 128, 2, 256, 512
cfg = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True, agg="x-attn", nb=2)
enc = mb.TransformerFusion(E, 2, 2, H, hd, F, transformer_actn="gelu", transformer_norm_first=True,
                           transformer_batch_first=False, transformer_agg="x-attn", precision=os.environ.get("PREC", "bf16"))
enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(cfg, 1).items()})
enc = enc.to(dev).eval()
tokens = torch.randn(B, T, E, device=dev); mask = torch.rand(B, T, device=dev) < 0.5; mask[:, 0] = False
with torch.no_grad():
    for _ in range(3): enc(tokens, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): enc(tokens, mask)
    e1.record(); torch.cuda.synchronize()
print("ok launches", enc.last_launch_count, "ms_per_forward", e0.elapsed_time(e1) / 5)
