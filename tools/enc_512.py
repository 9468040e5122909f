"""One forward of the hidden-512 config-5 encoder shape (T=4, latent 512, FFN 1024, mean) for an ncu launch list."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb, synth
dev = torch.device("cuda:0")
B, T, E, H, hd, F = 65536, 4, 128, 8, 64, 1024
cfg = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True, agg="mean", nb=0)
enc = mb.TransformerFusion(E, 0, 2, H, hd, F, transformer_actn="gelu", transformer_norm_first=True,
                           transformer_batch_first=False, transformer_agg="mean", precision="bf16")
enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(cfg, 1).items()})
enc = enc.to(dev).eval()
tokens = torch.randn(B, T, E, device=dev); mask = torch.rand(B, T, device=dev) < 0.5; mask[:, 0] = False
with torch.no_grad():
    for _ in range(3): enc(tokens, mask)
torch.cuda.synchronize(); print("ok", enc.last_launch_count)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    for _ in range(10): enc(tokens, mask)
    e1.record()
torch.cuda.synchronize(); print("ms_per_forward", e0.elapsed_time(e1) / 10)
