"""Phase timeline of one fused-encoder tile (CTA 0): where the ~200 us of a 128-row tile go."""
import os, sys, ctypes
os.environ["MDG_FUSION_TRACE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb, synth
from madrigal_b200 import _lib
dev = torch.device("cuda:0")
B, T, E, H, hd, F, agg = 4096, 4, 256, 8, 32, 512, "x-attn"
cfg = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True, agg=agg, nb=0)
enc = mb.TransformerFusion(E, 0, 2, H, hd, F, transformer_actn="gelu", transformer_norm_first=True,
                           transformer_batch_first=False, transformer_agg=agg, precision="bf16")
enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(cfg, 1).items()})
enc.x_attn_key_padding_mask = torch.zeros(1, T, dtype=torch.bool)
enc = enc.to(dev).eval()
tokens = torch.randn(B, T, E, device=dev); mask = torch.rand(B, T, device=dev) < 0.5; mask[:, 0] = False
with torch.no_grad():
    for _ in range(3): enc(tokens, mask)
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 1024)()
n = _lib.lib().mdg_fusion_trace_read(buf, 1024)
a = np.array(buf[:512], dtype=np.int64); m = np.array(buf[512:1024], dtype=np.int64)
# one tile: epilogue records = signal(tokens), then pairs (wait_mma, signal)...; MMA records = pairs (wait_epi, commit)
names = ["tokens"]
for l in range(2):
    names += [f"L{l} LN1"] + [f"L{l} attn ph{q}" for q in range(H * hd // 64 if hd == 32 else H // 4)] + [f"L{l} LN2"] + [f"L{l} FFN c{c}" for c in range((F + 255) // 256)]
if agg == "x-attn":
    names += ["LN_kv"] + [f"pool ph{q}" for q in range(4)] + ["xout+qres"]
else:
    names += ["copy H"]
names += ["z out"]
raw = np.array(buf[:512], dtype=np.uint64); raw = raw[raw > 0]
tags = (raw >> np.uint64(56)).astype(int); clk = (raw & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)
t0 = clk[0]
TAG = {1: "wait_mma done", 2: "signalled", 3: "work done (before fences)", 4: "LN stats done", 5: "LN barrier passed", 6: "q loaded",
       7: "k/v stashed", 8: "kv barrier passed"}
prev = t0
for tg, c in zip(tags[:70], clk[:70]):
    print(f"{c - t0:8d} (+{c - prev:6d})  {TAG.get(tg, tg)}")
    prev = c
print("total cycles", clk[-1] - t0)
