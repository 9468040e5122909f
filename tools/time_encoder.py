"""Device-side timing of the fusion encoder (mdg_fusion_encode) on BASELINE config-5-style shapes."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb  # noqa: E402
import synth  # noqa: E402

dev = torch.device("cuda:0")


def flops_per_drug(T, E, Dl, F, layers, agg):
    per_tok = layers * (8 * Dl * Dl + 4 * Dl * F) + 2 * E * Dl
    pool = 4 * Dl * Dl * T + 2 * Dl * Dl + 2 * Dl * E if agg == "x-attn" else 2 * Dl * E * T
    return T * per_tok + pool


def case(B, T, E, H, hd, F, agg, precision, iters=5):
    cfg = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True, agg=agg, nb=0)
    enc = mb.TransformerFusion(E, 0, 2, H, hd, F, transformer_actn="gelu", transformer_norm_first=True,
                               transformer_batch_first=False, transformer_agg=agg, precision=precision)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in synth.fusion_state_dict(cfg, 1).items()})
    if agg == "x-attn":
        enc.x_attn_key_padding_mask = torch.zeros(1, T, dtype=torch.bool)
    enc = enc.to(dev).eval()
    rng = np.random.default_rng(0)
    tokens = torch.randn(B, T, E, device=dev)
    mask = torch.rand(B, T, device=dev) < 0.5
    mask[:, 0] = False
    with torch.no_grad():
        for _ in range(2):
            enc(tokens, mask)
        torch.cuda.synchronize()
        ts, ws = [], []
        for _ in range(iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            s.record()
            enc(tokens, mask)
            e.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
            ws.append((t1 - t0) * 1e3)
    ms = float(np.median(ts))
    fl = flops_per_drug(T, E, H * hd, F, 2, agg) * B
    print(f"B={B} T={T} E={E} Dl={H*hd} F={F} {agg} {precision}: {ms:.3f} ms device, host enqueue {np.median(ws):.3f} ms, "
          f"{enc.last_launch_count} launches -> {B / ms / 1e3:.2f} M drugs/s, {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if "--small" in sys.argv:
        case(4096, 4, 256, 8, 32, 512, "x-attn", "bf16", iters=3)
        sys.exit(0)
    case(4096, 4, 256, 8, 32, 512, "x-attn", "bf16")
    case(4096, 4, 256, 8, 32, 512, "x-attn", "fp32")
    case(262144, 4, 128, 8, 32, 512, "mean", "bf16")
    case(262144, 4, 128, 8, 64, 1024, "mean", "bf16")
    case(262144, 4, 128, 8, 32, 512, "mean", "fp32")
    case(1 << 20, 4, 128, 8, 32, 512, "mean", "bf16", iters=3)
    case(16384, 23, 128, 8, 64, 256, "x-attn", "bf16")
    case(16384, 23, 128, 8, 64, 256, "x-attn", "fp32")
