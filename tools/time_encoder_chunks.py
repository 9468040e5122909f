import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import time_encoder as t
t.case(1 << 20, 4, 128, 8, 32, 512, "mean", "bf16", iters=3)
t.case(1 << 20, 4, 128, 8, 64, 1024, "mean", "bf16", iters=3)
t.case(65536, 23, 128, 8, 64, 256, "x-attn", "bf16", iters=3)
