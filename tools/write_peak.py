"""Write-only HBM bandwidth of this GPU (context for the rank kernel's roofline: its traffic is ~all writes)."""
import torch, numpy as np
dev = torch.device("cuda:0")
for nbytes in (2886 << 20, 8 << 30):
    x = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for fn, name in ((lambda: x.fill_(1), "torch fill_"), (lambda: x.zero_(), "torch zero_ (memset)")):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        print(f"{name}: {nbytes / 2**30:.2f} GiB, best {min(ts):.3f} ms -> {nbytes / min(ts) / 1e6:.0f} GB/s, median {np.median(ts):.3f} ms -> {nbytes / np.median(ts) / 1e6:.0f} GB/s")
    a = torch.empty(nbytes // 2, dtype=torch.uint8, device=dev); b = torch.empty_like(a)
    for _ in range(3): b.copy_(a)
    ts = []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); b.copy_(a); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print(f"copy: {nbytes / 2**30:.2f} GiB moved (read+write), best {min(ts):.3f} ms -> {nbytes / min(ts) / 1e6:.0f} GB/s")
    del x, a, b
