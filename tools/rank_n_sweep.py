"""Rank kernel (normaliser layout, exact LUT) against the catalogue size N at a fixed output volume.

Question (round 2, configs[3]): the kernel reaches 0.63-0.66 of the HBM copy bandwidth at N = 4096 and 0.58 at
N = 20,000.  Is that the row stride (40,000 B is not a multiple of 128), the slab size per outcome, or the SM clock
under a long launch?  Every case writes about the same number of bytes; SM clock and board power are sampled through
NVML while the launches run.
"""
import ctypes
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb  # noqa: E402
from madrigal_b200 import _lib, normalize  # noqa: E402
from synth import decoder_inputs  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6556.2


class Sampler:
    def __init__(self):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.run = False
        self.s = []

    def _loop(self):
        while self.run:
            try:
                self.s.append((self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                               self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        self.s = []
        self.run = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.run = False
        self.t.join()

    def summary(self):
        if not self.s:
            return "no samples"
        c = np.array([x[0] for x in self.s]); p = np.array([x[1] for x in self.s])
        return f"sm {np.median(c):.0f} MHz (min {c.min():.0f}) power {np.median(p):.0f} W (max {p.max():.0f}) n={len(c)}"


def kms(fn, iters):
    fn(); torch.cuda.synchronize()
    _lib.lib().mdg_profile_enable(iters)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    buf = (ctypes.c_float * 64)()
    n = _lib.lib().mdg_profile_read(buf, 64)
    _lib.lib().mdg_profile_enable(0)
    return float(np.mean(buf[:n])), float(np.min(buf[:n]))


def main():
    target = float(os.environ.get("SWEEP_GB", "12")) * 1e9
    sizes = [int(x) for x in os.environ.get("SWEEP_N", "4096,8192,12288,16384,19968,20000,20480").split(",")]
    kinds = os.environ.get("SWEEP_KINDS", "lut,pwl,packed").split(",")
    sampler = Sampler()
    for N in sizes:
        L = max(1, int(round(target / (2.0 * N * N))))
        z, W = decoder_inputs(N, 256, L, 0)
        zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
        dec = mb.PreparedDecoder(Wt, precision="bf16") if hasattr(mb, "PreparedDecoder") else Wt
        for kind in kinds:
            packed = kind == "packed"
            table = normalize.build_rank_table(zt, Wt, 16384, kind="pwl" if kind == "pwl" else "lut", panel=2048,
                                               precision="bf16")
            shape = (L, mb.decoder.packed_tiles_per_outcome(N), 32, 32) if packed else (L, N, N)
            out = torch.empty(shape, dtype=torch.uint16, device=dev)
            fn = lambda: mb.pair_score(zt, zt, dec, precision="bf16", out="rank", table=table, out_tensor=out,
                                       symmetric=True, packed=packed)
            iters = 6
            with sampler:
                ms, best = kms(fn, iters)
            gb = out.numel() * 2 / 1e9
            print(f"N={N:6d} L={L:4d} {kind:6s} {gb:6.2f} GB  kernel {ms:8.3f} ms (best {best:8.3f}) -> {gb / ms * 1e3:6.0f} GB/s "
                  f"= {gb / ms * 1e3 / HBM:.3f} of copy bw | {sampler.summary()}", flush=True)
            del out
        del zt, Wt, dec
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
