"""Per-launch times of the rank kernel over a long back-to-back sequence, with NVML clock / power / event-reason samples.

Question (round 2): the N = 20,000 launches slow down from 22.7 ms (first) to 25.0 ms (steady): clocks, power, or memory?"""
import ctypes, os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import madrigal_b200 as mb
from madrigal_b200 import _lib, normalize
from synth import decoder_inputs
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
N, L, iters = int(os.environ.get("N", "20000")), int(os.environ.get("L", "30")), int(os.environ.get("ITERS", "40"))
z, W = decoder_inputs(N, 256, L, 0)
zt, Wt = torch.from_numpy(z).to(dev), torch.from_numpy(W).to(dev)
dec = mb.PreparedDecoder(Wt, precision="bf16")
table = normalize.build_rank_table(zt, Wt, 16384, kind=os.environ.get("KIND", "lut"), panel=2048, precision="bf16")
out = torch.empty((L, N, N), dtype=torch.uint16, device=dev)
fn = lambda: mb.pair_score(zt, zt, dec, precision="bf16", out="rank", table=table, out_tensor=out, symmetric=True)
fn(); torch.cuda.synchronize(); time.sleep(0.5)
samples, run = [], [True]
def loop():
    while run[0]:
        try:
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h), pynvml.nvmlDeviceGetTemperature(h, 0)))
        except Exception as e:
            samples.append((time.perf_counter(), -1, -1, -1, str(e), -1)); break
        time.sleep(0.001)
t = threading.Thread(target=loop, daemon=True); t.start()
_lib.lib().mdg_profile_enable(iters)
t0 = time.perf_counter()
for _ in range(iters): fn()
torch.cuda.synchronize()
t1 = time.perf_counter(); run[0] = False; t.join()
buf = (ctypes.c_float * 256)(); n = _lib.lib().mdg_profile_read(buf, 256)
ms = [buf[i] for i in range(n)]
print(f"N {N} L {L}: {n} launches in {(t1 - t0) * 1e3:.0f} ms; per launch:", " ".join(f"{x:.2f}" for x in ms))
k = max(1, len(samples) // 12)
for i in range(0, len(samples), k):
    s = samples[i]
    print(f"  t={1e3 * (s[0] - t0):7.1f} ms  sm {s[1]} MHz  mem {s[2]} MHz  {s[3]:.0f} W  reasons {s[4] if isinstance(s[4], str) else hex(s[4])}  {s[5]} C")
