#!/usr/bin/env python
"""Hardware probe: write-only and read-only HBM bandwidth next to the copy figure of MEASURED_PEAKS.json (conv1_1 is a pure
write stream: 3 B in, 256 B out per pixel)."""
import torch

dev = torch.device("cuda", 0)
n = 1 << 30                      # 4 GiB of float32
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)


def best(fn, reps=8):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


fill = best(lambda: x.fill_(1.0))
copy = best(lambda: y.copy_(x))
red = best(lambda: x.sum())
gb = n * 4 / 1e9
print("write-only (fill_)  : %.3f ms  %.0f GB/s" % (fill, gb / fill * 1e3))
print("copy (read + write) : %.3f ms  %.0f GB/s total" % (copy, 2 * gb / copy * 1e3))
print("read-only (sum)     : %.3f ms  %.0f GB/s" % (red, gb / red * 1e3))
