#!/usr/bin/env python
"""Hardware probe: back-to-back tcgen05.mma dispatch rate (8 MMAs per elected region) for N = 64 / 128 / 256."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text-detection-ctpn_b200"))
os.environ["CTPN_B200_LIB"] = "dbg"      # the probes live in the test library
from ctpn_b200 import _native as N  # noqa: E402
torch.cuda.set_device(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
n = 48000
for bn in (64, 128, 256):
    for alt in (0, 1):
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(N.lib.ctpn_probe_mma_rate(bn, n, n, 0, alt, 2, sms, N.stream_ptr()), "probe")
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("tight bn=%3d alt_acc=%d : %.3f ms  %.1f ns/MMA  %.1f TFLOP/s" % (bn, alt, best, best * 1e6 / n, 2.0 * 128 * bn * 16 * n * sms / (best * 1e-3) / 1e12), flush=True)

grid2 = sms - sms % 2
for bn in (64, 128, 256):
    for alt in (0, 1):
        if alt and 2 * bn > 512:
            continue
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            N.check(N.lib.ctpn_probe_mma_rate_pair(bn, n, alt, grid2, N.stream_ptr()), "probe pair")
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("pair  bn=%3d alt_acc=%d : %.3f ms  %.1f ns/MMA(M=256)  %.1f TFLOP/s" % (bn, alt, best, best * 1e6 / n, 2.0 * 256 * bn * 16 * n * (grid2 // 2) / (best * 1e-3) / 1e12), flush=True)
