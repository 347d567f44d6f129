#!/usr/bin/env python
"""Hardware probe: tcgen05.mma dispatch rate of kind::f8f6f4 (e4m3, K=32) vs kind::f16 (K=16), one CTA (M=128) and a
CTA pair (cta_group::2, M=256), N = 64/128/256; mode 2 = 4 f16 + 4 f8 alternating on two accumulators."""
import ctypes as C
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text-detection-ctpn_b200"))
os.environ["CTPN_B200_LIB"] = "dbg"      # the probes live in the test library
from ctpn_b200 import _native as N  # noqa: E402
fn = N.lib.ctpn_probe_mma_kind
torch.cuda.set_device(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
n = 48000
for pair in (0, 1):
    grid = sms - (sms % 2 if pair else 0)
    M = 256 if pair else 128
    units = grid // 2 if pair else grid
    for bn in (64, 128, 256):
        for mode in (0, 1, 2):
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                N.check(fn(bn, pair, mode, n, grid, N.stream_ptr()), "probe kind")
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            kavg = {0: 16, 1: 32, 2: 24}[mode]
            print("%s bn=%3d mode=%s : %.3f ms  %.1f ns/MMA  %.1f TFLOP/s" % (
                "pair " if pair else "tight", bn, ["f16", "f8 ", "mix"][mode], best, best * 1e6 / n,
                2.0 * M * bn * kavg * n * units / (best * 1e-3) / 1e12), flush=True)
