#!/usr/bin/env python
"""Perf experiment: time conv1_1 (tensor-core kernel) alone on a 32 x 600 x 900 uint8 batch.  CTPN_C1_DEBUG selects which
part of the kernel is skipped (see conv1_tc.cu); results are wrong with any bit set -- timing only."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text-detection-ctpn_b200"))
from ctpn_b200 import _native as N  # noqa: E402
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, H, W = 32, 600, 900
torch.cuda.set_device(0)
im = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
lut = torch.randn(768, device="cuda"); w = torch.randn(27 * 64, device="cuda") * 0.01; b = torch.zeros(64, device="cuda")
out = torch.empty((planes, B, H, W, 64), dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for it in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    N.check(N.lib.ctpn_conv1_1_tc(N.ptr(im), 0, N.ptr(lut), N.ptr(w), N.ptr(b), N.ptr(out), B, H, W, planes, N.stream_ptr()), "conv1")
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("planes=%d CTPN_C1_DEBUG=%s : %.3f ms (min of %d, after 3 warm-up)" % (planes, os.environ.get("CTPN_C1_DEBUG", "0"), min(ts[3:]), len(ts) - 3), flush=True)
