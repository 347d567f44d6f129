#!/usr/bin/env python
"""Perf experiment (test library; CTPN_LSTM_NC=2 forces the 2-CTA-cluster variant): time the BiLSTM recurrence alone."""
import os
import sys

os.environ.setdefault("CTPN_B200_LIB", "dbg")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text-detection-ctpn_b200"))
import torch  # noqa: E402
from ctpn_b200 import _native as N  # noqa: E402

R, W, planes = int(sys.argv[1]) if len(sys.argv) > 1 else 1184, int(sys.argv[2]) if len(sys.argv) > 2 else 56, 2
dev = torch.device("cuda", 0)
x = torch.randn(R, W, 1024, device=dev)
wh = [torch.randn(128, 512, device=dev) * 0.05 for _ in range(2)]
out = torch.zeros((planes, R, W, 256), dtype=torch.bfloat16, device=dev)
run = lambda: N.check(N.lib.ctpn_bilstm_recurrent(N.ptr(x), N.ptr(wh[0]), N.ptr(wh[1]), N.ptr(out), R, W, planes, N.stream_ptr()), "bilstm")
for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print("bilstm R=%d W=%d NC=%s: median %.3f ms (min %.3f)" % (R, W, os.environ.get("CTPN_LSTM_NC", "auto"), ts[10], ts[0]), flush=True)
