#!/usr/bin/env python
"""Hardware probe (GPU box): sustained tcgen05.mma rate of one issuing thread per SM as a function of the
tcgen05.commit cadence and of waiting on the commit barriers.  Prints TFLOP/s for each configuration."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text-detection-ctpn_b200"))
os.environ["CTPN_B200_LIB"] = "dbg"      # the probes live in the test library
from ctpn_b200 import _native as N  # noqa: E402

torch.cuda.set_device(0)
sms = torch.cuda.get_device_properties(0).multi_processor_count


def run(bn, n_mma, commit_every, lag, alt, fence, grid=sms, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(N.lib.ctpn_probe_mma_rate(bn, n_mma, commit_every, lag, alt, fence, grid, N.stream_ptr()), "probe")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tf = 2.0 * 128 * bn * 16 * n_mma * grid / (best * 1e-3) / 1e12
    print("bn=%3d n=%6d commit_every=%5d lag=%2d alt_acc=%d fence=%d grid=%3d : %8.3f ms  %8.1f TFLOP/s" % (bn, n_mma, commit_every, lag, alt, fence, grid, best, tf), flush=True)


n = 48000
for bn in (256, 128, 64):
    run(bn, n, n, 0, 0, 0)                 # one commit at the end: pure issue rate
    run(bn, n, 4, 0, 0, 0)                 # commit every 4 MMAs, never wait
    run(bn, n, 12, 0, 0, 0)
    run(bn, n, 4, 4, 0, 0)                 # wait on the commit 4 groups back (ring of 4 stages)
    run(bn, n, 4, 2, 0, 0)
    run(bn, n, 4, 1, 0, 0)                 # wait for the previous group before issuing the next: full serialisation
    run(bn, n, 12, 2, 0, 0)
    run(bn, n, 12, 2, 1, 1)
run(256, n, 4, 2, 0, 0, grid=1)
run(256, n, n, 0, 0, 0, grid=1)
run(256, n, n, 0, 0, 0, grid=74)
