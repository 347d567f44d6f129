#!/usr/bin/env python
"""2-GPU experiment: where do the per-step gaps of the device-timed loop come from when the results are all-gathered?
    torchrun --nproc-per-node 2 tools/dbg_gather.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text-detection-ctpn_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ctpn_b200 import Engine, synthetic as synth  # noqa: E402
from ctpn_b200.dist import gather_packed  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(synth.make_weights(0), mode="f16f8", device=local)
im = torch.from_numpy(np.random.RandomState(rank).randint(0, 256, (32, 600, 900, 3), dtype=np.uint8)).to(dev)
info = torch.tensor([[600, 900, 1.0]] * 32, device=dev)
side = torch.cuda.Stream(device=dev)
K = 20


def run(variant):
    keep = []
    for _ in range(3):
        eng.detect_packed(im, info)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    host = []
    for _ in range(K):
        h0 = time.perf_counter()
        p = eng.detect_packed(im, info)
        h1 = time.perf_counter()
        if variant == "main":
            keep.append(gather_packed(p))
        elif variant == "side":
            ev = torch.cuda.Event(); ev.record()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                keep.append((p, gather_packed(p)))
        elif variant == "side_async":
            ev = torch.cuda.Event(); ev.record()
            with torch.cuda.stream(side):
                side.wait_event(ev)
                out = torch.empty((2 * p.numel(),), dtype=p.dtype, device=dev)
                w = dist.all_gather_into_tensor(out, p, async_op=True)
                keep.append((p, out, w))
        host.append((h1 - h0, time.perf_counter() - h1))
    torch.cuda.current_stream().wait_stream(side)
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    for k in keep:
        if isinstance(k, tuple) and len(k) == 3:
            k[2].wait()
    ms = e0.elapsed_time(e1) / K
    print("rank %d %-10s: %.2f ms/step (gpu), host enqueue: forward %.2f ms, gather call %.3f ms (max %.3f); wall %.2f ms/step" % (
        rank, variant, ms, 1e3 * np.mean([h[0] for h in host]), 1e3 * np.mean([h[1] for h in host]), 1e3 * max(h[1] for h in host), (t1 - t0) / K * 1e3), flush=True)


for v in ("none", "main", "side", "side_async", "none"):
    run(v)
dist.destroy_process_group()
