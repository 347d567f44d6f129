#!/usr/bin/env python
"""GPU experiment that located the cross-stream allocator race of Engine.rois_batches (a buffer allocated on the copy stream
aliasing memory the compute stream was still reading; fixed in engine.py, regression test
tests/test_net_gpu.py::test_streaming_api_with_large_distinct_unpinned_batches): streamed batches must equal the same batches run one by one."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "text-detection-ctpn_b200"))
import numpy as np, torch
from oracle import synth
from ctpn_b200 import Engine
w = synth.make_weights(0)
eng = Engine(w, planes=2)
ims = np.stack([synth.make_image(20 + i, 128, 192) for i in range(3)])
rev = ims[::-1].copy()
b_fwd = eng.rois_batch(ims); b_rev = eng.rois_batch(rev)
print("rev batch == fwd singles reversed:", [np.array_equal(b_rev[2-i], b_fwd[i]) for i in range(3)])
st = list(eng.rois_batches([ims, rev]))
print("stream[0]==fwd", [np.array_equal(st[0][i], b_fwd[i]) for i in range(3)])
print("stream[1]==rev", [np.array_equal(st[1][i], b_rev[i]) for i in range(3)])
print("stream[1]==fwd (un-reversed?)", [np.array_equal(st[1][i], b_fwd[i]) for i in range(3)])
st2 = list(eng.rois_batches([rev, ims, rev]))
print("3 batches:", [np.array_equal(st2[0][i], b_rev[i]) for i in range(3)], [np.array_equal(st2[1][i], b_fwd[i]) for i in range(3)], [np.array_equal(st2[2][i], b_rev[i]) for i in range(3)])
