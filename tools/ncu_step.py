#!/usr/bin/env python
"""One or more forward+proposal steps of the engine on synthetic data, for profiling under ncu (not a benchmark).

    ncu --set full --clock-control none -k regex:conv_tc -s 40 -c 16 -o gpurun_out/prof python tools/ncu_step.py --mode f16f8 --steps 2
In f16f8 mode the first step calibrates (24 extra conv_tc launches before its own 16); bf16 modes launch 16 conv_tc kernels per step."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "text-detection-ctpn_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ctpn_b200 import Engine, synthetic as synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="f16f8")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--height", type=int, default=600)
ap.add_argument("--width", type=int, default=900)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
eng = Engine(synth.make_weights(0), mode=a.mode)
rs = np.random.RandomState(100)
im = torch.from_numpy(rs.randint(0, 256, size=(a.batch, a.height, a.width, 3), dtype=np.uint8)).cuda()
info = torch.tensor([[a.height, a.width, 1.0]] * a.batch, dtype=torch.float32).cuda()
for _ in range(a.steps):
    eng.detect_packed(im, info)
torch.cuda.synchronize()
print("done")
