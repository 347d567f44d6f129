#!/usr/bin/env python
"""Host timing of the RPN training-target operator on one CPU core: this repo's
lib.rpn_msr.anchor_target_layer_tf.anchor_target_layer (one native pass + numpy sub-sampling) next to the CPU oracle
(oracle/train_targets.py) and, where /root/reference and oracle/_ref/bbox*.so exist (the build container), the reference's
own operator (numpy + its Cython bbox module).  All three produce bit-identical outputs (tests/test_train_targets_cpu.py).
    python tools/time_train_targets.py > profiles/r2_train_targets_cpu.txt
"""
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np  # noqa: E402

CASES = [("600x900, 6 lines", 37, 56, 600, 900, 6, 0), ("600x900, 12 lines + 8 dontcare", 37, 56, 600, 900, 12, 8),
         ("1200x1600, 24 lines + 8 dontcare", 75, 100, 1200, 1600, 24, 8)]


def timed(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(t))


def main():
    ref_layer = None
    if os.path.isdir("/root/reference/lib"):
        import make_golden
        import make_golden_train
        make_golden.load_reference()
        make_golden_train.load_bbox_module()
        from lib.rpn_msr import anchor_target_layer_tf as ref_layer       # the reference's (its lib/ is first on sys.path)
        assert os.path.realpath(ref_layer.__file__).startswith("/root/reference")
    from oracle import synth, train_targets as T
    # the product's operator lives in a package of the same name: load it by path
    from product_import import load_product_module          # the product's operator lives in a package of the same name
    ours = load_product_module("lib.rpn_msr.anchor_target_layer_tf").anchor_target_layer
    cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][:1]
    print("host: %s, %s, numpy %s, 1 thread" % (cpu[0] if cpu else platform.processor(), platform.platform(), np.__version__))
    print("%-34s %6s %8s | %12s %12s %12s | %s" % ("case", "gt", "anchors", "reference ms", "oracle ms", "this repo ms", "speed-up vs reference"))
    for name, H, W, ih, iw, lines, ndc in CASES:
        gt, hard, dc = synth.make_gt_boxes(1, ih, iw, lines, 1.0, ndc, 0.1)
        score = np.zeros((1, H, W, 20), np.float32)
        info = np.array([[ih, iw, 1.0]], np.float32)
        outs = {}

        def run(fn, key, *extra):
            def go():
                np.random.seed(5)
                outs[key] = fn(score, gt, hard, dc, info, *extra)
            return go
        t_or = timed(run(T.anchor_target_layer, "oracle", 16), 5)
        t_us = timed(run(ours, "ours", [16, ], [16, ]), 50)
        t_ref = None
        if ref_layer is not None:
            t_ref = timed(run(ref_layer.anchor_target_layer, "ref", [16, ], [16, ]), 5)
            assert all(np.array_equal(a, b) for a, b in zip(outs["ref"], outs["ours"]))
        assert all(np.array_equal(a, b) for a, b in zip(outs["oracle"], outs["ours"]))
        print("%-34s %6d %8d | %12s %12.2f %12.3f | %s" % (name, gt.shape[0], H * W * 10, "%.2f" % t_ref if t_ref else "-", t_or, t_us,
                                                        "%.0fx" % (t_ref / t_us) if t_ref else "-"))


if __name__ == "__main__":
    main()
