#!/usr/bin/env python
"""Build-container-only cross-check (needs /root/reference and oracle/_ref/bbox*.so): random training-target cases through
the REFERENCE's own anchor_target_layer, the oracle and the product's operator, all four output tensors compared bit for bit.
    python tests/golden/fuzz_train_vs_reference.py [cases]
The committed goldens (make_golden_train.py) are 7 of these cases; this run widens the pin without storing fixtures."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402
import make_golden_train  # noqa: E402
from product_import import load_product_module  # noqa: E402


def main(cases):
    cfg = make_golden.load_reference()[0]
    make_golden_train.load_bbox_module()
    from lib.rpn_msr import anchor_target_layer_tf as ref_layer
    from oracle import synth, train_targets as T
    ours = load_product_module("lib.rpn_msr.anchor_target_layer_tf").anchor_target_layer
    bad = 0
    stats = dict(fg=0, bg=0, hard=0, dontcare=0, outside=0, f64=0, weighted=0)
    for case in range(cases):
        rs = np.random.RandomState(case)
        H, W = int(rs.randint(3, 45)), int(rs.randint(6, 64))
        ih, iw = 16 * H - int(rs.randint(0, 16)), 16 * W - int(rs.randint(0, 16))
        scale = float(rs.choice([1.0, 0.61, 1.25, 1.6]))
        ndc = int(rs.randint(0, 10)) if case % 3 == 0 else 0
        hard_frac = float(rs.choice([0.0, 0.25]))
        outside = 1 if case % 11 == 0 else 0
        gt, hard, dc = synth.make_gt_boxes(500 + case, ih, iw, int(rs.randint(1, 7)), scale, ndc, hard_frac, outside)
        if case % 2:
            gt = gt.astype(np.float64) + rs.uniform(-0.45, 0.45, gt.shape) * (np.arange(5) < 4)
            stats["f64"] += 1
        if case % 7 == 0:      # a few degenerate annotations (x2 < x1, zero height)
            gt[0, 2] = gt[0, 0] - 3
            gt[-1, 3] = gt[-1, 1]
        pw = 0.4 if case % 5 == 0 else -1.0
        clobber = case % 4 == 0
        cfg.TRAIN.RPN_POSITIVE_WEIGHT, cfg.TRAIN.RPN_CLOBBER_POSITIVES = pw, clobber
        ocfg = dict(T.TRAIN_CFG, RPN_POSITIVE_WEIGHT=pw, RPN_CLOBBER_POSITIVES=clobber)
        score = np.zeros((1, H, W, 20), np.float32)
        info = np.array([[ih, iw, scale]], np.float32)
        outs = []
        with np.errstate(all="ignore"):
            for name, fn, extra in (("reference", ref_layer.anchor_target_layer, ([16, ], [16, ])),
                                    ("oracle", lambda *a: T.anchor_target_layer(*a, 16, ocfg), ()),
                                    ("product", lambda *a: product_call(ours, pw, clobber, *a), ([16, ], [16, ]))):
                np.random.seed(case)
                try:
                    outs.append(fn(score, gt.copy(), hard.copy(), dc.copy(), info, *extra))
                except Exception as e:      # noqa: BLE001  (a case the reference rejects must be rejected by all three)
                    outs.append(type(e).__name__)
        ref = outs[0]
        for name, got in zip(("oracle", "product"), outs[1:]):
            same = (isinstance(ref, str) and isinstance(got, str)) or (
                not isinstance(ref, str) and not isinstance(got, str) and all(same_tensor(a, b) for a, b in zip(ref, got)))
            if not same:
                bad += 1
                print("case %d: %s differs from the reference (%s vs %s)" % (case, name, ref if isinstance(ref, str) else "ok",
                                                                            got if isinstance(got, str) else "tensors"))
        if not isinstance(ref, str):
            stats["fg"] += int((ref[0] == 1).sum()); stats["bg"] += int((ref[0] == 0).sum())
            stats["hard"] += int(hard.sum()); stats["dontcare"] += ndc; stats["outside"] += outside; stats["weighted"] += pw > 0
    cfg.TRAIN.RPN_POSITIVE_WEIGHT, cfg.TRAIN.RPN_CLOBBER_POSITIVES = -1.0, False
    print("%d cases, %d mismatches; totals %s" % (cases, bad, stats))
    return bad


def same_tensor(a, b):
    """Bit-for-bit, except that NaNs (regression targets against a degenerate annotation: log of a negative width) only
    have to sit in the same places -- their sign / payload bits are not compared."""
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    nan = np.isnan(a)
    return np.array_equal(nan, np.isnan(b)) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))


def product_call(ours, pw, clobber, *args):
    pcfg = ours.__globals__["cfg"]
    saved = (pcfg.TRAIN.RPN_POSITIVE_WEIGHT, pcfg.TRAIN.RPN_CLOBBER_POSITIVES)
    pcfg.TRAIN.RPN_POSITIVE_WEIGHT, pcfg.TRAIN.RPN_CLOBBER_POSITIVES = pw, clobber
    try:
        return ours(*args)
    finally:
        pcfg.TRAIN.RPN_POSITIVE_WEIGHT, pcfg.TRAIN.RPN_CLOBBER_POSITIVES = saved


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 300) else 0)
