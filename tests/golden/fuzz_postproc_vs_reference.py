#!/usr/bin/env python
"""Build-container-only cross-check (needs /root/reference): random cases through the REFERENCE's own post-processing code,
the CPU oracle and -- where the product runs on the host -- the product:
  text lines   reference TextDetector (H and O)  vs  oracle/textline.py  vs  the product's TextDetector class (C++ filter /
               NMS / grouping + numpy fit: must be bit-identical) and its all-native path (same lines; coordinates differ
               from the numpy fit by float32 rounding only, reported in pixels)
  proposals    reference proposal_layer (numpy + py_cpu_nms)  vs  oracle/postproc.py   (the product's is a CUDA kernel,
               checked against the oracle by the GPU tests)
  nms          reference nms_wrapper.nms  vs  oracle/postproc.py
    python tests/golden/fuzz_postproc_vs_reference.py [text_cases] [proposal_cases]
The committed goldens (make_golden.py) hold 8 + 5 + 8 such cases; this run widens the pin without storing fixtures."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402
from product_import import load_product_module  # noqa: E402


def main(text_cases, prop_cases):
    cfg, nms_wrapper, _, proposal_layer, RefTextDetector = make_golden.load_reference()
    from oracle import postproc, synth, textline
    product = load_product_module("lib.text_connector.detectors")
    pcfg = product.cfg
    bad = 0
    lines_total = 0
    worst_native = {"H": 0.0, "O": 0.0}
    for case in range(text_cases):
        rs = np.random.RandomState(case)
        ih, iw = int(rs.randint(120, 1300)), int(rs.randint(200, 1700))
        tp, sc = synth.make_text_proposals(100 + case, ih, iw, int(rs.randint(0, 14)), int(rs.randint(0, 300)))
        for mode in ("H", "O"):
            cfg.TEST.DETECT_MODE = pcfg.TEST.DETECT_MODE = mode
            ref = np.asarray(RefTextDetector().detect(tp.copy(), sc.copy(), (ih, iw)), np.float64).reshape(-1, 9)
            orc = np.asarray(textline.detect(tp.copy(), sc.copy(), (ih, iw), mode), np.float64).reshape(-1, 9)
            got = np.asarray(product.TextDetector().detect(tp.copy(), sc.copy(), (ih, iw)), np.float64).reshape(-1, 9)
            nat = np.asarray(product.TextDetector(native=True).detect(tp.copy(), sc.copy(), (ih, iw)), np.float64).reshape(-1, 9)
            lines_total += ref.shape[0]
            for name, arr in (("oracle", orc), ("product", got)):
                if arr.shape != ref.shape or not np.array_equal(arr, ref):
                    bad += 1
                    print("text case %d mode %s: %s differs from the reference (%s vs %s rows)" % (case, mode, name, arr.shape[0], ref.shape[0]))
            if nat.shape != ref.shape:
                bad += 1
                print("text case %d mode %s: native path has %d lines, reference %d" % (case, mode, nat.shape[0], ref.shape[0]))
            elif ref.size:
                worst_native[mode] = max(worst_native[mode], float(np.abs(nat - ref).max()))
    cfg.TEST.DETECT_MODE = pcfg.TEST.DETECT_MODE = "H"
    print("text lines: %d cases x 2 modes, %d lines, %d mismatches (oracle and product class vs reference, bit for bit); "
          "all-native path: same lines, worst coordinate deviation H %.2e px, O %.2e px"
          % (text_cases, lines_total, bad, worst_native["H"], worst_native["O"]))

    bad_p = 0
    canon_diff = 0
    tied = 0
    rows = 0
    for case in range(prop_cases):
        rs = np.random.RandomState(1000 + case)
        H, W = int(rs.randint(4, 40)), int(rs.randint(4, 58))
        scale = float(rs.choice([1.0, 0.8, 1.5]))
        ih, iw = 16 * H - int(rs.randint(0, 16)), 16 * W - int(rs.randint(0, 16))
        pre, post = (12000, 1000) if case % 3 else (int(rs.randint(50, 3000)), int(rs.randint(10, 400)))
        cls_prob, bbox = synth.make_head_outputs(200 + case, H, W, logit_std=float(rs.choice([0.5, 1.0, 1.5])),
                                                 delta_std=float(rs.choice([0.1, 0.3, 0.8])))
        info = np.array([[ih, iw, scale]], np.float32)
        fg = cls_prob[..., 10:]
        if np.unique(fg).size != fg.size:
            # saturated softmax values collide in float32: the reference's `argsort()[::-1]` leaves the order of equal scores
            # unspecified (quicksort), so there is no reference answer to compare with (ties -> ascending index)
            tied += 1
            continue
        cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = pre, post
        blob, deltas = proposal_layer(cls_prob.copy(), bbox.copy(), info, "TEST")
        # exp_mode="numpy" is the reference's own float32 np.exp; the canonical oracle (and the CUDA kernel) round exp
        # correctly instead: same rows, coordinates within 1 ulp
        oblob, odeltas = postproc.proposal_layer(cls_prob, bbox, info, pre_nms_topN=pre, post_nms_topN=post, exp_mode="numpy")
        cblob, _ = postproc.proposal_layer(cls_prob, bbox, info, pre_nms_topN=pre, post_nms_topN=post, exp_mode="rounded")
        rows += blob.shape[0]
        if blob.shape != oblob.shape or not np.array_equal(blob, oblob) or not np.array_equal(deltas, odeltas):
            bad_p += 1
            print("proposal case %d (%dx%d, pre %d post %d): oracle differs from the reference" % (case, H, W, pre, post))
        if cblob.shape != blob.shape or not np.array_equal(cblob[:, 0], blob[:, 0]) or not np.allclose(cblob, blob, rtol=3e-7, atol=1e-4):
            canon_diff += 1
            print("proposal case %d: correctly-rounded-exp oracle selects different rows" % case)
        dets = synth.make_boxes(300 + case, int(rs.randint(1, 1500)), ctpn_like=bool(case % 2))
        thr = float(rs.choice([0.2, 0.3, 0.5, 0.7]))
        if list(nms_wrapper.nms(dets.copy(), thr)) != list(postproc.nms(dets, thr)):
            bad_p += 1
            print("nms case %d: oracle differs from the reference" % case)
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = 12000, 1000
    print("proposal layer + nms: %d cases, %d output rows, %d mismatches (numpy-exp oracle, bit for bit); correctly rounded exp: "
          "%d cases with a different row set; %d generated cases skipped for tied scores" % (prop_cases - tied, rows, bad_p, canon_diff, tied))
    return bad + bad_p + canon_diff


if __name__ == "__main__":
    a = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    sys.exit(1 if main(a, b) else 0)
