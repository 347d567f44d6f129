#!/usr/bin/env python
"""Generate tests/golden/reference_anchor_targets.npz by running the REFERENCE's own training-target code, imported
unmodified from /root/reference: lib/rpn_msr/anchor_target_layer_tf.py on top of its Cython lib/utils/bbox.pyx, which
oracle/Makefile re-cythonizes and compiles into oracle/_ref/ (the reference's own build is not run).

Run in the build container only (the GPU box has no /root/reference):
    make -C oracle && python tests/golden/make_golden_train.py

Shims are those of make_golden.py (easydict stand-in, np.float / np.int aliases); the compiled module is registered as
lib.utils.bbox.  numpy's global RNG is seeded per case (the layer sub-samples with numpy.random.choice); the seed is stored.
Inputs come from oracle/synth.py::make_gt_boxes; only the case parameters and the outputs are stored.
"""
import glob
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

CASES = [  # tag, seed, H, W, im_h, im_w, scale, lines, dontcare, hard_frac, outside, positive_weight, clobber
    ("cfgA", 1, 37, 56, 600, 900, 1.0, 6, 0, 0.0, 0, -1.0, False),
    ("few", 2, 37, 56, 600, 900, 1.0, 1, 0, 0.0, 0, -1.0, False),
    ("scaled_hard_dontcare", 3, 37, 56, 600, 900, 1.171875, 5, 4, 0.2, 0, -1.0, False),
    ("ragged", 4, 9, 13, 150, 215, 1.0, 2, 1, 0.0, 0, -1.0, False),
    ("outside_gt", 5, 12, 18, 192, 288, 1.0, 2, 0, 0.0, 1, -1.0, False),
    ("weighted_clobber", 6, 37, 56, 600, 900, 0.8, 4, 2, 0.1, 0, 0.5, True),
    ("tall", 7, 62, 38, 1000, 608, 1.6, 8, 3, 0.1, 0, -1.0, False),
]
BBOX_CASES = [(1, 300, 40), (2, 1, 1), (3, 64, 200)]        # seed, N, K


def load_bbox_module():
    paths = glob.glob(os.path.join(ROOT, "oracle", "_ref", "bbox.*.so"))
    assert paths, "run `make -C oracle` first"
    spec = importlib.util.spec_from_file_location("lib.utils.bbox", paths[0])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lib.utils.bbox"] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    cfg = make_golden.load_reference()[0]
    bbox = load_bbox_module()
    from lib.rpn_msr import anchor_target_layer_tf as ref_layer
    assert os.path.realpath(ref_layer.__file__).startswith(make_golden.REF)
    assert ref_layer.bbox_overlaps is bbox.bbox_overlaps
    from oracle import synth
    out = {"meta_numpy": np.array(np.__version__), "meta_reference_commit": np.array("c04a571e2593fc361c1aff3127e58dc13fdc4e5a")}

    for seed, n, k in BBOX_CASES:
        a = synth.make_boxes(seed, n)[:, :4].astype(np.float64)
        b = synth.make_boxes(seed + 50, k, ctpn_like=True)[:, :4].astype(np.float64)
        out["bbox_%d_cfg" % seed] = np.array([seed, n, k], np.int64)
        out["bbox_%d_overlaps" % seed] = bbox.bbox_overlaps(a, b)
        out["bbox_%d_intersections" % seed] = bbox.bbox_intersections(a, b)

    for tag, seed, H, W, ih, iw, scale, lines, ndc, hard_frac, outside, pw, clobber in CASES:
        gt, hard, dc = synth.make_gt_boxes(seed, ih, iw, lines, scale, ndc, hard_frac, outside)
        cfg.TRAIN.RPN_POSITIVE_WEIGHT, cfg.TRAIN.RPN_CLOBBER_POSITIVES = pw, clobber
        np.random.seed(7000 + seed)
        score = np.zeros((1, H, W, 20), np.float32)
        info = np.array([[ih, iw, scale]], np.float32)
        res = ref_layer.anchor_target_layer(score, gt.copy(), hard.copy(), dc.copy(), info, [16, ], [16, ])
        out["atl_%s_cfg" % tag] = np.array([seed, H, W, ih, iw, lines, ndc, outside, int(clobber)], np.int64)
        out["atl_%s_params" % tag] = np.array([scale, hard_frac, pw], np.float64)
        for name, arr in zip(("labels", "targets", "inside", "outside"), res):
            assert arr.dtype == np.float32
            out["atl_%s_%s" % (tag, name)] = arr
        print(tag, "gt", gt.shape[0], "hard", int(hard.sum()), "fg", int((res[0] == 1).sum()), "bg", int((res[0] == 0).sum()))
    cfg.TRAIN.RPN_POSITIVE_WEIGHT, cfg.TRAIN.RPN_CLOBBER_POSITIVES = -1.0, False
    path = os.path.join(HERE, "reference_anchor_targets.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
