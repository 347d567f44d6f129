#!/usr/bin/env python
"""Generate tests/golden/reference_postproc.npz by running the REFERENCE's own
numpy post-processing code, imported unmodified from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Shims (nothing in the reference is edited):
  * an `easydict` stand-in module (the package is not installed);
  * np.float / np.int aliases, only if missing (numpy >= 1.24 removed them);
    np.bool is NOT touched.
  * text.yml's three TEST overrides are applied on `cfg` directly because
    config.py:292 calls yaml.load() without a Loader (TypeError on PyYAML >= 6).
The reference falls back to py_cpu_nms (nms_wrapper.py:3-8) because its Cython
modules are not built -- that is the reference's own CPU NMS path.

Inputs come from oracle/synth.py (seeded); only seeds + outputs are stored.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


class _EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def load_reference():
    mod = types.ModuleType("easydict")
    mod.EasyDict = _EasyDict
    sys.modules["easydict"] = mod
    for name, typ in (("float", float), ("int", int)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    sys.path.insert(0, REF)
    sys.path.insert(1, ROOT)
    from lib.fast_rcnn.config import cfg
    from lib.fast_rcnn import nms_wrapper
    from lib.rpn_msr.generate_anchors import generate_anchors
    from lib.rpn_msr.proposal_layer_tf import proposal_layer
    from lib.text_connector.detectors import TextDetector
    assert os.path.realpath(nms_wrapper.__file__).startswith(REF)
    cfg.TEST.HAS_RPN = True
    cfg.TEST.DETECT_MODE = "H"
    return cfg, nms_wrapper, generate_anchors, proposal_layer, TextDetector


PROPOSAL_CASES = [  # (tag, seed, H, W, im_h, im_w, scale, pre, post)
    ("small", 1, 12, 18, 192, 288, 1.0, 12000, 1000),
    ("small_topn", 2, 12, 18, 192, 288, 1.0, 300, 50),
    ("scaled", 3, 10, 14, 160, 224, 1.5, 12000, 1000),
    ("cfgA", 4, 37, 56, 600, 900, 1.0, 12000, 1000),
    ("ragged", 5, 9, 13, 150, 215, 1.0, 12000, 1000),
]
NMS_CASES = [  # (tag, seed, n, ctpn_like, thresh)
    ("one", 1, 1, False, 0.7), ("generic500", 2, 500, False, 0.7), ("generic500_t02", 2, 500, False, 0.2),
    ("generic64", 3, 64, False, 0.5), ("generic65", 4, 65, False, 0.5),
    ("ctpn2000", 5, 2000, True, 0.7), ("ctpn2000_t02", 5, 2000, True, 0.2), ("dense3000", 6, 3000, False, 0.3),
]
TEXT_CASES = [0, 1, 2, 3]


def main():
    cfg, nms_wrapper, generate_anchors, proposal_layer, TextDetector = load_reference()
    from oracle import synth
    out = {"meta_numpy": np.array(np.__version__), "meta_reference_commit": np.array("c04a571e2593fc361c1aff3127e58dc13fdc4e5a")}
    out["anchors"] = generate_anchors()

    for tag, seed, H, W, ih, iw, scale, pre, post in PROPOSAL_CASES:
        cls_prob, bbox = synth.make_head_outputs(seed, H, W)
        cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = pre, post
        info = np.array([[ih, iw, scale]], np.float32)
        blob, deltas = proposal_layer(cls_prob.copy(), bbox.copy(), info, "TEST")
        out["prop_%s_cfg" % tag] = np.array([seed, H, W, ih, iw, pre, post], np.int64)
        out["prop_%s_scale" % tag] = np.float32(scale)
        out["prop_%s_blob" % tag] = blob
        out["prop_%s_deltas" % tag] = deltas
        print("proposal", tag, blob.shape)
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = 12000, 1000

    assert nms_wrapper.nms(np.zeros((0, 5), np.float32), 0.7) == []
    for tag, seed, n, ctpn_like, thresh in NMS_CASES:
        dets = synth.make_boxes(seed, n, ctpn_like=ctpn_like)
        keep = nms_wrapper.nms(dets.copy(), thresh)
        out["nms_%s_cfg" % tag] = np.array([seed, n, int(ctpn_like)], np.int64)
        out["nms_%s_thresh" % tag] = np.float64(thresh)
        out["nms_%s_keep" % tag] = np.asarray(keep, np.int64)
        print("nms", tag, len(keep))

    for mode in ("H", "O"):
        cfg.TEST.DETECT_MODE = mode
        for seed in TEXT_CASES:
            tp, sc = synth.make_text_proposals(seed)
            recs = TextDetector().detect(tp.copy(), sc.copy(), (600, 900))
            out["text_%s_%d" % (mode, seed)] = np.asarray(recs, np.float64)
            print("text", mode, seed, recs.shape)
    path = os.path.join(HERE, "reference_postproc.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
