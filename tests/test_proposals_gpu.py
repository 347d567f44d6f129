"""GPU: the fused proposal layer (ctpn_proposals behind lib.rpn_msr.proposal_layer_tf.proposal_layer)
against the reference-generated goldens and the CPU oracle.  Index order is checked exactly;
coordinates are bit-exact against the oracle's 'rounded' exp mode and within 1 ulp (<= 1e-4 px)
of the reference's numpy-exp output."""
import os

import numpy as np
import pytest

from oracle import postproc, synth

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_postproc.npz"))
PROP_TAGS = sorted(k[5:-5] for k in G.files if k.startswith("prop_") and k.endswith("_blob"))


@pytest.mark.parametrize("tag", PROP_TAGS)
def test_proposal_layer_matches_reference_golden(tag):
    from lib.fast_rcnn.config import cfg
    from lib.rpn_msr.proposal_layer_tf import proposal_layer
    seed, H, W, ih, iw, pre, post = (int(v) for v in G["prop_%s_cfg" % tag])
    scale = float(G["prop_%s_scale" % tag])
    cls_prob, bbox = synth.make_head_outputs(seed, H, W)
    old = cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = pre, post
    try:
        blob, deltas = proposal_layer(cls_prob, bbox, np.array([[ih, iw, scale]], np.float32), "TEST")
    finally:
        cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = old
    ref_blob, ref_deltas = G["prop_%s_blob" % tag], G["prop_%s_deltas" % tag]
    assert blob.shape == ref_blob.shape
    np.testing.assert_array_equal(blob[:, 0], ref_blob[:, 0])           # same proposals, same order
    np.testing.assert_array_equal(deltas, ref_deltas)
    np.testing.assert_allclose(blob, ref_blob, rtol=3e-7, atol=1e-4)     # north_star: 1e-3
    want, _ = postproc.proposal_layer(cls_prob, bbox, np.array([[ih, iw, scale]], np.float32), pre, post, exp_mode="rounded")
    np.testing.assert_array_equal(blob, want)                            # bit-exact vs canonical oracle


def test_batched_equals_loop_of_single_and_ties():
    """Batch of 4 images with heavy score ties (non-unique scores): per-image result equals the
    oracle's canonical order exactly, including the anchor index of every row."""
    import torch
    from ctpn_b200.engine import Engine
    eng = Engine(None)
    H, W = 19, 31
    cls, box = [], []
    for s in range(4):
        c, b = synth.make_head_outputs(100 + s, H, W, unique=False)
        c = (np.round(c * 64) / 64).astype(np.float32)                 # quantise -> many ties
        cls.append(c[0]); box.append(b[0])
    cls, box = np.stack(cls), np.stack(box)
    info = np.array([[H * 16, W * 16, 1.0]] * 4, np.float32)
    rois, index, count = eng.proposals(torch.from_numpy(cls).cuda(), torch.from_numpy(box).cuda(), torch.from_numpy(info),
                                       cls_is_logit=False, cfg=dict(RPN_PRE_NMS_TOP_N=2000, RPN_POST_NMS_TOP_N=300))
    for b in range(4):
        want, _, idx = postproc.proposal_layer(cls[b:b + 1], box[b:b + 1], info[b:b + 1], 2000, 300, return_index=True)
        n = int(count[b])
        assert n == want.shape[0]
        np.testing.assert_array_equal(index[b, :n].cpu().numpy(), idx)
        np.testing.assert_array_equal(rois[b, :n].cpu().numpy(), want)
        assert float(rois[b, n:].abs().sum()) == 0.0


def test_cfgB_sized_and_logit_input():
    """75x100 feature map (75 000 anchors, cfgB), logits in, pair softmax fused on the device."""
    import torch
    from ctpn_b200.engine import Engine
    eng = Engine(None)
    H, W = 75, 100
    rs = np.random.RandomState(5)
    logits = (rs.standard_normal((1, H, W, 20)) * 2).astype(np.float32)
    bbox = (rs.standard_normal((1, H, W, 40)) * 0.3).astype(np.float32)
    info = np.array([[1200, 1600, 1.0]], np.float32)
    rois, index, count = eng.proposals(torch.from_numpy(logits).cuda(), torch.from_numpy(bbox).cuda(), torch.from_numpy(info), cls_is_logit=True)
    n = int(count[0])
    idx = index[0, :n].cpu().numpy()
    got = rois[0, :n].cpu().numpy()
    # feed the device's own probabilities (softmax may differ from numpy's in the last ulp) to the oracle
    l = logits.reshape(-1, 2).astype(np.float64)
    p = np.exp(l - l.max(1, keepdims=True)); p = (p / p.sum(1, keepdims=True))
    assert np.abs(got[:, 0] - p[idx, 1]).max() < 1e-6
    assert n == 1000 and (np.diff(got[:, 0]) <= 0).all()
    assert got[:, 1:].min() >= 0 and got[:, [1, 3]].max() <= 1599 and got[:, [2, 4]].max() <= 1199


def test_generic_nms_forced_for_every_image():
    """CTPN_GENERIC_NMS=1 (a switch of the test library only) disables the column path altogether."""
    from test_conv_gpu import DBG, run_check
    run_check("proposals_generic", env=dict(DBG, CTPN_GENERIC_NMS="1"))


def test_unstructured_boxes_take_the_generic_nms_path():
    """im_info narrower than the feature map: boxes of many columns are clipped onto the same x range, so
    the column decomposition does not hold.  The device detects it per image and falls back to the generic
    bitmask NMS; results still equal the oracle exactly.  Image 1 of the batch is a normal (structured) one."""
    import torch
    from ctpn_b200.engine import Engine
    eng = Engine(None)
    H, W = 12, 18
    cls = np.concatenate([synth.make_head_outputs(200 + s, H, W)[0] for s in range(2)])
    box = np.concatenate([synth.make_head_outputs(200 + s, H, W)[1] for s in range(2)])
    info = np.array([[192, 100, 1.0], [192, 288, 1.0]], np.float32)
    rois, index, count = eng.proposals(torch.from_numpy(cls).cuda(), torch.from_numpy(box).cuda(), torch.from_numpy(info), cls_is_logit=False)
    for b in range(2):
        want, _, idx = postproc.proposal_layer(cls[b:b + 1], box[b:b + 1], info[b:b + 1], return_index=True)
        n = int(count[b])
        assert n == want.shape[0]
        np.testing.assert_array_equal(index[b, :n].cpu().numpy(), idx)
        np.testing.assert_array_equal(rois[b, :n].cpu().numpy(), want)
