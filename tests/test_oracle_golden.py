"""CPU: pin the oracle (oracle/postproc.py, oracle/textline.py) against the
golden vectors produced by the reference's own code
(tests/golden/make_golden.py -> tests/golden/reference_postproc.npz)."""
import os

import numpy as np
import pytest

from oracle import postproc, synth, textline

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_postproc.npz"))
PROP_TAGS = sorted(k[5:-5] for k in G.files if k.startswith("prop_") and k.endswith("_blob"))
NMS_TAGS = sorted(k[4:-5] for k in G.files if k.startswith("nms_") and k.endswith("_keep"))


def test_anchor_table():
    np.testing.assert_array_equal(postproc.anchors(), G["anchors"])          # py3 table
    py2 = postproc.anchors(py2=True)
    assert py2[:, 1].tolist() == [2, 0, -3, -8, -16, -26, -40, -61, -91, -133]
    assert py2[:, 3].tolist() == [12, 15, 18, 23, 31, 41, 55, 76, 106, 148]


@pytest.mark.parametrize("tag", PROP_TAGS)
def test_proposal_layer_matches_reference(tag):
    seed, H, W, ih, iw, pre, post = (int(v) for v in G["prop_%s_cfg" % tag])
    scale = float(G["prop_%s_scale" % tag])
    cls_prob, bbox = synth.make_head_outputs(seed, H, W)
    info = np.array([[ih, iw, scale]], np.float32)
    # numpy-exp mode reproduces the reference bit for bit
    blob, deltas = postproc.proposal_layer(cls_prob, bbox, info, pre, post, exp_mode="numpy")
    np.testing.assert_array_equal(blob, G["prop_%s_blob" % tag])
    np.testing.assert_array_equal(deltas, G["prop_%s_deltas" % tag])
    # canonical (correctly rounded exp) mode: same rows selected, coordinates within 1 ulp
    blob_r, _ = postproc.proposal_layer(cls_prob, bbox, info, pre, post, exp_mode="rounded")
    assert blob_r.shape == blob.shape
    np.testing.assert_array_equal(blob_r[:, 0], blob[:, 0])
    np.testing.assert_allclose(blob_r, blob, rtol=3e-7, atol=1e-4)


@pytest.mark.parametrize("tag", NMS_TAGS)
def test_nms_matches_reference(tag):
    seed, n, ctpn_like = (int(v) for v in G["nms_%s_cfg" % tag])
    dets = synth.make_boxes(seed, n, ctpn_like=bool(ctpn_like))
    keep = postproc.nms(dets, float(G["nms_%s_thresh" % tag]))
    np.testing.assert_array_equal(np.asarray(keep, np.int64), G["nms_%s_keep" % tag])


def test_nms_empty():
    assert postproc.nms(np.zeros((0, 5), np.float32), 0.7) == []


@pytest.mark.parametrize("mode", ["H", "O"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_text_detector_matches_reference(mode, seed):
    tp, sc = synth.make_text_proposals(seed)
    recs = textline.detect(tp, sc, (600, 900), mode)
    ref = G["text_%s_%d" % (mode, seed)]
    assert recs.shape == ref.shape and recs.dtype == np.float64
    np.testing.assert_array_equal(recs, ref)


def test_column_decomposition_property():
    """CTPN proposals from different feature columns never
    suppress each other (IoU <= 1/15 < 0.2), so NMS == per-column NMS."""
    dets = synth.make_boxes(11, 1500, ctpn_like=True)
    keep = set(postproc.nms(dets, 0.2))
    per_col = set()
    for x in np.unique(dets[:, 0]):
        idx = np.where(dets[:, 0] == x)[0]
        per_col.update(int(idx[k]) for k in postproc.nms(dets[idx], 0.2))
    assert keep == per_col
