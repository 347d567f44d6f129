"""ctpn/demo.py mirror against outputs of the reference's own draw_boxes / resize_im (tests/golden/make_golden_draw.py
executes the two functions straight from the unmodified reference source): result-file bytes, annotated image, and the
resize rule + pixels.  Also pins oracle/resize.py against the reference's resize_im output."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import resize as R  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_draw_boxes.npz"))


@pytest.mark.parametrize("k", range(4))
def test_draw_boxes_matches_reference(k, tmp_path, monkeypatch):
    from ctpn import demo
    monkeypatch.setattr(demo, "RESULTS_DIR", str(tmp_path))
    img = np.full((600, 900, 3), 40 * (k + 1), np.uint8)
    img[::50] = 255 - 40 * (k + 1)
    demo.draw_boxes(img, "some/dir/pic_%d.png" % k, GOLD["boxes_%d" % k], float(GOLD["scale_%d" % k]))
    txt = open(tmp_path / ("res_pic_%d.txt" % k), "rb").read()
    assert txt == GOLD["res_%d" % k].tobytes()
    assert txt.count(b"\r\n") == len(GOLD["boxes_%d" % k]) - 1          # the first box trips the |x1 - y1| < 5 skip
    np.testing.assert_array_equal(cv2.imread(str(tmp_path / ("pic_%d.png" % k))), GOLD["image_%d" % k])


@pytest.mark.parametrize("k", range(5))
def test_resize_im_matches_reference(k):
    from ctpn import demo
    h, w = [int(v) for v in GOLD["resize_shape_%d" % k]]
    im = np.random.RandomState(90 + k).randint(0, 256, (h, w, 3)).astype(np.uint8)
    out, f = demo.resize_im(im, scale=120, max_scale=240)
    assert f == float(GOLD["resize_f_%d" % k]) == R.resize_im_scale(h, w, 120, 240)
    np.testing.assert_array_equal(out, GOLD["resize_out_%d" % k])
    np.testing.assert_array_equal(R.resize_linear_u8(im, f), GOLD["resize_out_%d" % k])   # the oracle, without cv2 in the loop


BLOB = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_image_blob.npz"))


@pytest.mark.parametrize("k", range(5))
def test_get_image_blob_matches_reference(k):
    """lib/fast_rcnn/test.py mirror against the reference's _get_image_blob (tests/golden/make_golden_blob.py), with the
    same reduced TEST.SCALES / MAX_SIZE.  At scale 1 the mirror hands the uint8 image through (conv1_1 subtracts the
    means on the device as float32(double(v) - mean)); that arithmetic must reproduce the reference blob bit for bit."""
    from lib.fast_rcnn.config import cfg
    from lib.fast_rcnn import test as T
    old = cfg.TEST.SCALES, cfg.TEST.MAX_SIZE
    try:
        cfg.TEST.SCALES, cfg.TEST.MAX_SIZE = tuple(int(v) for v in BLOB["scales"]), int(BLOB["max_size"])
        h, w = [int(v) for v in BLOB["shape_%d" % k]]
        im = np.random.RandomState(300 + k).randint(0, 256, (h, w, 3)).astype(np.uint8)
        blob, factors = T._get_image_blob(im)
    finally:
        cfg.TEST.SCALES, cfg.TEST.MAX_SIZE = old
    np.testing.assert_array_equal(factors, BLOB["factors_%d" % k])
    ref = BLOB["blob_%d" % k]
    if blob.dtype == np.uint8:
        assert factors[0] == 1.0
        blob = (blob.astype(np.float64) - np.asarray(cfg.PIXEL_MEANS, np.float64)).astype(np.float32)
    assert blob.dtype == np.float32 and blob.shape == ref.shape
    np.testing.assert_array_equal(blob, ref)
