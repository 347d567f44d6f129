"""GPU: ctpn_resize_linear_u8 (image front-end, SURVEY.md §8 f rank 2) against the oracle restatement of OpenCV's
fixed-point INTER_LINEAR and against cv2.resize itself when cv2 is importable.  Bit-exact."""
import numpy as np
import pytest

from oracle import resize as R, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from ctpn_b200 import Engine
    return Engine(synth.make_weights(0), planes=1)


def _reference(im, fx, fy):
    out = R.resize_linear_u8(im, fx, fy)
    try:
        import cv2
        np.testing.assert_array_equal(out, cv2.resize(im, None, None, fx=fx, fy=fy, interpolation=cv2.INTER_LINEAR).reshape(out.shape))
    except ImportError:
        pass
    return out


@pytest.mark.parametrize("shape,fx,fy", [((2, 37, 53, 3), 1.5, 1.5), ((1, 480, 640, 3), 1.25, 1.25), ((3, 301, 203, 3), 0.5, 0.5),
                                          ((2, 300, 202, 3), 0.5, 0.5), ((1, 90, 160, 3), 600 / 90.0, 600 / 90.0),
                                          ((2, 64, 48, 1), 1.7, 0.6), ((1, 1200, 1600, 3), 0.75, 0.75), ((1, 20, 30, 3), 1.0, 1.0)])
def test_resize_matches_opencv(engine, shape, fx, fy):
    rs = np.random.RandomState(5)
    ims = rs.randint(0, 256, shape).astype(np.uint8)
    got = engine.resize_images(ims, fx, fy).cpu().numpy()
    for b in range(shape[0]):
        ref = _reference(ims[b], fx, fy)
        assert got[b].shape == ref.shape
        np.testing.assert_array_equal(got[b], ref)


def test_detect_resized_equals_host_resize_then_detect(engine):
    """resize_im on the device + detector == cv2-exact host resize + detector (same proposals, bit for bit)."""
    rs = np.random.RandomState(9)
    ims = rs.randint(0, 256, (2, 240, 400, 3)).astype(np.uint8)
    res, f = engine.detect_resized(ims)
    assert f == R.resize_im_scale(240, 400) == 2.5
    host = np.stack([R.resize_linear_u8(im, f) for im in ims])
    assert host.shape == (2, 600, 1000, 3)
    want = engine.detect_batch(host)
    for (s0, b0), (s1, b1) in zip(res, want):
        np.testing.assert_array_equal(s0, s1)
        np.testing.assert_array_equal(b0, b1)
    with pytest.raises(ValueError):
        engine.resize_images(ims.astype(np.float32), 2.0)
    with pytest.raises(RuntimeError):
        from ctpn_b200 import _native as N
        import torch
        t = torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device="cuda")
        o = torch.zeros((1, 9, 9, 3), dtype=torch.uint8, device="cuda")     # cv2 would produce 12 x 12
        N.check(N.lib.ctpn_resize_linear_u8(N.ptr(t), 1, 8, 8, 3, 1.5, 1.5, N.ptr(o), 9, 9, N.stream_ptr()), "resize")


@pytest.mark.parametrize("h,w,f", [(600, 1100, 1000.0 / 1100), (37, 53, 0.73), (300, 750, 1000.0 / 750), (128, 300, 2.0), (64, 64, 0.5), (65, 63, 0.5)])
def test_image_blob_f32_bit_exact_with_the_oracle(h, w, f):
    """ctpn_image_blob_f32 (mean subtraction + float32 INTER_LINEAR, the second rescale of _get_image_blob) == the numpy
    restatement of OpenCV's own float code, which tests/test_resize_cpu.py pins against cv2 with IPP off."""
    from ctpn_b200 import Engine
    from oracle.resize import resize_linear_f32
    eng = Engine(None)
    rs = np.random.RandomState(h + w)
    ims = rs.randint(0, 256, (2, h, w, 3)).astype(np.uint8)
    got = eng.image_blob(ims, f).cpu().numpy()
    for b in range(2):
        im = ims[b].astype(np.float32)
        im -= np.array([[[102.9801, 115.9465, 122.7717]]])
        np.testing.assert_array_equal(got[b], resize_linear_f32(im, f))


def test_detect_scaled_equals_host_blob_path():
    """Engine.detect_scaled (rescale rule + blob on the device) == test_ctpn's host route fed with the same blob."""
    from ctpn_b200 import Engine
    from oracle import synth
    from oracle.resize import resize_linear_f32
    eng = Engine(synth.make_weights(0), planes=2)
    im = synth.make_image(3, 300, 560)                        # short side 300 -> x2 would give 1120 > 1000: scale = 1000/560
    scale = 1000.0 / 560
    src = im.astype(np.float32)
    src -= np.array([[[102.9801, 115.9465, 122.7717]]])
    blob = resize_linear_f32(src, scale)[None]
    want = eng.rois_batch(blob, np.array([[blob.shape[1], blob.shape[2], scale]], np.float32))[0]
    scores, boxes = eng.detect_scaled(im[None])[0]
    np.testing.assert_array_equal(scores, want[:, 0])
    np.testing.assert_array_equal(boxes, want[:, 1:5] / np.float32(scale))
