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
