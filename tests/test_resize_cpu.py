"""Pin oracle/resize.py (restatement of OpenCV's fixed-point INTER_LINEAR for uint8) against cv2.resize itself -- the
arithmetic of the reference's resize_im (ctpn/demo.py:21-25).  Bit-exact on every case."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import resize as R  # noqa: E402


@pytest.mark.parametrize("seed", range(6))
def test_resize_linear_matches_cv2(seed):
    rs = np.random.RandomState(seed)
    for _ in range(25):
        h, w = int(rs.randint(20, 300)), int(rs.randint(20, 400))
        c = int(rs.choice([1, 3]))
        im = rs.randint(0, 256, (h, w, c)).astype(np.uint8)
        f = float(rs.choice([R.resize_im_scale(h, w), 0.75, 1.25, 1.5, 2.0, 0.8333333, 0.5, 0.25, 1.0, rs.uniform(0.3, 3.0)]))
        ref = cv2.resize(im, None, None, fx=f, fy=f, interpolation=cv2.INTER_LINEAR)
        got = R.resize_linear_u8(im, f)
        if c == 1:
            got = got[:, :, 0]
        assert got.shape == ref.shape, (h, w, f)
        np.testing.assert_array_equal(got, ref, err_msg="h=%d w=%d f=%r" % (h, w, f))


def test_resize_im_rule_and_anisotropic():
    assert R.resize_im_scale(600, 900) == 1.0                       # BASELINE configs: identity
    assert R.resize_im_scale(1200, 1600) == 0.5                     # config 4 would be halved by demo.py
    assert R.resize_im_scale(300, 2000) == 1200.0 / 2000            # long-side cap
    rs = np.random.RandomState(11)
    im = rs.randint(0, 256, (57, 91, 3)).astype(np.uint8)
    ref = cv2.resize(im, None, None, fx=1.7, fy=0.6, interpolation=cv2.INTER_LINEAR)
    np.testing.assert_array_equal(R.resize_linear_u8(im, 1.7, 0.6), ref)
    for shape in ((301, 203), (300, 203), (301, 202)):               # exact 1/2 with odd borders: INTER_AREA branch
        im = rs.randint(0, 256, shape + (3,)).astype(np.uint8)
        np.testing.assert_array_equal(R.resize_linear_u8(im, 0.5), cv2.resize(im, None, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR))


def test_float32_linear_matches_opencv_own_code_and_bounds_the_ipp_difference():
    """The float32 rescale of _get_image_blob (lib/fast_rcnn/test.py:17-25).  opencv-python dispatches float32 INTER_LINEAR
    to Intel IPP; with IPP switched off OpenCV runs its own resize.cpp code, which the restatement reproduces bit for bit.
    The IPP result differs from that through its float32 coordinate arithmetic: the gap grows with the image width and reaches
    ~1.4e-2 (on 8-bit-range data) for a 1100-px-wide image; asserted < 5e-2, i.e. 2e-4 of the pixel range."""
    import cv2
    from oracle.resize import resize_linear_f32
    means = np.array([102.9801, 115.9465, 122.7717])
    rs = np.random.RandomState(5)
    was = cv2.ipp.useIPP()
    try:
        for (h, w, f) in [(600, 1100, 1000.0 / 1100), (37, 53, 0.73), (300, 750, 1000.0 / 750), (128, 300, 2.0), (50, 70, 1.37),
                          (64, 64, 0.5), (65, 63, 0.5)]:
            im = rs.randint(0, 256, (h, w, 3)).astype(np.float32)
            im -= means
            mine = resize_linear_f32(im, f)
            cv2.ipp.setUseIPP(False)
            own = cv2.resize(im, None, None, fx=f, fy=f, interpolation=cv2.INTER_LINEAR)
            cv2.ipp.setUseIPP(True)
            ipp = cv2.resize(im, None, None, fx=f, fy=f, interpolation=cv2.INTER_LINEAR)
            np.testing.assert_array_equal(mine, own)
            assert np.abs(ipp - own).max() < 5e-2
    finally:
        cv2.ipp.setUseIPP(was)
