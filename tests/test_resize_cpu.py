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
