"""Oracle (test infrastructure; never imported by the product): numpy restatement of cv2.resize(..., INTER_LINEAR) for
uint8 images, the arithmetic behind the reference's resize_im (ctpn/demo.py:21-25) and the final un-scaling in
draw_boxes (:50).  OpenCV (opencv-python, un-vendored dependency; 4.13.0 in this image) computes it in fixed point:

  * dst size = cvRound(src * f) (round half to even);
  * per destination column: fx = float((dx + 0.5) / f - 0.5), sx = floor(fx), fx -= sx; at the image border the
    fraction is dropped (sx < 0 -> sx = 0, fx = 0; sx >= w - 1 -> sx = w - 1, fx = 0); weights
    cvRound((1 - fx) * 2048), cvRound(fx * 2048) as int16; horizontal pass in int32;
  * per destination row the same WITHOUT dropping the fraction at the border (the two taps are clamped to the
    same row instead), and dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  * a scale of exactly 1/2 in both directions is routed to INTER_AREA: rounded 2x2 mean, partial blocks at odd
    borders averaged over the pixels that exist.
Pinned against cv2.resize itself in tests/test_resize_cpu.py (bit-exact on every case)."""
import numpy as np


def cv_round(x):
    return np.rint(x).astype(np.int64)


def out_size(sh, sw, fx, fy):
    return int(cv_round(np.float64(sh) * np.float64(fy))), int(cv_round(np.float64(sw) * np.float64(fx)))


def _taps(dn, sn, scale, drop_border_fraction):
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if drop_border_fraction:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= sn - 1
        f[hi] = 0
        s[hi] = sn - 1
    w0 = cv_round(((np.float32(1.0) - f) * np.float32(2048)).astype(np.float32)).astype(np.int32)
    w1 = cv_round((f * np.float32(2048)).astype(np.float32)).astype(np.int32)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), w0, w1


def _area2(src):
    sh, sw = src.shape[:2]
    dh, dw = out_size(sh, sw, 0.5, 0.5)
    S = src.astype(np.int64)
    out = np.zeros((dh, dw) + src.shape[2:], np.uint8)
    for dy in range(dh):
        ys = [y for y in (2 * dy, 2 * dy + 1) if y < sh]
        for dx in range(dw):
            xs = [x for x in (2 * dx, 2 * dx + 1) if x < sw]
            blk = S[np.ix_(ys, xs)].reshape(len(ys) * len(xs), -1).sum(0)
            if len(ys) * len(xs) == 4:
                out[dy, dx] = (blk + 2) >> 2
            else:
                out[dy, dx] = np.clip(cv_round((blk.astype(np.float32) / np.float32(len(ys) * len(xs))).astype(np.float32)), 0, 255)
    return out


def _area2_fast(src):
    """Vectorised _area2 for full 2x2 blocks (even sizes); odd borders fall back to the loop."""
    sh, sw = src.shape[:2]
    if sh % 2 or sw % 2:
        return _area2(src)
    S = src.astype(np.int32)
    return ((S[0::2, 0::2] + S[0::2, 1::2] + S[1::2, 0::2] + S[1::2, 1::2] + 2) >> 2).astype(np.uint8)


def resize_linear_u8(src, fx, fy=None):
    """cv2.resize(src, None, None, fx=fx, fy=fy, interpolation=cv2.INTER_LINEAR) for uint8 [H,W] or [H,W,C]."""
    fy = fx if fy is None else fy
    src = np.asarray(src)
    assert src.dtype == np.uint8
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    sh, sw = src.shape[:2]
    if 1.0 / fx == 2.0 and 1.0 / fy == 2.0:
        out = _area2_fast(src)
        return out[:, :, 0] if squeeze else out
    dh, dw = out_size(sh, sw, fx, fy)
    sx, sx1, ax0, ax1 = _taps(dw, sw, 1.0 / fx, True)
    sy, sy1, ay0, ay1 = _taps(dh, sh, 1.0 / fy, False)
    S = src.astype(np.int32)
    rows = S[:, sx] * ax0[None, :, None] + S[:, sx1] * ax1[None, :, None]
    out = (((ay0[:, None, None] * (rows[sy] >> 4)) >> 16) + ((ay1[:, None, None] * (rows[sy1] >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def _taps_f32(dn, sn, scale, drop_border_fraction):
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if drop_border_fraction:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= sn - 1
        f[hi] = 0
        s[hi] = sn - 1
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), (np.float32(1.0) - f).astype(np.float32), f


def resize_linear_f32(src, fx, fy=None):
    """cv2.resize(src, None, None, fx=fx, fy=fy, interpolation=cv2.INTER_LINEAR) for float32 [H,W,C] images as OpenCV's OWN
    code computes it (the second rescale of _get_image_blob, lib/fast_rcnn/test.py:17-25, runs on the float32
    mean-subtracted image): float weights (1 - f, f), horizontal pass S[sx] * a0 + S[sx+1] * a1, vertical pass
    R0 * b0 + R1 * b1, every product and sum rounded to float32; an exact 1/2 scale is routed to INTER_AREA,
    (((s00 + s01) + s10) + s11) * 0.25.  Pinned against cv2 with cv2.ipp.setUseIPP(False) in tests/test_resize_cpu.py;
    IPP-dispatching builds (the opencv-python wheel) differ from this by up to ~1.4e-2 on 8-bit-range data (1100-px-wide image;
    the gap grows with the width)."""
    fy = fx if fy is None else fy
    src = np.asarray(src)
    assert src.dtype == np.float32 and src.ndim == 3
    sh, sw = src.shape[:2]
    dh, dw = out_size(sh, sw, fx, fy)
    if 1.0 / fx == 2.0 and 1.0 / fy == 2.0:
        out = np.zeros((dh, dw, src.shape[2]), np.float32)
        for dy in range(dh):
            ys = [y for y in (2 * dy, 2 * dy + 1) if y < sh]
            for dx in range(dw):
                xs = [x for x in (2 * dx, 2 * dx + 1) if x < sw]
                vals = [src[y, x] for y in ys for x in xs]
                acc = vals[0].copy()
                for v in vals[1:]:
                    acc = (acc + v).astype(np.float32)
                out[dy, dx] = acc * np.float32(0.25) if len(vals) == 4 else acc / np.float32(len(vals))
        return out
    sx, sx1, ax0, ax1 = _taps_f32(dw, sw, 1.0 / fx, True)
    sy, sy1, ay0, ay1 = _taps_f32(dh, sh, 1.0 / fy, False)
    rows = (src[:, sx] * ax0[None, :, None]).astype(np.float32) + (src[:, sx1] * ax1[None, :, None]).astype(np.float32)
    return ((rows[sy] * ay0[:, None, None]).astype(np.float32) + (rows[sy1] * ay1[:, None, None]).astype(np.float32)).astype(np.float32)


def resize_im_scale(h, w, scale=600, max_scale=1200):
    """The factor resize_im (ctpn/demo.py:21-25) applies: short side -> scale unless the long side would exceed max_scale."""
    f = float(scale) / min(h, w)
    if max_scale is not None and f * max(h, w) > max_scale:
        f = float(max_scale) / max(h, w)
    return f
