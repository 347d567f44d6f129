"""C++ text-line connector (ctpn_text_lines_host, SURVEY.md §8 f rank 3) against the oracle restatement of
TextDetector.detect and the golden outputs of the reference itself.  Pure host code: runs without a GPU.

Bar: identical line SETS (count and order); every value within 1e-4 px (one float32 ulp at 1000 px is 6e-5) and at
least 98 % of the values bit-identical -- the rest are 2-box lines, whose end points are evaluated exactly half-way
between two float32 numbers, where the last bit of LAPACK's double-precision solve decides the rounding in numpy."""
import ctypes as C
import os

import numpy as np
import pytest

from ctpn_b200 import _native as N
from ctpn_b200.textlines import text_lines
from oracle import synth, textline

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_postproc.npz"))
TOL = 1e-4


@pytest.mark.parametrize("mode", ["H", "O"])
def test_matches_reference_goldens(mode):
    for seed in range(4):
        tp, sc = synth.make_text_proposals(seed)
        ref = GOLD["text_%s_%d" % (mode, seed)]
        got = text_lines(tp, sc, (600, 900), mode)
        assert got.shape == ref.shape and got.dtype == np.float64
        assert np.abs(got - ref).max() <= TOL


@pytest.mark.parametrize("mode", ["H", "O"])
def test_matches_oracle_on_many_layouts(mode):
    total = exact = lines = 0
    worst = 0.0
    for seed in range(120):
        tp, sc = synth.make_text_proposals(100 + seed, n_lines=1 + seed % 14, n_noise=20 + 5 * (seed % 30))
        ref = textline.detect(tp, sc.reshape(-1, 1), (600, 900), mode)
        got = text_lines(tp, sc, (600, 900), mode)
        assert got.shape == ref.shape, "seed %d: %s lines vs %s" % (seed, got.shape, ref.shape)
        if ref.size:
            worst = max(worst, float(np.abs(got - ref).max()))
            total += ref.size
            exact += int((got == ref).sum())
            lines += len(ref)
    print("mode %s: %d lines, %d/%d values bit-exact, max |diff| %.2e" % (mode, lines, exact, total, worst))
    assert lines > 300 and worst <= TOL and exact >= 0.98 * total
    np.testing.assert_array_equal(got[:, 8], ref[:, 8]) if ref.size else None     # scores: numpy's pairwise sum, exact


def _round_f32(fr):
    """Correctly rounded (nearest, ties to even) float32 of an exact Fraction."""
    from fractions import Fraction
    x = np.float32(float(fr))                 # double rounding is harmless here: candidates are checked exactly below
    lo, hi = np.nextafter(x, np.float32(-np.inf)), np.nextafter(x, np.float32(np.inf))
    best = min((lo, x, hi), key=lambda c: (abs(Fraction(float(c)) - fr), int(np.float32(c).view(np.uint32)) & 1))
    return np.float32(best)


def test_native_fit_is_the_correctly_rounded_exact_fit_on_two_box_lines():
    """Where the all-C++ path and numpy's np.polyfit disagree (1 float32 ulp), the C++ value is the correctly rounded
    exact least-squares result and numpy's is not: for a 2-box chain the fitted line passes through both points, so the
    exact value at x is a rational number; it is checked here with fractions.Fraction."""
    from fractions import Fraction
    from ctpn_b200 import textlines
    checked = ties = numpy_off = 0
    for seed in range(120):
        tp, sc = synth.make_text_proposals(100 + seed, n_lines=1 + seed % 14, n_noise=20 + 5 * (seed % 30))
        keep = textlines.filter_nms(tp, sc)
        b, s = tp[keep], sc[keep]
        chains = textlines.groups(b, s, (600, 900))
        two = [c for c in chains if len(c) == 2 and b[c[0], 0] != b[c[1], 0]]
        if not two:
            continue
        native = textlines.text_lines(b, s, (600, 900), "H", (0.0, 2.0, 50, 0.7, 0.7, 0.0, 0.0, 0, 0))   # filters off: every chain comes back
        via_np = textlines.fit_lines(b, s, chains, (600, 900), "H")
        assert native.shape == via_np.shape == (len(chains), 9)
        for c in two:
            row = chains.index(c)
            g = b[c]
            left, right = g[:, 0].min(), g[:, 2].max()
            half = (g[0, 2] - g[0, 0]) * np.float32(0.5)
            xa, xb = Fraction(float(left + half)), Fraction(float(right - half))
            for col, ycol in ((1, 1), (5, 3)):                       # top edge -> min, bottom edge -> max
                x0, x1, y0, y1 = (Fraction(float(v)) for v in (g[0, 0], g[1, 0], g[0, ycol], g[1, ycol]))
                ends = [y0 + (y1 - y0) * (x - x0) / (x1 - x0) for x in (xa, xb)]
                exact = min(ends) if col == 1 else max(ends)
                want = _round_f32(exact)
                want = np.float32(min(max(want, np.float32(0)), np.float32(599)))
                assert np.float32(native[row, col]) == want, (seed, c, col)
                checked += 1
                ties += int(Fraction(float(want)) != exact and abs(Fraction(float(want)) - exact) * 2 == Fraction(float(np.spacing(want))))
                numpy_off += int(np.float32(via_np[row, col]) != want)
    print("2-box edges checked %d, exact float32 ties %d, numpy != correctly rounded %d" % (checked, ties, numpy_off))
    assert checked > 100


def test_other_image_sizes_and_mirror_class():
    from lib.fast_rcnn.config import cfg
    from lib.text_connector.detectors import TextDetector
    tp, sc = synth.make_text_proposals(7, im_h=900, im_w=600, n_lines=9)
    ref = textline.detect(tp, sc.reshape(-1, 1), (900, 600), "H")
    old = cfg.TEST.DETECT_MODE
    try:
        cfg.TEST.DETECT_MODE = "H"
        got = TextDetector(native=True).detect(tp, sc.reshape(-1, 1), (900, 600))
    finally:
        cfg.TEST.DETECT_MODE = old
    assert got.shape == ref.shape and np.abs(got - ref).max() <= TOL


def test_edge_cases_and_errors():
    empty = text_lines(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), (600, 900))
    assert empty.shape == (0, 9)
    tp, sc = synth.make_text_proposals(2)
    assert text_lines(tp, np.full_like(sc, 0.5), (600, 900)).shape == (0, 9)          # nothing above 0.7
    assert text_lines(tp[:1], sc[:1], (600, 900)).shape == (0, 9)                    # a single box is no line
    with pytest.raises(ValueError):
        text_lines(tp, sc[:-1], (600, 900))
    with pytest.raises(ValueError):
        text_lines(tp, sc, (600, 900), mode="X")
    with pytest.raises(RuntimeError):                                                # the reference raises IndexError here
        text_lines(tp, np.full_like(sc, 0.99), (600, 100))
    # custom constants: with a 0.99 line-score bar no line survives; with no NMS-surviving neighbours within 1 px neither
    base = (0.7, 0.2, 50, 0.7, 0.7, 0.5, 0.9, 16, 2)
    assert text_lines(tp, sc, (600, 900), cfg=base).shape == text_lines(tp, sc, (600, 900)).shape
    assert text_lines(tp, sc, (600, 900), cfg=base[:6] + (0.999,) + base[7:]).shape == (0, 9)
    assert text_lines(tp, sc, (600, 900), cfg=base[:2] + (1,) + base[3:]).shape == (0, 9)
    # output buffer too small: status + the number of lines found
    b = np.ascontiguousarray(tp, np.float32)
    s = np.ascontiguousarray(sc, np.float32).ravel()
    out = np.zeros((1, 9))
    num = C.c_int(0)
    rc = N.lib.ctpn_text_lines_host(b.ctypes.data, s.ctypes.data, len(s), 600, 900, 0, None, out.ctypes.data, 1, C.byref(num))
    assert rc != 0 and num.value == len(text_lines(tp, sc, (600, 900))) > 1
    assert b"lines found" in N.lib.ctpn_last_error()


def test_fuzz_random_layouts_sizes_and_ties():
    """Unstructured inputs: arbitrary image sizes, float x1 (not column aligned), clustered rows, tied scores.  The line
    SETS must agree exactly; coordinates within 2 float32 ulp at the largest coordinate (1700 px -> 2.5e-4)."""
    rs = np.random.RandomState(123)
    lines = 0
    for t in range(200):
        h, w = int(rs.randint(100, 1300)), int(rs.randint(100, 1700))
        n = int(rs.randint(0, 300))
        x1 = rs.uniform(0, w - 17, n) if rs.rand() < 0.5 else 16.0 * rs.randint(0, (w - 17) // 16 + 1, n)
        yc, hh = rs.uniform(10, h - 10, n), rs.uniform(6, 80, n)
        if rs.rand() < 0.6 and n:
            rows = rs.uniform(20, h - 20, max(1, n // 20))
            yc, hh = rows[rs.randint(0, len(rows), n)] + rs.normal(0, 2, n), 20 + rs.normal(0, 1.5, n)
        b = np.stack([x1, yc - hh / 2, x1 + 16, yc + hh / 2], 1).astype(np.float32)
        b[:, 0::2] = np.clip(b[:, 0::2], 0, w - 1)
        b[:, 1::2] = np.clip(b[:, 1::2], 0, h - 1)
        s = rs.uniform(0.6, 1.0, n).astype(np.float32)
        if rs.rand() < 0.3:
            s = np.round(s, 2).astype(np.float32)
        for mode in ("H", "O"):
            ref = textline.detect(b, s.reshape(-1, 1), (h, w), mode)
            got = text_lines(b, s, (h, w), mode)
            assert got.shape == ref.shape, (t, mode)
            if ref.size:
                assert np.abs(got - ref).max() <= 3e-4, (t, mode)
                lines += len(ref)
    assert lines > 200
