"""Text lines from proposals through the C++ connector of the library (ctpn_text_lines_host): the whole of
TextDetector.detect (lib/text_connector/detectors.py:19-49) in 0.1-0.4 ms per image instead of the 4-9 ms of
the Python connector.  Same line sets as the Python mirror / the reference; coordinates agree to float32 rounding
(2-box lines evaluate the fit exactly half-way between two float32 values, where LAPACK's last bit decides in numpy)."""
import ctypes as C

import numpy as np

from . import _native as N

MAX_LINES = 4096


def text_lines(text_proposals, scores, size, mode="H", cfg=None):
    """text_proposals [n,4] float32, scores [n] or [n,1] float32 (test_ctpn output), size = (h, w) -> float64 [m,9]
    rows (x1,y1,x2,y2,x3,y3,x4,y4,score) like TextDetector.detect.  cfg: optional 9-tuple (min_score, nms_thresh,
    max_gap, min_v_overlaps, min_size_sim, min_ratio, line_min_score, proposal_width, min_num_proposals)."""
    if mode not in ("H", "O"):
        raise ValueError("mode must be 'H' or 'O' (got %r)" % (mode,))
    b = np.ascontiguousarray(text_proposals, np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(scores, np.float32).reshape(-1)
    if s.shape[0] != b.shape[0]:
        raise ValueError("%d proposals but %d scores" % (b.shape[0], s.shape[0]))
    c = None if cfg is None else np.ascontiguousarray(cfg, np.float32).reshape(9)
    out = np.empty((min(MAX_LINES, max(b.shape[0], 1)), 9), np.float64)     # a line needs >= 2 proposals
    num = C.c_int(0)
    N.check(N.lib.ctpn_text_lines_host(b.ctypes.data, s.ctypes.data, b.shape[0], int(size[0]), int(size[1]),
                                       1 if mode == "O" else 0, None if c is None else c.ctypes.data,
                                       out.ctypes.data, out.shape[0], C.byref(num)), "ctpn_text_lines_host")
    return out[:num.value].copy()
