"""Text lines from proposals (TextDetector.detect, lib/text_connector/detectors.py:19-49) on top of the host C++
connector of the library (csrc/textline.cu).  Two ways to the same lines:

  detect_lines()  the drop-in path behind lib.text_connector.TextDetector: score filter + order + NMS 0.2, proposal
                  graph and chains in C++ (ctpn_text_filter_nms_host, ctpn_text_groups_host -- the O(n^2) Python loops
                  of the reference), line fitting here with numpy so that np.polyfit's own LAPACK solve produces the
                  coordinates: bit-identical to the reference's output on the same machine.
  text_lines()    everything in C++ (ctpn_text_lines_host), 0.1-0.4 ms per image, for the throughput pipeline
                  (Engine.detect_lines_batches).  Same line sets; coordinates are the correctly rounded least-squares fit
                  and differ from numpy's by at most one float32 ulp where LAPACK's ~1e-16 noise breaks an exact float32
                  tie (2-box lines; ~1 % of values).  np.polyfit's last bit depends on the LAPACK build, so bit-equality
                  with it is not a portable target; tests/test_textline_cpu.py checks the tie claim with exact rationals.

Assumes NumPy >= 2 scalar promotion (python floats are weak: float32 arrays stay float32), which is what the goldens were
generated under."""
import ctypes as C

import numpy as np

from . import _native as N

MAX_LINES = 4096
# (min_score, nms_thresh, max_gap, min_v_overlaps, min_size_sim, min_ratio, line_min_score, proposal_width, min_num)
DEFAULT_CFG = (0.7, 0.2, 50, 0.7, 0.7, 0.5, 0.9, 16, 2)


def _cfg_ptr(cfg):
    if cfg is None:
        return None, None
    c = np.ascontiguousarray(cfg, np.float32).reshape(9)
    return c, c.ctypes.data


def _as_inputs(text_proposals, scores):
    b = np.ascontiguousarray(text_proposals, np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(scores, np.float32).reshape(-1)
    if s.shape[0] != b.shape[0]:
        raise ValueError("%d proposals but %d scores" % (b.shape[0], s.shape[0]))
    return b, s


def text_lines(text_proposals, scores, size, mode="H", cfg=None):
    """text_proposals [n,4] float32, scores [n] or [n,1] float32 (test_ctpn output), size = (h, w) -> float64 [m,9]
    rows (x1,y1,x2,y2,x3,y3,x4,y4,score) like TextDetector.detect.  cfg: optional 9-tuple (see DEFAULT_CFG)."""
    if mode not in ("H", "O"):
        raise ValueError("mode must be 'H' or 'O' (got %r)" % (mode,))
    b, s = _as_inputs(text_proposals, scores)
    _keep, cptr = _cfg_ptr(cfg)
    out = np.empty((min(MAX_LINES, max(b.shape[0], 1)), 9), np.float64)     # a line needs >= 2 proposals
    num = C.c_int(0)
    N.check(N.lib.ctpn_text_lines_host(b.ctypes.data, s.ctypes.data, b.shape[0], int(size[0]), int(size[1]),
                                       1 if mode == "O" else 0, cptr, out.ctypes.data, out.shape[0], C.byref(num)),
            "ctpn_text_lines_host")
    return out[:num.value].copy()


def filter_nms(text_proposals, scores, cfg=None):
    """detectors.py:21-28 on the host: indices (into the input) of the proposals with score > min_score that survive
    the greedy NMS, in visiting order (score descending, index ascending on ties)."""
    b, s = _as_inputs(text_proposals, scores)
    _keep, cptr = _cfg_ptr(cfg)
    keep = np.empty(max(b.shape[0], 1), np.int32)
    num = C.c_int(0)
    N.check(N.lib.ctpn_text_filter_nms_host(b.ctypes.data, s.ctypes.data, b.shape[0], cptr, keep.ctypes.data, C.byref(num)),
            "ctpn_text_filter_nms_host")
    return keep[:num.value].astype(np.int64)


def groups(text_proposals, scores, im_size, cfg=None):
    """Proposal graph + chain walk (text_proposal_graph_builder.py:56-78, other.py:16-29): list of index lists, one per
    chain, in the reference's order (chains by ascending first node, members left to right)."""
    b, s = _as_inputs(text_proposals, scores)
    _keep, cptr = _cfg_ptr(cfg)
    m = b.shape[0]
    offsets = np.zeros(m + 1, np.int32)
    num, total = C.c_int(0), C.c_int(0)
    capacity = 2 * m + 16                     # chains that merge share their tails: the member total can exceed m
    while True:
        members = np.empty(capacity, np.int32)
        rc = N.lib.ctpn_text_groups_host(b.ctypes.data, s.ctypes.data, m, int(im_size[1]), cptr, offsets.ctypes.data,
                                         members.ctypes.data, capacity, C.byref(num), C.byref(total))
        if rc == N.ERR_WORKSPACE and total.value > capacity:
            capacity = total.value
            continue
        if rc and "outside the image width" in N.last_error():
            raise IndexError("list index out of range")      # boxes_table[int(box[0])], graph_builder.py:62-64
        N.check(rc, "ctpn_text_groups_host")
        break
    return [members[offsets[g]:offsets[g + 1]].tolist() for g in range(num.value)]


# ---- numpy line fitting (the reference's own arithmetic: float32 boxes, float64 np.polyfit) ------------------------
def _edge_at(xs, ys, xa, xb):
    """text_proposal_connector.py:13-19: ordinates of the fitted edge at xa and xb (float64 unless degenerate)."""
    if (xs == xs[0]).all():
        return ys[0], ys[0]
    line = np.poly1d(np.polyfit(xs, ys, 1))
    return line(xa), line(xb)


def _chain_table(boxes, scores, chains, centre_line):
    """One float32 row per chain: (x_left, y_top, x_right, y_bottom, mean score[, slope, intercept, mean height + 2.5])."""
    table = np.zeros((len(chains), 8 if centre_line else 5), np.float32)
    for row, members in zip(table, chains):
        g = boxes[members]
        left, right = g[:, 0].min(), g[:, 2].max()
        half = (g[0, 2] - g[0, 0]) * 0.5
        tops = _edge_at(g[:, 0], g[:, 1], left + half, right - half)
        bottoms = _edge_at(g[:, 0], g[:, 3], left + half, right - half)
        row[:5] = (left, min(tops), right, max(bottoms), scores[members].sum() / float(len(members)))
        if centre_line:
            row[5:7] = np.polyfit((g[:, 0] + g[:, 2]) / 2, (g[:, 1] + g[:, 3]) / 2, 1)
            row[7] = np.mean(g[:, 3] - g[:, 1]) + 2.5
    return table


def fit_lines(text_proposals, scores, chains, im_size, mode="H"):
    """get_text_lines of the H connector (text_proposal_connector.py:21-64: axis-aligned, clipped) or the O connector
    (text_proposal_connector_oriented.py:24-105: skew-compensated parallelogram, not clipped) -> float64 [L,9]."""
    boxes = np.asarray(text_proposals)
    scores = np.asarray(scores)
    table = _chain_table(boxes, scores, chains, mode == "O")
    recs = np.zeros((len(table), 9), np.float64)
    if not len(table):
        return recs
    if mode == "H":
        table[:, 0::2] = np.maximum(np.minimum(table[:, 0::2], im_size[1] - 1), 0)     # other.py:7-13 (clips the score column too)
        table[:, 1::2] = np.maximum(np.minimum(table[:, 1::2], im_size[0] - 1), 0)
        recs[:, [0, 4]] = table[:, [0]]
        recs[:, [2, 6]] = table[:, [2]]
        recs[:, [1, 3]] = table[:, [1]]
        recs[:, [5, 7]] = table[:, [3]]
        recs[:, 8] = table[:, 4]
        return recs
    xl, xr, slope, icpt, height = table[:, 0], table[:, 2], table[:, 5], table[:, 6], table[:, 7]
    upper, lower = icpt - height / 2, icpt + height / 2
    corners = np.stack([xl, slope * xl + upper, xr, slope * xr + upper, xl, slope * xl + lower, xr, slope * xr + lower], 1)
    run, rise = corners[:, 2] - corners[:, 0], corners[:, 3] - corners[:, 1]
    length = np.sqrt(run * run + rise * rise)
    proj = (corners[:, 5] - corners[:, 1]) * rise / length
    sx, sy = np.fabs(proj * run / length), np.fabs(proj * rise / length)
    down = slope < 0
    shift = np.zeros_like(corners)                       # float32: the corner updates stay in float32 like the scalars
    shift[:, 0], shift[:, 1], shift[:, 6], shift[:, 7] = -sx, sy, sx, -sy
    up = np.zeros_like(corners)
    up[:, 2], up[:, 3], up[:, 4], up[:, 5] = sx, sy, -sx, -sy
    recs[:, :8] = corners + np.where(down[:, None], shift, up)
    recs[:, 8] = table[:, 4]
    return recs


def keep_lines(recs, cfg=None):
    """filter_boxes (detectors.py:37-49): indices of lines with w/h > min_ratio, score > line_min_score, w > width*min_num."""
    c = DEFAULT_CFG if cfg is None else cfg
    recs = np.asarray(recs, np.float64).reshape(-1, 9)
    h = (np.abs(recs[:, 5] - recs[:, 1]) + np.abs(recs[:, 7] - recs[:, 3])) / 2.0 + 1
    w = (np.abs(recs[:, 2] - recs[:, 0]) + np.abs(recs[:, 6] - recs[:, 4])) / 2.0 + 1
    return np.where((w / h > c[5]) & (recs[:, 8] > c[6]) & (w > c[7] * c[8]))[0]


def detect_lines(text_proposals, scores, size, mode="H", cfg=None):
    """TextDetector.detect with numpy's own fit: proposals [n,4] f32, scores [n,1] f32 -> float64 [m,9]."""
    keep = filter_nms(text_proposals, scores, cfg)
    tp, sc = np.asarray(text_proposals)[keep], np.asarray(scores).reshape(-1, 1)[keep]
    recs = fit_lines(tp, sc, groups(tp, sc, size, cfg), size, mode)
    return recs[keep_lines(recs, cfg)]
