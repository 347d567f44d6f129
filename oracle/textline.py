"""Oracle (test infrastructure): loop-level restatement of the reference's
text-line construction.  Not imported by the product.

Follows (paths relative to /root/reference):
  detect()            lib/text_connector/detectors.py:19-35
  line_filter()       lib/text_connector/detectors.py:37-49
  build_edges()       lib/text_connector/text_proposal_graph_builder.py:10-78
  chains()            lib/text_connector/other.py:16-29
  lines_horizontal()  lib/text_connector/text_proposal_connector.py:13-64 (+ other.py:7-13)
  lines_oriented()    lib/text_connector/text_proposal_connector_oriented.py:24-105
Constants: lib/text_connector/text_connect_cfg.py:1-12.

Scalar/array dtype behaviour is that of numpy >= 2 (the interpreter this repo
runs on): python-float thresholds are compared in float32 against float32
operands.  Ties in the score sort follow oracle.postproc.order_desc.
"""
import numpy as np

from .postproc import F32, nms, order_desc

TEXT_PROPOSALS_WIDTH = 16
MIN_NUM_PROPOSALS = 2
MIN_RATIO = 0.5
LINE_MIN_SCORE = 0.9
MAX_HORIZONTAL_GAP = 50
TEXT_PROPOSALS_MIN_SCORE = 0.7
TEXT_PROPOSALS_NMS_THRESH = 0.2
MIN_V_OVERLAPS = 0.7
MIN_SIZE_SIM = 0.7


def _compatible(tp, heights, a, b):
    """meet_v_iou (graph_builder.py:40-54): vertical overlap and height similarity."""
    ha, hb = heights[a], heights[b]
    y0 = max(tp[b][1], tp[a][1])
    y1 = min(tp[b][3], tp[a][3])
    ov = max(0, y1 - y0 + 1) / min(ha, hb)
    sim = min(ha, hb) / max(ha, hb)
    return bool(ov >= MIN_V_OVERLAPS) and bool(sim >= MIN_SIZE_SIM)


def build_edges(tp, scores, im_size):
    """Returns bool adjacency [N,N]; graph_builder.py:56-78."""
    n = tp.shape[0]
    im_w = im_size[1]
    heights = tp[:, 3] - tp[:, 1] + 1
    table = [[] for _ in range(im_w)]
    for i in range(n):
        table[int(tp[i][0])].append(i)

    def successions(i):
        x = int(tp[i][0])
        for left in range(x + 1, min(x + MAX_HORIZONTAL_GAP + 1, im_w)):
            hit = [j for j in table[left] if _compatible(tp, heights, j, i)]
            if hit:
                return hit
        return []

    def precursors(i):
        x = int(tp[i][0])
        lo = max(int(tp[i][0] - MAX_HORIZONTAL_GAP), 0)
        for left in range(x - 1, lo - 1, -1):
            hit = [j for j in table[left] if _compatible(tp, heights, j, i)]
            if hit:
                return hit
        return []

    graph = np.zeros((n, n), bool)
    for i in range(n):
        succ = successions(i)
        if not succ:
            continue
        s = succ[int(np.argmax(scores[succ]))]
        prec = precursors(s)
        if scores[i] >= np.max(scores[prec]):
            graph[i, s] = True
    return graph


def chains(graph):
    """other.py:16-29: start at nodes with out-edges and no in-edge; follow the
    first out-edge only."""
    out = []
    for i in range(graph.shape[0]):
        if not graph[:, i].any() and graph[i, :].any():
            v = i
            out.append([v])
            while graph[v, :].any():
                v = int(np.where(graph[v, :])[0][0])
                out[-1].append(v)
    return out


def _fit_y(X, Y, x1, x2):
    if np.sum(X == X[0]) == len(X):
        return Y[0], Y[0]
    p = np.poly1d(np.polyfit(X, Y, 1))
    return p(x1), p(x2)


def lines_horizontal(tp, scores, im_size):
    groups = chains(build_edges(tp, scores, im_size))
    lines = np.zeros((len(groups), 5), F32)
    for k, g in enumerate(groups):
        b = tp[list(g)]
        x0 = np.min(b[:, 0])
        x1 = np.max(b[:, 2])
        off = (b[0, 2] - b[0, 0]) * 0.5
        lt, rt = _fit_y(b[:, 0], b[:, 1], x0 + off, x1 - off)
        lb, rb = _fit_y(b[:, 0], b[:, 3], x0 + off, x1 - off)
        sc = scores[list(g)].sum() / float(len(g))
        lines[k] = (x0, min(lt, rt), x1, max(lb, rb), sc)
    # other.py:7-13 clip (acts on the first four columns pairwise AND on the score
    # column, which sits at an even index: score is clipped to [0, w-1] -- harmless)
    lines[:, 0::2] = np.maximum(np.minimum(lines[:, 0::2], im_size[1] - 1), 0)
    lines[:, 1::2] = np.maximum(np.minimum(lines[:, 1::2], im_size[0] - 1), 0)
    recs = np.zeros((len(lines), 9), np.float64)
    for k, ln in enumerate(lines):
        xmin, ymin, xmax, ymax = ln[0], ln[1], ln[2], ln[3]
        recs[k] = (xmin, ymin, xmax, ymin, xmin, ymax, xmax, ymax, ln[4])
    return recs


def lines_oriented(tp, scores, im_size):
    groups = chains(build_edges(tp, scores, im_size))
    lines = np.zeros((len(groups), 8), F32)
    for k, g in enumerate(groups):
        b = tp[list(g)]
        X = (b[:, 0] + b[:, 2]) / 2
        Y = (b[:, 1] + b[:, 3]) / 2
        z1 = np.polyfit(X, Y, 1)
        x0 = np.min(b[:, 0])
        x1 = np.max(b[:, 2])
        off = (b[0, 2] - b[0, 0]) * 0.5
        lt, rt = _fit_y(b[:, 0], b[:, 1], x0 + off, x1 - off)
        lb, rb = _fit_y(b[:, 0], b[:, 3], x0 + off, x1 - off)
        sc = scores[list(g)].sum() / float(len(g))
        height = np.mean(b[:, 3] - b[:, 1])
        lines[k] = (x0, min(lt, rt), x1, max(lb, rb), sc, z1[0], z1[1], height + 2.5)
    recs = np.zeros((len(lines), 9), np.float64)
    for k, ln in enumerate(lines):
        b1 = ln[6] - ln[7] / 2
        b2 = ln[6] + ln[7] / 2
        x1, y1 = ln[0], ln[5] * ln[0] + b1
        x2, y2 = ln[2], ln[5] * ln[2] + b1
        x3, y3 = ln[0], ln[5] * ln[0] + b2
        x4, y4 = ln[2], ln[5] * ln[2] + b2
        dx, dy = x2 - x1, y2 - y1
        width = np.sqrt(dx * dx + dy * dy)
        t0 = y3 - y1
        t1 = t0 * dy / width
        x = np.fabs(t1 * dx / width)
        y = np.fabs(t1 * dy / width)
        if ln[5] < 0:
            x1 -= x; y1 += y; x4 += x; y4 -= y
        else:
            x2 += x; y2 += y; x3 -= x; y3 -= y
        recs[k] = (x1, y1, x2, y2, x3, y3, x4, y4, ln[4])
    return recs


def line_filter(recs):
    """detectors.py:37-49 (float64)."""
    if len(recs) == 0:
        return np.zeros((0,), np.int64)
    heights = (np.abs(recs[:, 5] - recs[:, 1]) + np.abs(recs[:, 7] - recs[:, 3])) / 2.0 + 1
    widths = (np.abs(recs[:, 2] - recs[:, 0]) + np.abs(recs[:, 6] - recs[:, 4])) / 2.0 + 1
    sc = recs[:, 8]
    return np.where((widths / heights > MIN_RATIO) & (sc > LINE_MIN_SCORE) &
                    (widths > TEXT_PROPOSALS_WIDTH * MIN_NUM_PROPOSALS))[0]


def detect(text_proposals, scores, size, mode="H"):
    """TextDetector.detect (detectors.py:19-35).  text_proposals [N,4] f32,
    scores [N,1] f32, size=(h,w).  Returns float64 [M,9]."""
    tp = np.asarray(text_proposals)
    sc = np.asarray(scores)
    keep = np.where(sc > TEXT_PROPOSALS_MIN_SCORE)[0]
    tp, sc = tp[keep], sc[keep]
    order = order_desc(sc.ravel())
    tp, sc = tp[order], sc[order]
    keep = nms(np.hstack((tp, sc)), TEXT_PROPOSALS_NMS_THRESH)
    tp, sc = tp[keep], sc[keep]
    recs = lines_horizontal(tp, sc, size) if mode == "H" else lines_oriented(tp, sc, size)
    return recs[line_filter(recs)]
