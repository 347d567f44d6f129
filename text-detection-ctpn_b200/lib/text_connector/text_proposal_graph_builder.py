"""TextProposalGraphBuilder (lib/text_connector/text_proposal_graph_builder.py:6-78), with the
per-proposal Python scans replaced by dense float32 array operations.  Semantics kept:
  * buckets are int(x1) (truncation), scans look <= MAX_HORIZONTAL_GAP px to the right / left and
    stop at the first bucket that holds any compatible box (:10-32);
  * compatibility = vertical overlap >= 0.7 and height similarity >= 0.7 in float32 (:40-54);
  * successor = best-scoring box of that bucket, first one on ties (:72);
  * the edge is kept iff score[i] >= max score of the successor's nearest precursors (:34-38).
Comparisons against the python-float thresholds run in float32 (NumPy >= 2 weak-scalar promotion); the reference on
NumPy 1.x compared float32 scalars in float64, which differs only when a ratio equals float32(0.7) exactly.  The goldens
were generated under NumPy 2."""
import numpy as np

from .text_connect_cfg import Config as TextLineCfg
from .other import Graph


class TextProposalGraphBuilder:
    def build_graph(self, text_proposals, scores, im_size):
        tp = np.asarray(text_proposals)
        n = tp.shape[0]
        graph = np.zeros((n, n), bool)
        if n == 0:
            return Graph(graph)
        sc = np.asarray(scores).reshape(-1)
        im_w = im_size[1]
        key = tp[:, 0].astype(np.int64)                      # int(box[0])
        if key.min() < -im_w or key.max() >= im_w:
            raise IndexError("list index out of range")      # boxes_table[int(box[0])], :62-64
        key = np.where(key < 0, key + im_w, key)             # python negative indexing of boxes_table
        heights = tp[:, 3] - tp[:, 1] + 1
        y0 = np.maximum(tp[None, :, 1], tp[:, None, 1])
        y1 = np.minimum(tp[None, :, 3], tp[:, None, 3])
        hmin = np.minimum(heights[None, :], heights[:, None])
        hmax = np.maximum(heights[None, :], heights[:, None])
        ov = np.maximum(y1 - y0 + 1, 0) / hmin
        compat = (ov >= TextLineCfg.MIN_V_OVERLAPS) & (hmin / hmax >= TextLineCfg.MIN_SIZE_SIM)
        x_int = tp[:, 0].astype(np.int64)                    # un-wrapped int(box[0]) drives the scan ranges
        gap = TextLineCfg.MAX_HORIZONTAL_GAP
        # successors of i: buckets x_i+1 .. min(x_i+gap, im_w-1)
        d = key[None, :] - x_int[:, None]
        cand = compat & (d >= 1) & (d <= gap) & (key[None, :] < im_w)
        big = np.iinfo(np.int64).max
        dmin = np.where(cand, d, big).min(axis=1)
        has_succ = dmin != big
        succ_set = cand & (d == dmin[:, None])
        succ = np.where(succ_set, sc[None, :], -np.inf).argmax(axis=1)
        # precursors of s: buckets x_s-1 down to max(int(x_s - gap), 0)
        lo = np.maximum((tp[:, 0] - gap).astype(np.int64), 0)
        e = x_int[:, None] - key[None, :]
        pcand = compat & (e >= 1) & (key[None, :] >= lo[:, None])
        emin = np.where(pcand, e, big).min(axis=1)
        pset = pcand & (e == emin[:, None])
        pmax = np.where(pset, sc[None, :], -np.inf).max(axis=1)
        idx = np.where(has_succ)[0]
        ok = sc[idx] >= pmax[succ[idx]]
        graph[idx[ok], succ[idx[ok]]] = True
        return Graph(graph)
