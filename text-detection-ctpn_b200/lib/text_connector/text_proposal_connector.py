"""Horizontal text-line fitting (lib/text_connector/text_proposal_connector.py:21-64)."""
import numpy as np

from .other import clip_boxes
from .text_proposal_graph_builder import TextProposalGraphBuilder


def fit_y(X, Y, x1, x2):
    # a single distinct X gives the horizontal line y = Y[0] (connector.py:13-19)
    if np.sum(X == X[0]) == len(X):
        return Y[0], Y[0]
    p = np.poly1d(np.polyfit(X, Y, 1))
    return p(x1), p(x2)


class TextProposalConnector:
    def __init__(self):
        self.graph_builder = TextProposalGraphBuilder()

    def group_text_proposals(self, text_proposals, scores, im_size):
        return self.graph_builder.build_graph(text_proposals, scores, im_size).sub_graphs_connected()

    def fit_y(self, X, Y, x1, x2):
        return fit_y(X, Y, x1, x2)

    def get_text_lines(self, text_proposals, scores, im_size):
        tp_groups = self.group_text_proposals(text_proposals, scores, im_size)
        text_lines = np.zeros((len(tp_groups), 5), np.float32)
        for index, tp_indices in enumerate(tp_groups):
            b = text_proposals[list(tp_indices)]
            x0 = np.min(b[:, 0])
            x1 = np.max(b[:, 2])
            offset = (b[0, 2] - b[0, 0]) * 0.5
            lt_y, rt_y = fit_y(b[:, 0], b[:, 1], x0 + offset, x1 - offset)
            lb_y, rb_y = fit_y(b[:, 0], b[:, 3], x0 + offset, x1 - offset)
            score = scores[list(tp_indices)].sum() / float(len(tp_indices))
            text_lines[index] = (x0, min(lt_y, rt_y), x1, max(lb_y, rb_y), score)
        text_lines = clip_boxes(text_lines, im_size)
        text_recs = np.zeros((len(text_lines), 9), np.float64)
        if len(text_lines):
            xmin, ymin, xmax, ymax = text_lines[:, 0], text_lines[:, 1], text_lines[:, 2], text_lines[:, 3]
            text_recs[:, 0], text_recs[:, 1], text_recs[:, 2], text_recs[:, 3] = xmin, ymin, xmax, ymin
            text_recs[:, 4], text_recs[:, 5], text_recs[:, 6], text_recs[:, 7] = xmin, ymax, xmax, ymax
            text_recs[:, 8] = text_lines[:, 4]
        return text_recs
