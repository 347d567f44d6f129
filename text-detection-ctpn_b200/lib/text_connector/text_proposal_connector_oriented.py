"""Oriented text-line fitting (lib/text_connector/text_proposal_connector_oriented.py:24-105):
centre-line fit, mean height + 2.5, skew-compensated parallelogram corners; no clipping."""
import numpy as np

from .text_proposal_connector import fit_y
from .text_proposal_graph_builder import TextProposalGraphBuilder


class TextProposalConnector:
    def __init__(self):
        self.graph_builder = TextProposalGraphBuilder()

    def group_text_proposals(self, text_proposals, scores, im_size):
        return self.graph_builder.build_graph(text_proposals, scores, im_size).sub_graphs_connected()

    def fit_y(self, X, Y, x1, x2):
        return fit_y(X, Y, x1, x2)

    def get_text_lines(self, text_proposals, scores, im_size):
        tp_groups = self.group_text_proposals(text_proposals, scores, im_size)
        text_lines = np.zeros((len(tp_groups), 8), np.float32)
        for index, tp_indices in enumerate(tp_groups):
            b = text_proposals[list(tp_indices)]
            X = (b[:, 0] + b[:, 2]) / 2
            Y = (b[:, 1] + b[:, 3]) / 2
            z1 = np.polyfit(X, Y, 1)
            x0 = np.min(b[:, 0])
            x1 = np.max(b[:, 2])
            offset = (b[0, 2] - b[0, 0]) * 0.5
            lt_y, rt_y = fit_y(b[:, 0], b[:, 1], x0 + offset, x1 - offset)
            lb_y, rb_y = fit_y(b[:, 0], b[:, 3], x0 + offset, x1 - offset)
            score = scores[list(tp_indices)].sum() / float(len(tp_indices))
            height = np.mean(b[:, 3] - b[:, 1])
            text_lines[index] = (x0, min(lt_y, rt_y), x1, max(lb_y, rb_y), score, z1[0], z1[1], height + 2.5)
        text_recs = np.zeros((len(text_lines), 9), np.float64)
        for index, line in enumerate(text_lines):
            b1 = line[6] - line[7] / 2
            b2 = line[6] + line[7] / 2
            x1, y1 = line[0], line[5] * line[0] + b1
            x2, y2 = line[2], line[5] * line[2] + b1
            x3, y3 = line[0], line[5] * line[0] + b2
            x4, y4 = line[2], line[5] * line[2] + b2
            disX, disY = x2 - x1, y2 - y1
            width = np.sqrt(disX * disX + disY * disY)
            fTmp1 = (y3 - y1) * disY / width
            x = np.fabs(fTmp1 * disX / width)
            y = np.fabs(fTmp1 * disY / width)
            if line[5] < 0:
                x1 -= x; y1 += y; x4 += x; y4 -= y
            else:
                x2 += x; y2 += y; x3 -= x; y3 -= y
            text_recs[index] = (x1, y1, x2, y2, x3, y3, x4, y4, line[4])
        return text_recs
