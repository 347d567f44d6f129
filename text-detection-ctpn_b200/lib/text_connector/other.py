"""Helpers of lib/text_connector/other.py: clip_boxes (:7-13) and the chain walk of
Graph.sub_graphs_connected (:16-29)."""
import numpy as np


def threshold(coords, min_, max_):
    return np.maximum(np.minimum(coords, max_), min_)


def clip_boxes(boxes, im_shape):
    """In place, like the reference: x columns to [0, width-1], then y columns to [0, height-1]."""
    height, width = im_shape[0], im_shape[1]
    for first_col, limit in ((0, width - 1), (1, height - 1)):
        boxes[:, first_col::2] = threshold(boxes[:, first_col::2], 0, limit)
    return boxes


class Graph:
    def __init__(self, graph):
        self.graph = graph

    def sub_graphs_connected(self):
        g = self.graph
        if g.shape[0] == 0:
            return []
        has_in = g.any(axis=0)
        has_out = g.any(axis=1)
        first_out = g.argmax(axis=1)          # index of the first out-edge (other.py:27)
        sub_graphs = []
        for index in np.where(~has_in & has_out)[0]:
            v = int(index)
            chain = [v]
            while has_out[v]:
                v = int(first_out[v])
                chain.append(v)
            sub_graphs.append(chain)
        return sub_graphs
