"""TextDetector (lib/text_connector/detectors.py:19-49): score filter > 0.7, sort, NMS 0.2
(on the GPU through nms()), text-line construction (host side, as in the reference) and the
final line filter.  TextDetector(native=True) runs all of detect() in the library's C++ connector instead
(ctpn_text_lines_host: same line sets, coordinates equal to float32 rounding, ~20x faster)."""
import numpy as np

from lib.fast_rcnn.nms_wrapper import nms
from lib.fast_rcnn.config import cfg
from .text_proposal_connector import TextProposalConnector
from .text_proposal_connector_oriented import TextProposalConnector as TextProposalConnectorOriented
from .text_connect_cfg import Config as TextLineCfg


class TextDetector:
    def __init__(self, native=False):
        self.native = bool(native)
        self.mode = cfg.TEST.DETECT_MODE
        if self.mode == "H":
            self.text_proposal_connector = TextProposalConnector()
        elif self.mode == "O":
            self.text_proposal_connector = TextProposalConnectorOriented()

    def detect(self, text_proposals, scores, size):
        if self.native:
            from ctpn_b200.textlines import text_lines
            c = TextLineCfg
            return text_lines(text_proposals, scores, size, self.mode,
                              (c.TEXT_PROPOSALS_MIN_SCORE, c.TEXT_PROPOSALS_NMS_THRESH, c.MAX_HORIZONTAL_GAP, c.MIN_V_OVERLAPS,
                               c.MIN_SIZE_SIM, c.MIN_RATIO, c.LINE_MIN_SCORE, c.TEXT_PROPOSALS_WIDTH, c.MIN_NUM_PROPOSALS))
        keep_inds = np.where(scores > TextLineCfg.TEXT_PROPOSALS_MIN_SCORE)[0]
        text_proposals, scores = text_proposals[keep_inds], scores[keep_inds]
        # score descending, index ascending on ties (the reference's argsort()[::-1] is unstable)
        sorted_indices = np.argsort(-scores.ravel(), kind="stable")
        text_proposals, scores = text_proposals[sorted_indices], scores[sorted_indices]
        keep_inds = nms(np.hstack((text_proposals, scores)), TextLineCfg.TEXT_PROPOSALS_NMS_THRESH)
        text_proposals, scores = text_proposals[keep_inds], scores[keep_inds]
        text_recs = self.text_proposal_connector.get_text_lines(text_proposals, scores, size)
        keep_inds = self.filter_boxes(text_recs)
        return text_recs[keep_inds]

    def filter_boxes(self, boxes):
        boxes = np.asarray(boxes, np.float64).reshape(-1, 9)
        heights = (np.abs(boxes[:, 5] - boxes[:, 1]) + np.abs(boxes[:, 7] - boxes[:, 3])) / 2.0 + 1
        widths = (np.abs(boxes[:, 2] - boxes[:, 0]) + np.abs(boxes[:, 6] - boxes[:, 4])) / 2.0 + 1
        scores = boxes[:, 8]
        return np.where((widths / heights > TextLineCfg.MIN_RATIO) & (scores > TextLineCfg.LINE_MIN_SCORE) &
                        (widths > (TextLineCfg.TEXT_PROPOSALS_WIDTH * TextLineCfg.MIN_NUM_PROPOSALS)))[0]
