"""TextProposalConnector, horizontal lines (lib/text_connector/text_proposal_connector.py:21-64): a delegate to the
library's host connector -- grouping in C++ (ctpn_text_groups_host), fitting with numpy's own np.polyfit so the
coordinates are the reference's bit for bit (ctpn_b200/textlines.py)."""
from ctpn_b200 import textlines

from .text_connect_cfg import native_cfg


def fit_y(X, Y, x1, x2):
    return textlines._edge_at(X, Y, x1, x2)


class TextProposalConnector:
    MODE = "H"

    def group_text_proposals(self, text_proposals, scores, im_size):
        return textlines.groups(text_proposals, scores, im_size, native_cfg())

    def fit_y(self, X, Y, x1, x2):
        return fit_y(X, Y, x1, x2)

    def get_text_lines(self, text_proposals, scores, im_size):
        chains = self.group_text_proposals(text_proposals, scores, im_size)
        return textlines.fit_lines(text_proposals, scores, chains, im_size, self.MODE)
