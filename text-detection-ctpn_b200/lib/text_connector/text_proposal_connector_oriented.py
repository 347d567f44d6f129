"""TextProposalConnector, oriented lines (lib/text_connector/text_proposal_connector_oriented.py:24-105): the same
delegate as the horizontal connector with the parallelogram fit (centre line, mean height + 2.5, no clipping)."""
from .text_proposal_connector import TextProposalConnector as _Horizontal


class TextProposalConnector(_Horizontal):
    MODE = "O"
