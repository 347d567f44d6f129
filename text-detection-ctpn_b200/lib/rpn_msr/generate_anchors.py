"""generate_anchors() -- the 10 fixed-width CTPN anchors of lib/rpn_msr/generate_anchors.py:24-32
(heights 11..283, width 16), as an int32 [10,4] table.  The device kernels carry the same table
(csrc/proposal.cu); this host copy exists for API parity.  `scales`/`ratios` are ignored exactly
as in the reference."""
import numpy as np

_HEIGHTS = [11, 16, 23, 33, 48, 68, 97, 139, 198, 283]


def generate_anchors(base_size=16, ratios=[0.5, 1, 2], scales=2 ** np.arange(3, 6), py2=False):
    ctr = (base_size - 1) * 0.5
    out = np.zeros((len(_HEIGHTS), 4), np.int32)
    for i, h in enumerate(_HEIGHTS):
        hw, hh = (8, h // 2) if py2 else (8.0, h / 2)
        out[i] = [int(ctr - hw), int(ctr - hh), int(ctr + hw), int(ctr + hh)]
    return out
