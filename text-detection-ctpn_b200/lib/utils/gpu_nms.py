"""gpu_nms(dets, thresh, device_id=0): same signature and result as the reference's Cython
module lib/utils/gpu_nms.pyx:18-33, backed by ctpn_nms_host (include/ctpn_b200.h), the
C-ABI replacement of `_nms` (lib/utils/gpu_nms.hpp:1-2).

Ties: the reference orders with `scores.argsort()[::-1]` (unstable); here equal scores are
visited in ascending index order (stable), the repo-wide canonical rule."""
import ctypes as C

import numpy as np

from ctpn_b200 import _native as N

assert C.sizeof(C.c_int) == 4   # gpu_nms.pyx:13


def gpu_nms(dets, thresh, device_id=0):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2 or dets.shape[1] < 5:
        raise ValueError("dets must be [N, >=5] float32 (x1, y1, x2, y2, score)")
    boxes_num, boxes_dim = dets.shape
    keep = np.zeros(boxes_num, dtype=np.int32)
    num_out = C.c_int(0)
    order = np.argsort(-dets[:, 4], kind="stable")
    sorted_dets = np.ascontiguousarray(dets[order, :])
    N.check(N.lib.ctpn_nms_host(N.ptr(keep), C.byref(num_out), N.ptr(sorted_dets), boxes_num, boxes_dim,
                                np.float32(thresh), int(device_id)), "ctpn_nms_host")
    keep = keep[:num_out.value]
    return list(order[keep])
