"""bbox_overlaps(boxes, query_boxes) / bbox_intersections(boxes, query_boxes): same signatures, dtype rules and results as
the reference's Cython module lib/utils/bbox.pyx:15-55, :57-95 (float64 [N,>=4] x [K,>=4] -> [N,K]), backed by
ctpn_bbox_overlaps_host / ctpn_bbox_intersections_host (include/ctpn_b200.h).  Host code: the reference's is too."""
import numpy as np

from ctpn_b200 import _native as N


def _pairwise(fn, name, boxes, query_boxes):
    # the Cython signature is np.ndarray[np.float_t, ndim=2]: anything else is a ValueError there as well
    for arr, what in ((boxes, "boxes"), (query_boxes, "query_boxes")):
        if not isinstance(arr, np.ndarray) or arr.dtype != np.float64 or arr.ndim != 2:
            raise ValueError("%s: %s must be a 2-d float64 ndarray" % (name, what))
        if arr.shape[0] > 0 and arr.shape[1] < 4:
            raise IndexError("%s: %s needs at least 4 columns" % (name, what))
    boxes = np.ascontiguousarray(boxes)
    query_boxes = np.ascontiguousarray(query_boxes)
    n, k = boxes.shape[0], query_boxes.shape[0]
    out = np.zeros((n, k), np.float64)
    if n and k:
        N.check(fn(N.ptr(boxes), n, boxes.shape[1], N.ptr(query_boxes), k, query_boxes.shape[1], N.ptr(out)), name)
    return out


def bbox_overlaps(boxes, query_boxes):
    return _pairwise(N.lib.ctpn_bbox_overlaps_host, "bbox_overlaps", boxes, query_boxes)


def bbox_intersections(boxes, query_boxes):
    return _pairwise(N.lib.ctpn_bbox_intersections_host, "bbox_intersections", boxes, query_boxes)
