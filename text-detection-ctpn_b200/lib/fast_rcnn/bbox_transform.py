"""bbox_transform(ex_rois, gt_rois): the regression targets of lib/fast_rcnn/bbox_transform.py:3-34 (dx, dy relative to
the example box size; dw, dh as log ratios), [n,4].  Host numpy, used by training-side callers; anchor_target_layer gets
the same quantity from ctpn_anchor_targets_host.  The inference-side inverse and clipping (:36-80) run on the device
inside ctpn_proposals and have no host copy."""
import numpy as np


def bbox_transform(ex_rois, gt_rois):
    ex_size = ex_rois[:, 2:4] - ex_rois[:, 0:2] + 1.0
    ex_ctr = ex_rois[:, 0:2] + 0.5 * ex_size
    if not (np.min(ex_size[:, 0]) > 0.1 and np.min(ex_size[:, 1]) > 0.1):
        raise AssertionError("Invalid boxes found: %s %s" % (ex_rois[np.argmin(ex_size[:, 0]), :], ex_rois[np.argmin(ex_size[:, 1]), :]))
    gt_size = gt_rois[:, 2:4] - gt_rois[:, 0:2] + 1.0
    gt_ctr = gt_rois[:, 0:2] + 0.5 * gt_size
    return np.hstack(((gt_ctr - ex_ctr) / ex_size, np.log(gt_size / ex_size)))
