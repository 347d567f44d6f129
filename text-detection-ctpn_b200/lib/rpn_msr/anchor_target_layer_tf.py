"""anchor_target_layer(rpn_cls_score, gt_boxes, gt_ishard, dontcare_areas, im_info, _feat_stride, anchor_scales)
-> (rpn_labels [1,H,W,A], rpn_bbox_targets, rpn_bbox_inside_weights, rpn_bbox_outside_weights [1,H,W,4A]), float32:
the operator interface of lib/rpn_msr/anchor_target_layer_tf.py:10 (called through tf.py_func at
lib/networks/network.py:225-243 in the training graph).  Training is outside this engine's scope; this
operator exists so that a training loop written against the reference finds it, and it is host code as the reference's is.

Overlaps, label assignment and regression targets come from one native pass (ctpn_anchor_targets_host; no
[anchors x ground truth] matrix).  The random sub-sampling of :181-198 stays here, on numpy's global RNG with the
reference's two npr.choice calls in the reference's order, so a seeded run draws the same anchors."""
import numpy as np
import numpy.random as npr

from ctpn_b200 import _native as N
from lib.fast_rcnn.config import cfg

from .generate_anchors import generate_anchors


def anchor_target_layer(rpn_cls_score, gt_boxes, gt_ishard, dontcare_areas, im_info, _feat_stride=[16, ], anchor_scales=[16, ]):
    A = generate_anchors(scales=np.array(anchor_scales)).shape[0]
    assert rpn_cls_score.shape[0] == 1, 'Only single item batches are supported'
    height, width = rpn_cls_score.shape[1:3]
    im_info = im_info[0]
    gt_boxes = np.asarray(gt_boxes)
    if gt_boxes.ndim != 2 or gt_boxes.shape[0] == 0 or gt_boxes.shape[1] != 5:
        raise ValueError("gt_boxes must be [G>=1, 5] (x1, y1, x2, y2, class); got %s" % (gt_boxes.shape,))
    # numpy keeps the ground-truth side of bbox_transform in the annotations' own precision when that is float32
    # (float16 annotations are widened to float32 here; the reference would carry on in half precision)
    if gt_boxes.dtype == np.float16:
        gt_boxes = gt_boxes.astype(np.float32)
    gt_is_f32 = gt_boxes.dtype == np.float32
    gt = np.ascontiguousarray(gt_boxes[:, :4], np.float64)
    T = cfg.TRAIN
    hard = None
    if T.PRECLUDE_HARD_SAMPLES and gt_ishard is not None and gt_ishard.shape[0] > 0:
        assert gt_ishard.shape[0] == gt_boxes.shape[0]
        hard = np.ascontiguousarray(np.asarray(gt_ishard).astype(int).reshape(-1) == 1, np.uint8)
    dontcare = None
    if dontcare_areas is not None and dontcare_areas.shape[0] > 0:
        dontcare = np.ascontiguousarray(np.asarray(dontcare_areas)[:, :4], np.float64)
    cfg5 = np.array([T.RPN_NEGATIVE_OVERLAP, T.RPN_POSITIVE_OVERLAP, float(bool(T.RPN_CLOBBER_POSITIVES)),
                     T.DONTCARE_AREA_INTERSECTION_HI, float(bool(T.PRECLUDE_HARD_SAMPLES))], np.float64)
    total = int(height * width * A)
    labels = np.empty((total,), np.float32)
    bbox_targets = np.empty((total, 4), np.float32)
    N.check(N.lib.ctpn_anchor_targets_host(N.ptr(gt), gt.shape[0], int(gt_is_f32), N.ptr(hard) if hard is not None else None,
                                           N.ptr(dontcare) if dontcare is not None else None,
                                           0 if dontcare is None else dontcare.shape[0], int(height), int(width),
                                           int(np.asarray(_feat_stride).reshape(-1)[0]), float(im_info[0]), float(im_info[1]),
                                           N.ptr(cfg5), N.ptr(labels), N.ptr(bbox_targets)), "ctpn_anchor_targets_host")

    # subsample positive, then negative labels if there are too many (:181-198); positions in the label vector are in the
    # same order as the reference's inside-anchor vector, so npr.choice picks the same anchors
    num_fg = int(T.RPN_FG_FRACTION * T.RPN_BATCHSIZE)
    fg_inds = np.where(labels == 1)[0]
    if len(fg_inds) > num_fg:
        labels[npr.choice(fg_inds, size=(len(fg_inds) - num_fg), replace=False)] = -1
    num_bg = T.RPN_BATCHSIZE - np.sum(labels == 1)
    bg_inds = np.where(labels == 0)[0]
    if len(bg_inds) > num_bg:
        labels[npr.choice(bg_inds, size=(len(bg_inds) - num_bg), replace=False)] = -1

    bbox_inside_weights = np.zeros((total, 4), np.float32)
    bbox_inside_weights[labels == 1, :] = np.array(T.RPN_BBOX_INSIDE_WEIGHTS)
    bbox_outside_weights = np.zeros((total, 4), np.float32)
    if T.RPN_POSITIVE_WEIGHT < 0:
        positive_weights, negative_weights = np.ones((1, 4)), np.zeros((1, 4))
    else:
        assert (T.RPN_POSITIVE_WEIGHT > 0) & (T.RPN_POSITIVE_WEIGHT < 1)
        # the reference's own expressions (:213-216): `w / count + 1`, not `w / (count + 1)`
        positive_weights = T.RPN_POSITIVE_WEIGHT / np.sum(labels == 1) + 1
        negative_weights = (1.0 - T.RPN_POSITIVE_WEIGHT) / np.sum(labels == 0) + 1
    bbox_outside_weights[labels == 1, :] = positive_weights
    bbox_outside_weights[labels == 0, :] = negative_weights

    return (labels.reshape((1, height, width, A)), bbox_targets.reshape((1, height, width, A * 4)),
            bbox_inside_weights.reshape((1, height, width, A * 4)), bbox_outside_weights.reshape((1, height, width, A * 4)))
