"""Oracle (test infrastructure): numpy restatement of the reference's RPN
post-processing.  Not imported by the product.

Follows, function by function (paths relative to /root/reference):
  anchors()              lib/rpn_msr/generate_anchors.py:3-32
  decode_boxes()         lib/fast_rcnn/bbox_transform.py:36-65  (CTPN variant: dx, dw ignored)
  clip()                 lib/fast_rcnn/bbox_transform.py:67-80
  size_filter()          lib/rpn_msr/proposal_layer_tf.py:160-165
  nms()                  lib/fast_rcnn/nms_wrapper.py:23-47 (py_cpu_nms) ==
                         lib/utils/nms_kernel.cu:24-32,124-139 (same IoU, strict '>')
  proposal_layer()       lib/rpn_msr/proposal_layer_tf.py:14-157

Two deliberate, documented choices where the reference is under-specified:
  * Tie order.  The reference sorts with ``argsort()[::-1]`` (unstable); equal
    scores come out in an unspecified order.  The oracle's canonical rule is
    "score descending, original index ascending" (``order_desc``).
  * exp().  ``exp_mode='numpy'`` uses numpy's float32 exp exactly like the
    reference; ``exp_mode='rounded'`` (canonical for CUDA parity) evaluates
    exp in float64 and rounds once to float32, which the CUDA kernel
    reproduces bit-for-bit.  The two differ by at most 1 ulp(float32).
"""
import numpy as np

F32 = np.float32

# lib/fast_rcnn/config.py:175-183 (TEST block) and ctpn/text.yml
DEFAULTS = dict(pre_nms_topN=12000, post_nms_topN=1000, nms_thresh=0.7, min_size=8, feat_stride=16)

ANCHOR_HEIGHTS = (11, 16, 23, 33, 48, 68, 97, 139, 198, 283)  # generate_anchors.py:26


def anchors(py2=False):
    """int32 [10,4] base anchors.  generate_anchors.py:13-21 writes float
    expressions into an int32 array (truncation toward zero); ``h / 2`` is true
    division on py3 (canonical here) and floor division on py2."""
    out = np.zeros((len(ANCHOR_HEIGHTS), 4), np.int32)
    x_ctr = (0 + 15) * 0.5
    y_ctr = (0 + 15) * 0.5
    w = 16
    for i, h in enumerate(ANCHOR_HEIGHTS):
        hw = (w // 2) if py2 else (w / 2)
        hh = (h // 2) if py2 else (h / 2)
        vals = (x_ctr - hw, y_ctr - hh, x_ctr + hw, y_ctr + hh)
        out[i] = [int(v) for v in vals]  # C-style truncation, as numpy int32 assignment does
    return out


def shifted_anchors(height, width, feat_stride=16, py2=False):
    """[H*W*10, 4] int anchors, rows ordered (h, w, a).  proposal_layer_tf.py:83-99"""
    base = anchors(py2).astype(np.int64)
    sx = np.arange(width) * feat_stride
    sy = np.arange(height) * feat_stride
    gx, gy = np.meshgrid(sx, sy)
    shifts = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)
    return (base[None, :, :] + shifts[:, None, :]).reshape(-1, 4)


def exp_f32(x, mode="rounded"):
    x = np.asarray(x, F32)
    if mode == "numpy":
        return np.exp(x)
    return np.exp(x.astype(np.float64)).astype(F32)


def decode_boxes(anch, deltas, exp_mode="rounded"):
    """bbox_transform.py:36-65.  All arithmetic in float32 (deltas' dtype)."""
    b = anch.astype(F32)
    d = deltas.astype(F32, copy=False)
    widths = b[:, 2] - b[:, 0] + F32(1.0)
    heights = b[:, 3] - b[:, 1] + F32(1.0)
    ctr_x = b[:, 0] + F32(0.5) * widths
    ctr_y = b[:, 1] + F32(0.5) * heights
    dy = d[:, 1]
    dh = d[:, 3]
    pred_ctr_x = ctr_x
    pred_ctr_y = dy * heights + ctr_y
    pred_w = widths
    pred_h = exp_f32(dh, exp_mode) * heights
    out = np.zeros(d.shape, F32)
    out[:, 0] = pred_ctr_x - F32(0.5) * pred_w
    out[:, 1] = pred_ctr_y - F32(0.5) * pred_h
    out[:, 2] = pred_ctr_x + F32(0.5) * pred_w
    out[:, 3] = pred_ctr_y + F32(0.5) * pred_h
    return out


def clip(boxes, im_h, im_w):
    """bbox_transform.py:67-80: max(min(v, dim-1), 0), float32."""
    hx = F32(F32(im_w) - F32(1))
    hy = F32(F32(im_h) - F32(1))
    out = boxes.copy()
    out[:, 0] = np.maximum(np.minimum(out[:, 0], hx), F32(0))
    out[:, 1] = np.maximum(np.minimum(out[:, 1], hy), F32(0))
    out[:, 2] = np.maximum(np.minimum(out[:, 2], hx), F32(0))
    out[:, 3] = np.maximum(np.minimum(out[:, 3], hy), F32(0))
    return out


def size_filter(boxes, min_size):
    """proposal_layer_tf.py:160-165; min_size is float32 (8 * im_info[2])."""
    ms = F32(min_size)
    ws = boxes[:, 2] - boxes[:, 0] + F32(1)
    hs = boxes[:, 3] - boxes[:, 1] + F32(1)
    return np.where((ws >= ms) & (hs >= ms))[0]


def order_desc(scores):
    """Canonical visiting order: score descending, index ascending on ties."""
    s = np.asarray(scores, F32).ravel()
    return np.argsort(-s, kind="stable")


def iou_row(box, others):
    """nms_kernel.cu:24-32 / nms_wrapper.py:30,37-44 -- float32, +1 pixel convention."""
    one = F32(1)
    xx1 = np.maximum(box[0], others[:, 0])
    yy1 = np.maximum(box[1], others[:, 1])
    xx2 = np.minimum(box[2], others[:, 2])
    yy2 = np.minimum(box[3], others[:, 3])
    w = np.maximum(F32(0), xx2 - xx1 + one)
    h = np.maximum(F32(0), yy2 - yy1 + one)
    inter = w * h
    area_b = (box[2] - box[0] + one) * (box[3] - box[1] + one)
    area_o = (others[:, 2] - others[:, 0] + one) * (others[:, 3] - others[:, 1] + one)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (area_b + area_o - inter)


def nms_sorted(boxes, thresh, max_keep=0):
    """Greedy NMS over boxes ALREADY in visiting order; returns kept positions.
    Suppress iff IoU > float32(thresh) (nms_kernel.cu:71; nms_wrapper.py:45 keeps '<=')."""
    b = np.ascontiguousarray(boxes[:, :4], F32)
    t = F32(thresh)
    n = b.shape[0]
    alive = np.ones(n, bool)
    keep = []
    for i in range(n):
        if not alive[i]:
            continue
        keep.append(i)
        if max_keep and len(keep) >= max_keep:
            break
        if i + 1 < n:
            ovr = iou_row(b[i], b[i + 1:])
            alive[i + 1:] &= ~(ovr > t)
    return np.asarray(keep, np.int64)


def nms(dets, thresh):
    """nms(dets[N,5]=[x1,y1,x2,y2,score], thresh) -> indices into dets in visiting
    order.  nms_wrapper.py:11-20; [] for empty input."""
    dets = np.asarray(dets, F32)
    if dets.shape[0] == 0:
        return []
    order = order_desc(dets[:, 4])
    kept = nms_sorted(dets[order], thresh)
    return [int(v) for v in order[kept]]


def proposal_layer(cls_prob, bbox_pred, im_info, pre_nms_topN=12000, post_nms_topN=1000,
                   nms_thresh=0.7, min_size=8, feat_stride=16, exp_mode="rounded", py2=False,
                   return_index=False):
    """proposal_layer_tf.py:14-157.  cls_prob [1,H,W,20], bbox_pred [1,H,W,40],
    im_info [1,3] = (blob_h, blob_w, scale).  Returns (blob [n,5] f32 =
    [score,x1,y1,x2,y2], deltas [n,4]) and optionally the flat (h,w,a) anchor index
    of every output row."""
    cls_prob = np.asarray(cls_prob, F32)
    bbox_pred = np.asarray(bbox_pred, F32)
    info = np.asarray(im_info, F32).reshape(-1, 3)[0]
    assert cls_prob.shape[0] == 1, "Only single item batches are supported"
    H, W = cls_prob.shape[1:3]
    A = 10
    scores = cls_prob.reshape(1, H, W, A, 2)[..., 1].reshape(-1)          # fg prob, (h,w,a) order
    deltas = bbox_pred.reshape(-1, 4)
    anch = shifted_anchors(H, W, feat_stride, py2)
    props = decode_boxes(anch, deltas, exp_mode)
    props = clip(props, info[0], info[1])
    keep = size_filter(props, F32(min_size) * info[2])
    idx = keep
    props, scores, deltas = props[keep], scores[keep], deltas[keep]
    order = order_desc(scores)
    if pre_nms_topN > 0:
        order = order[:pre_nms_topN]
    props, scores, deltas, idx = props[order], scores[order], deltas[order], idx[order]
    kept = nms_sorted(props, nms_thresh)
    if post_nms_topN > 0:
        kept = kept[:post_nms_topN]
    props, scores, deltas, idx = props[kept], scores[kept], deltas[kept], idx[kept]
    blob = np.concatenate([scores[:, None], props], axis=1).astype(F32)
    if return_index:
        return blob, deltas, idx
    return blob, deltas
