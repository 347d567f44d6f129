"""Seeded synthetic weights / inputs shared by the golden generator, the tests,
smoke() and bench.py (test infrastructure; the product never imports this).

The reference's own initialisers (truncated_normal stddev 0.01, network.py:151,168)
collapse activations to ~0 after 14 layers, which makes every score 0.5 and turns NMS
order into pure tie-breaking; SURVEY.md App. A.6 therefore prescribes a
variance-preserving init.  Variable names/shapes follow SURVEY.md App. A.2.
"""
import numpy as np

from .net_cpu import CONV_LAYERS, LSTM_BW, LSTM_FW


def make_weights(seed=0):
    rs = np.random.RandomState(seed)
    w = {}
    for name, cin, cout, _ in CONV_LAYERS:
        std = np.sqrt(2.0 / (9 * cin))
        if name == "conv1_1":
            std /= 75.0     # mean-subtracted uint8 pixels have RMS ~75: bring activations to O(1)
        w[name + "/weights"] = (rs.standard_normal((3, 3, cin, cout)) * std).astype(np.float32)
        w[name + "/biases"] = (rs.standard_normal(cout) * 0.01).astype(np.float32)
    lim = np.sqrt(6.0 / (640 + 512))
    for scope in (LSTM_FW, LSTM_BW):
        w[scope + "/kernel"] = rs.uniform(-lim, lim, (640, 512)).astype(np.float32)
        w[scope + "/bias"] = (rs.standard_normal(512) * 0.01).astype(np.float32)
    w["lstm_o/weights"] = (rs.standard_normal((256, 512)) * np.sqrt(1.0 / 256)).astype(np.float32)
    w["lstm_o/biases"] = (rs.standard_normal(512) * 0.01).astype(np.float32)
    # head scales chosen so that (measured on 600x900 synthetic images) logits ~ N(0, 2)
    # and dy, dh ~ N(0, 0.3); see tests/test_oracle_net.py::test_head_statistics
    w["rpn_cls_score/weights"] = (rs.standard_normal((512, 20)) * HEAD_CLS_STD).astype(np.float32)
    w["rpn_cls_score/biases"] = (rs.standard_normal(20) * 0.01).astype(np.float32)
    w["rpn_bbox_pred/weights"] = (rs.standard_normal((512, 40)) * HEAD_BOX_STD).astype(np.float32)
    w["rpn_bbox_pred/biases"] = (rs.standard_normal(40) * 0.01).astype(np.float32)
    return w


HEAD_CLS_STD = 0.21
HEAD_BOX_STD = 0.026


def make_image(seed, h=600, w=900):
    """uint8 HWC BGR image, i.i.d. uniform pixels."""
    rs = np.random.RandomState(1000 + seed)
    return rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)


def make_head_outputs(seed, H, W, logit_std=2.0, delta_std=0.3, unique=True):
    """Direct synthetic head outputs for the proposal stage:
    cls_prob [1,H,W,20] (pair softmax of N(0,logit_std) logits) and bbox_pred
    [1,H,W,40] ~ N(0,delta_std).  With unique=True the fg scores are made pairwise
    distinct so the reference's unstable argsort has a single valid answer."""
    rs = np.random.RandomState(2000 + seed)
    logits = (rs.standard_normal((1, H, W, 10, 2)) * logit_std).astype(np.float32)
    m = logits.max(axis=-1, keepdims=True)
    e = np.exp(logits - m)
    prob = (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)
    if unique:
        fg = prob[..., 1].ravel()
        for _ in range(50):
            _, first = np.unique(fg, return_index=True)
            dup = np.setdiff1d(np.arange(fg.size), first)
            if dup.size == 0:
                break
            fg[dup] = np.nextafter(fg[dup], np.float32(0.0)).astype(np.float32)
        prob[..., 1] = fg.reshape(prob[..., 1].shape)
        prob[..., 0] = np.float32(1.0) - prob[..., 1]
    bbox = (rs.standard_normal((1, H, W, 40)) * delta_std).astype(np.float32)
    return prob.reshape(1, H, W, 20), bbox


def make_boxes(seed, n, im_w=900, im_h=600, ctpn_like=False):
    """Random dets [n,5] = [x1,y1,x2,y2,score] with distinct scores, for nms()."""
    rs = np.random.RandomState(3000 + seed)
    if ctpn_like:
        x1 = (rs.randint(0, im_w // 16, n) * 16).astype(np.float32)
        x2 = x1 + 16
    else:
        x1 = rs.uniform(0, im_w - 40, n).astype(np.float32)
        x2 = x1 + rs.uniform(4, 120, n).astype(np.float32)
    y1 = rs.uniform(0, im_h - 40, n).astype(np.float32)
    y2 = y1 + rs.uniform(4, 150, n).astype(np.float32)
    sc = rs.permutation(n).astype(np.float32) / np.float32(n)
    sc = (sc * np.float32(0.999) + np.float32(0.0005)).astype(np.float32)
    return np.stack([x1, y1, np.minimum(x2, im_w - 1), np.minimum(y2, im_h - 1), sc], 1).astype(np.float32)


def make_text_proposals(seed, im_h=600, im_w=900, n_lines=6, n_noise=120):
    """Proposals shaped like test_ctpn() output for an image with a few text
    lines: 16-px wide strips along slowly drifting baselines, plus clutter."""
    rs = np.random.RandomState(4000 + seed)
    boxes, scores = [], []
    for _ in range(n_lines):
        c0 = rs.randint(0, im_w // 16 - 8)
        c1 = rs.randint(c0 + 3, min(c0 + 40, im_w // 16))
        yc = rs.uniform(40, im_h - 40)
        hh = rs.uniform(12, 45)
        slope = rs.uniform(-0.08, 0.08)
        base = rs.uniform(0.86, 0.965)
        for c in range(c0, c1):
            if rs.rand() < 0.08:
                continue
            y = yc + slope * 16 * (c - c0) + rs.normal(0, 1.0)
            h = hh * (1 + rs.normal(0, 0.04))
            boxes.append([16 * c, y - h / 2, 16 * c + 16, y + h / 2])
            scores.append(base + rs.uniform(0.0, 0.03))
            if rs.rand() < 0.5:      # overlapping duplicate for the 0.2 NMS to remove
                boxes.append([16 * c, y - h / 2 + rs.normal(0, 1.5), 16 * c + 16, y + h / 2 + rs.normal(0, 1.5)])
                scores.append(rs.uniform(0.71, 0.95))
    for _ in range(n_noise):
        c = rs.randint(0, im_w // 16)
        y = rs.uniform(10, im_h - 10)
        h = rs.uniform(8, 120)
        boxes.append([16 * c, y - h / 2, 16 * c + 16, y + h / 2])
        scores.append(rs.uniform(0.3, 0.98))
    b = np.asarray(boxes, np.float32)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, im_w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, im_h - 1)
    s = np.asarray(scores, np.float32)
    # distinct scores (unstable-sort safety), descending as test_ctpn returns them
    s = s + (np.arange(s.size, dtype=np.float32) * np.float32(1e-6))
    order = np.argsort(-s, kind="stable")
    return b[order], s[order][:, None].astype(np.float32)


def make_gt_boxes(seed, im_h=600, im_w=900, n_lines=6, scale=1.0, n_dontcare=0, hard_frac=0.0, n_outside=0):
    """Training annotations shaped like the reference's split text lines (prepare_training_data/split_label.py): 16-px
    wide ground-truth strips along a few lines, multiplied by the image scale (so coordinates are fractional), class 1.
    Returns gt_boxes [G,5] float32, gt_ishard [G] int32, dontcare_areas [D,4] float32."""
    rs = np.random.RandomState(6000 + seed)
    gt = []
    for _ in range(n_lines):
        c0 = rs.randint(0, max(1, int(im_w / scale) // 16 - 6))
        c1 = rs.randint(c0 + 2, min(c0 + 30, int(im_w / scale) // 16) + 1)
        yc = rs.uniform(30, im_h / scale - 30)
        hh = rs.uniform(10, 60)
        slope = rs.uniform(-0.05, 0.05)
        for c in range(c0, c1):
            y = yc + slope * 16 * (c - c0)
            gt.append([16 * c, max(0.0, y - hh / 2), 16 * c + 15, min(im_h / scale - 1, y + hh / 2), 1.0])
    for _ in range(n_outside):          # annotations that fall off the resized image: no inside anchor touches them
        gt.append([im_w / scale + 40, 10, im_w / scale + 55, 40, 1.0])
    gt = np.asarray(gt, np.float32)
    gt[:, :4] *= np.float32(scale)
    hard = (rs.rand(gt.shape[0]) < hard_frac).astype(np.int32)
    dc = np.zeros((n_dontcare, 4), np.float32)
    for i in range(n_dontcare):
        x, y = rs.uniform(0, im_w - 120), rs.uniform(0, im_h - 80)
        dc[i] = [x, y, x + rs.uniform(30, 110), y + rs.uniform(20, 70)]
    return gt, hard, dc
