"""Oracle (test infrastructure): CPU restatement of the reference's TF1 test
graph (VGG16 -> rpn_conv -> BiLSTM -> FC -> heads -> pair softmax).
Not imported by the product.  PARITY UNPINNED: TensorFlow 1.3 is not available
(requirements.txt:2), so this restates TF 1.3's documented op semantics; it is
checked against a float64 evaluation of itself and against oracle/net_alt.py, an
independently written second restatement (tests/test_oracle_net_cpu.py).  The graph wiring (layer
order, variable names, row sequences, reshapes, proposal-layer call) is pinned to the reference's own graph-building code
run on a numpy TensorFlow stand-in (tests/golden/make_golden_net.py, reference_net_wiring.npz); the per-op semantics are not.

Follows (paths relative to /root/reference):
  topology            lib/networks/VGGnet_test.py:16-55
  conv (+bias, ReLU)  lib/networks/network.py:160-183   3x3, stride 1, SAME, HWIO weights
  max_pool            lib/networks/network.py:189-196   2x2, stride 2, VALID
  Bilstm              lib/networks/network.py:88-113    LSTMCell(128) fw/bw over W, concat, 256->512
  lstm_fc             lib/networks/network.py:144-158   512->40, 512->20
  pair softmax        lib/networks/network.py:269-277, 332-337
  image blob          lib/fast_rcnn/test.py:7-31, lib/utils/blob.py:6-19, config.py:200

TF 1.3 LSTMCell (no peepholes/projection): gates = [x, h] . kernel[640,512] + bias,
split order i, j, f, o; c = sigmoid(f + 1.0) * c + sigmoid(i) * tanh(j);
h = sigmoid(o) * tanh(c); zero initial state.
"""
import numpy as np
import torch
import torch.nn.functional as F

CONV_LAYERS = [  # (name, cin, cout, pool_after)
    ("conv1_1", 3, 64, False), ("conv1_2", 64, 64, True),
    ("conv2_1", 64, 128, False), ("conv2_2", 128, 128, True),
    ("conv3_1", 128, 256, False), ("conv3_2", 256, 256, False), ("conv3_3", 256, 256, True),
    ("conv4_1", 256, 512, False), ("conv4_2", 512, 512, False), ("conv4_3", 512, 512, True),
    ("conv5_1", 512, 512, False), ("conv5_2", 512, 512, False), ("conv5_3", 512, 512, False),
    ("rpn_conv/3x3", 512, 512, False),
]
PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])  # config.py:200 (float64, BGR)
LSTM_FW = "lstm_o/bidirectional_rnn/fw/lstm_cell"
LSTM_BW = "lstm_o/bidirectional_rnn/bw/lstm_cell"


def image_blob(im):
    """test.py:7-31 for the scale==1 case + the generic cv2 path.  Returns
    (blob [1,H,W,3] f32, im_scale).  ``im -= PIXEL_MEANS`` is an in-place float32
    op with a float64 operand: computed in float64, rounded to float32."""
    import cv2
    im_orig = im.astype(np.float32, copy=True)
    im_orig -= PIXEL_MEANS
    size_min = min(im_orig.shape[0:2])
    size_max = max(im_orig.shape[0:2])
    scale = float(600) / float(size_min)               # cfg.TEST.SCALES = (600,)
    if np.round(scale * size_max) > 1000:              # cfg.TEST.MAX_SIZE
        scale = float(1000) / float(size_max)
    out = cv2.resize(im_orig, None, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
    return out[None].astype(np.float32), scale


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def lstm_dir(x, kernel, bias, reverse):
    """x [R, W, 512]; kernel [640,512]; returns h for every step, [R, W, 128]."""
    R, W, _ = x.shape
    wx, wh = kernel[:512], kernel[512:]
    xp = x @ wx + bias                                   # [R, W, 512]
    h = x.new_zeros(R, 128)
    c = x.new_zeros(R, 128)
    out = x.new_zeros(R, W, 128)
    steps = range(W - 1, -1, -1) if reverse else range(W)
    for t in steps:
        g = xp[:, t] + h @ wh
        i, j, f, o = g[:, :128], g[:, 128:256], g[:, 256:384], g[:, 384:]
        c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(j)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def forward(blob, weights, dtype=torch.float32, taps=None):
    """blob [N,H,W,3] float; weights: dict name -> ndarray (TF variable names,
    App. A.2 of SURVEY.md).  Returns dict with 'rpn_cls_prob_reshape' [N,H',W',20],
    'rpn_bbox_pred' [N,H',W',40] (numpy, NHWC) plus any intermediate named in taps."""
    taps = set(taps or [])
    res = {}
    with torch.no_grad():
        x = _t(blob, dtype).permute(0, 3, 1, 2)           # NCHW for torch
        for name, cin, cout, pool in CONV_LAYERS:
            w = _t(weights[name + "/weights"], dtype).permute(3, 2, 0, 1)   # HWIO -> OIHW
            b = _t(weights[name + "/biases"], dtype)
            x = F.relu(F.conv2d(x, w, b, stride=1, padding=1))
            if name in taps:
                res[name] = x.permute(0, 2, 3, 1).numpy().copy()
            if pool:
                x = F.max_pool2d(x, 2, 2)                  # VALID: floor
                if name + "+pool" in taps:
                    res[name + "+pool"] = x.permute(0, 2, 3, 1).numpy().copy()
        N, C, H, W = x.shape
        seq = x.permute(0, 2, 3, 1).reshape(N * H, W, C)
        fw = lstm_dir(seq, _t(weights[LSTM_FW + "/kernel"], dtype), _t(weights[LSTM_FW + "/bias"], dtype), False)
        bw = lstm_dir(seq, _t(weights[LSTM_BW + "/kernel"], dtype), _t(weights[LSTM_BW + "/bias"], dtype), True)
        lstm_out = torch.cat([fw, bw], dim=-1).reshape(N * H * W, 256)
        if "lstm_out" in taps:
            res["lstm_out"] = lstm_out.reshape(N, H, W, 256).numpy().copy()
        fc = lstm_out @ _t(weights["lstm_o/weights"], dtype) + _t(weights["lstm_o/biases"], dtype)
        if "lstm_o" in taps:
            res["lstm_o"] = fc.reshape(N, H, W, 512).numpy().copy()
        bbox = fc @ _t(weights["rpn_bbox_pred/weights"], dtype) + _t(weights["rpn_bbox_pred/biases"], dtype)
        score = fc @ _t(weights["rpn_cls_score/weights"], dtype) + _t(weights["rpn_cls_score/biases"], dtype)
        prob = torch.softmax(score.reshape(-1, 2), dim=-1).reshape(N, H, W, 20)
        res["rpn_cls_score"] = score.reshape(N, H, W, 20).numpy().copy()
        res["rpn_cls_prob_reshape"] = prob.numpy().copy()
        res["rpn_bbox_pred"] = bbox.reshape(N, H, W, 40).numpy().copy()
    return res
