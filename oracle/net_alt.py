"""Oracle cross-check (test infrastructure): a SECOND, independently written restatement of the reference's test graph,
used only to narrow the "parity unpinned" gap of oracle/net_cpu.py (TensorFlow 1.3 cannot run here): the two restatements
share no code and use different primitives, so a slip in either one's reading of the graph shows up as a mismatch
(tests/test_oracle_net_cpu.py asserts agreement to ~1e-12 in float64).

  conv       lib/networks/network.py:160-183  -> explicit im2col (zero-padded gather) + one matmul per layer, numpy
  max_pool   lib/networks/network.py:189-196  -> reshape to 2x2 blocks after cropping odd rows/columns (VALID), numpy
  Bilstm     lib/networks/network.py:88-113   -> torch.nn.LSTM(bidirectional): torch's gate order is (i, f, g, o) and it has
                                                 no forget bias, TF 1.3's LSTMCell is (i, j, f, o) with forget_bias 1.0 added
                                                 inside the sigmoid, so TF columns are permuted and +1 goes into the f bias
  lstm_fc    lib/networks/network.py:144-158  -> numpy matmul
  softmax    lib/networks/network.py:269-277, 332-337 -> sigmoid of the pair difference (softmax over two logits)
Not imported by the product."""
import numpy as np
import torch

LAYERS = ["conv1_1", "conv1_2", "P", "conv2_1", "conv2_2", "P", "conv3_1", "conv3_2", "conv3_3", "P",
          "conv4_1", "conv4_2", "conv4_3", "P", "conv5_1", "conv5_2", "conv5_3", "rpn_conv/3x3"]


def conv3x3_same(x, w, b):
    """x [N,H,W,Cin], w [3,3,Cin,Cout] (HWIO), b [Cout]: SAME padding, stride 1, then ReLU."""
    n, h, wd, cin = x.shape
    xp = np.zeros((n, h + 2, wd + 2, cin), x.dtype)
    xp[:, 1:-1, 1:-1] = x
    cols = np.concatenate([xp[:, dy:dy + h, dx:dx + wd] for dy in range(3) for dx in range(3)], axis=-1)   # [N,H,W,9*Cin], (dy,dx,c)
    y = cols.reshape(-1, 9 * cin) @ w.reshape(9 * cin, -1) + b
    return np.maximum(y, 0).reshape(n, h, wd, -1)


def pool2x2_valid(x):
    n, h, w, c = x.shape
    x = x[:, :h - h % 2, :w - w % 2]
    return x.reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


def _torch_lstm(params, prefix):
    """TF kernel [640,512] / bias [512], columns (i, j, f, o) -> torch weight_ih [4*128, 512], weight_hh [4*128, 128], rows (i, f, g, o)."""
    k, b = params[prefix + "/kernel"].astype(np.float64), params[prefix + "/bias"].astype(np.float64)
    i, j, f, o = (slice(q * 128, (q + 1) * 128) for q in range(4))
    order = [i, f, j, o]
    w_ih = np.concatenate([k[:512, s].T for s in order])
    w_hh = np.concatenate([k[512:, s].T for s in order])
    bias = np.concatenate([b[i], b[f] + 1.0, b[j], b[o]])          # forget_bias = 1.0
    return w_ih, w_hh, bias


def forward(blob, params):
    """blob [N,H,W,3] (mean-subtracted); float64 throughout.  Returns rpn_cls_score, rpn_cls_prob_reshape, rpn_bbox_pred."""
    x = np.asarray(blob, np.float64)
    for name in LAYERS:
        if name == "P":
            x = pool2x2_valid(x)
        else:
            x = conv3x3_same(x, params[name + "/weights"].astype(np.float64), params[name + "/biases"].astype(np.float64))
    n, h, w, c = x.shape
    lstm = torch.nn.LSTM(512, 128, batch_first=True, bidirectional=True).double()
    fw, bw = (_torch_lstm(params, "lstm_o/bidirectional_rnn/%s/lstm_cell" % d) for d in ("fw", "bw"))
    with torch.no_grad():
        for sfx, (w_ih, w_hh, bias) in (("", fw), ("_reverse", bw)):
            getattr(lstm, "weight_ih_l0" + sfx).copy_(torch.from_numpy(w_ih))
            getattr(lstm, "weight_hh_l0" + sfx).copy_(torch.from_numpy(w_hh))
            getattr(lstm, "bias_ih_l0" + sfx).copy_(torch.from_numpy(bias))
            getattr(lstm, "bias_hh_l0" + sfx).zero_()
        seq, _ = lstm(torch.from_numpy(x.reshape(n * h, w, c)))     # rows of the feature map are the sequences (network.py:91-93)
    feat = seq.numpy().reshape(n * h * w, 256)
    fc = feat @ params["lstm_o/weights"].astype(np.float64) + params["lstm_o/biases"].astype(np.float64)
    bbox = fc @ params["rpn_bbox_pred/weights"].astype(np.float64) + params["rpn_bbox_pred/biases"].astype(np.float64)
    score = fc @ params["rpn_cls_score/weights"].astype(np.float64) + params["rpn_cls_score/biases"].astype(np.float64)
    pairs = score.reshape(-1, 2)                                    # channel 2a = bg, 2a+1 = fg (VGGnet_test.py:46-52)
    fg = 1.0 / (1.0 + np.exp(pairs[:, 0] - pairs[:, 1]))
    prob = np.stack([1.0 - fg, fg], axis=1).reshape(n, h, w, 20)
    return {"rpn_cls_score": score.reshape(n, h, w, 20), "rpn_cls_prob_reshape": prob, "rpn_bbox_pred": bbox.reshape(n, h, w, 40)}
