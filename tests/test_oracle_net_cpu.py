"""CPU: the network oracle (oracle/net_cpu.py, "parity unpinned": TF 1.3 cannot run here) against an independently
written second restatement (oracle/net_alt.py: im2col matmul convs, reshape pooling, torch.nn.LSTM with permuted gate
columns and the forget bias folded into the bias).  A shared misreading of the reference graph is still possible; a slip
in only one of them (gate order, forget bias, padding phase, pool flooring, tap order, scan direction) is not."""
import numpy as np
import torch

from oracle import net_alt, net_cpu, synth


def _blob(seed, h, w):
    return (synth.make_image(seed, h, w).astype(np.float32) - net_cpu.PIXEL_MEANS).astype(np.float32)[None]


def test_two_independent_restatements_agree_in_float64():
    w = synth.make_weights(0)
    for seed, (h, wd) in [(0, (64, 96)), (1, (50, 83))]:            # even and odd sizes: pooling floors at every level
        blob = _blob(seed, h, wd)
        a = net_cpu.forward(blob, w, dtype=torch.float64)
        b = net_alt.forward(blob, w)
        for k in ("rpn_cls_score", "rpn_bbox_pred", "rpn_cls_prob_reshape"):
            assert a[k].shape == b[k].shape
            d = np.abs(a[k] - b[k]).max()
            assert d <= 1e-11 * max(1.0, np.abs(b[k]).max()), (k, d)


def test_restatements_agree_on_a_batch_and_reverse_direction_matters():
    w = synth.make_weights(3)
    blob = np.concatenate([_blob(5, 48, 64), _blob(6, 48, 64)])
    a = net_cpu.forward(blob, w, dtype=torch.float64)
    b = net_alt.forward(blob, w)
    assert np.abs(a["rpn_cls_score"] - b["rpn_cls_score"]).max() <= 1e-11
    # sanity of the check itself: swapping the two directions' weights must be visible
    w2 = dict(w)
    w2[net_cpu.LSTM_FW + "/kernel"], w2[net_cpu.LSTM_BW + "/kernel"] = w[net_cpu.LSTM_BW + "/kernel"], w[net_cpu.LSTM_FW + "/kernel"]
    c = net_alt.forward(blob, w2)
    assert np.abs(a["rpn_cls_score"] - c["rpn_cls_score"]).max() > 1e-3
