"""Seeded random-init weights and synthetic images for benchmarking / smoke runs (BASELINE.json: "random-init VGG16",
"synthetic 600x900 batches").  No checkpoint or dataset is available offline, so the engine is exercised with a
variance-preserving initialisation (the reference's own stddev-0.01 initialisers collapse the activations to zero after
14 layers).  Variable names and shapes are those of the reference's TF checkpoint (App. A.2).
tests/test_host_cpu.py checks that this generator and the test-side one (oracle/synth.py) produce identical tensors."""
import numpy as np

CONV_LAYERS = [("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
               ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv4_1", 256, 512),
               ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv5_1", 512, 512), ("conv5_2", 512, 512),
               ("conv5_3", 512, 512), ("rpn_conv/3x3", 512, 512)]
LSTM_SCOPES = ("lstm_o/bidirectional_rnn/fw/lstm_cell", "lstm_o/bidirectional_rnn/bw/lstm_cell")
HEAD_CLS_STD, HEAD_BOX_STD = 0.21, 0.026     # logits ~ N(0, 2), (dy, dh) ~ N(0, 0.3) on 600x900 synthetic images


def make_weights(seed=0):
    rs = np.random.RandomState(seed)
    w = {}
    for name, cin, cout in CONV_LAYERS:
        std = np.sqrt(2.0 / (9 * cin))
        if name == "conv1_1":
            std /= 75.0     # mean-subtracted uint8 pixels have RMS ~75: bring activations to O(1)
        w[name + "/weights"] = (rs.standard_normal((3, 3, cin, cout)) * std).astype(np.float32)
        w[name + "/biases"] = (rs.standard_normal(cout) * 0.01).astype(np.float32)
    lim = np.sqrt(6.0 / (640 + 512))
    for scope in LSTM_SCOPES:
        w[scope + "/kernel"] = rs.uniform(-lim, lim, (640, 512)).astype(np.float32)
        w[scope + "/bias"] = (rs.standard_normal(512) * 0.01).astype(np.float32)
    w["lstm_o/weights"] = (rs.standard_normal((256, 512)) * np.sqrt(1.0 / 256)).astype(np.float32)
    w["lstm_o/biases"] = (rs.standard_normal(512) * 0.01).astype(np.float32)
    w["rpn_cls_score/weights"] = (rs.standard_normal((512, 20)) * HEAD_CLS_STD).astype(np.float32)
    w["rpn_cls_score/biases"] = (rs.standard_normal(20) * 0.01).astype(np.float32)
    w["rpn_bbox_pred/weights"] = (rs.standard_normal((512, 40)) * HEAD_BOX_STD).astype(np.float32)
    w["rpn_bbox_pred/biases"] = (rs.standard_normal(40) * 0.01).astype(np.float32)
    return w


def make_image(seed, h=600, w=900):
    """uint8 HWC BGR image, i.i.d. uniform pixels."""
    return np.random.RandomState(1000 + seed).randint(0, 256, size=(h, w, 3)).astype(np.uint8)
