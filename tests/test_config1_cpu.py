"""SURVEY.md §8(d) Config 1 (plumbing, no GPU): one synthetic image through the reference-facing call chain
ctpn(sess, net, image_name) -> resize_im -> test_ctpn -> sess.run -> TextDetector -> draw_boxes, with the session played by
the CPU oracle (torch-CPU network + numpy proposal layer) and the text lines built by the library's host C++ connector.
Pass = it runs, the result file and the annotated image are written in the reference's format."""
import os
import re

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import net_cpu, postproc, synth  # noqa: E402


class OracleSession:
    """Duck-types what test_ctpn needs from tf.Session: run([rois], {net.data, net.im_info, net.keep_prob})."""

    def __init__(self, weights):
        self.weights = weights
        self.calls = 0

    def run(self, fetches, feed_dict=None):
        data = info = None
        for k, v in feed_dict.items():
            if getattr(k, "name", None) == "data":
                data = v
            elif getattr(k, "name", None) == "im_info":
                info = v
        blob = np.asarray(data)
        if blob.dtype == np.uint8:                        # the mirror hands uint8 through at scale 1
            blob = (blob.astype(np.float64) - net_cpu.PIXEL_MEANS).astype(np.float32)
        out = net_cpu.forward(blob, self.weights)
        rois, _ = postproc.proposal_layer(out["rpn_cls_prob_reshape"], out["rpn_bbox_pred"], np.asarray(info, np.float32))
        self.calls += 1
        return [rois for _ in fetches]


def test_ctpn_call_chain_on_cpu(tmp_path, monkeypatch):
    import torch
    from ctpn import demo
    from lib.networks.factory import get_network
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    monkeypatch.setattr(demo, "RESULTS_DIR", str(tmp_path / "results"))
    monkeypatch.setattr(demo, "NATIVE_CONNECTOR", True)          # host C++ connector: no GPU anywhere on this path
    monkeypatch.setattr(demo.TextLineCfg, "SCALE", 96)           # keep the CPU network small: 48x72 -> 96x144
    monkeypatch.setattr(demo.TextLineCfg, "MAX_SCALE", 192)
    os.makedirs(demo.RESULTS_DIR)
    img = synth.make_image(0, 48, 72)
    name = str(tmp_path / "img_007.png")
    cv2.imwrite(name, img)
    sess, net = OracleSession(synth.make_weights(0)), get_network("VGGnet_test")
    demo.ctpn(sess, net, name)
    assert sess.calls == 1
    res = open(os.path.join(demo.RESULTS_DIR, "res_img_007.txt"), "rb").read()
    assert res == b"" or re.fullmatch(rb"(\d+,\d+,\d+,\d+\r\n)+", res)
    out_img = cv2.imread(os.path.join(demo.RESULTS_DIR, "img_007.png"))
    assert out_img is not None and out_img.shape == (48, 72, 3)   # annotated image scaled back by 1 / f
    with pytest.raises(KeyError):
        get_network("VGGnet_nope")
