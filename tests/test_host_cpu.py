"""CPU: host-side logic that mirrors the reference's Python (config, image pre-processing decisions,
text-line connector, result files).  The connector's second NMS normally runs on the GPU; here it is
replaced by the oracle's NMS so the host code can be checked against the reference goldens without a GPU."""
import os

import numpy as np
import pytest

from oracle import postproc, synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_postproc.npz"))


def test_config_defaults_and_yaml_overlay(tmp_path):
    from lib.fast_rcnn.config import cfg, cfg_from_file, cfg_from_list
    assert cfg.TEST.RPN_PRE_NMS_TOP_N == 12000 and cfg.TEST.RPN_POST_NMS_TOP_N == 1000
    assert cfg.TEST.RPN_NMS_THRESH == 0.7 and cfg.TEST.RPN_MIN_SIZE == 8 and cfg.TEST.SCALES == (600,)
    assert cfg.TEST.MAX_SIZE == 1000 and cfg.ANCHOR_SCALES == [16]
    np.testing.assert_allclose(cfg.PIXEL_MEANS.ravel(), [102.9801, 115.9465, 122.7717])
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "text-detection-ctpn_b200")
    cfg_from_file(os.path.join(pkg, "ctpn", "text.yml"))
    assert cfg.TEST.DETECT_MODE == "H" and cfg.TEST.HAS_RPN is True
    bad = tmp_path / "bad.yml"
    bad.write_text("NOT_A_KEY: 1\n")
    with pytest.raises(KeyError):
        cfg_from_file(str(bad))
    bad.write_text("TEST:\n  RPN_NMS_THRESH: 'x'\n")
    with pytest.raises(ValueError):
        cfg_from_file(str(bad))
    cfg_from_list(["TEST.DETECT_MODE", "O"])
    assert cfg.TEST.DETECT_MODE == "O"
    cfg_from_list(["TEST.DETECT_MODE", "H"])


def test_image_blob_decisions_match_reference_rules():
    from lib.fast_rcnn.test import _get_image_blob
    from oracle import net_cpu
    im = synth.make_image(1, 600, 900)
    blob, scales = _get_image_blob(im)                      # scale 1: raw uint8 goes to the device
    assert blob.dtype == np.uint8 and blob.shape == (1, 600, 900, 3) and scales[0] == 1.0
    im2 = synth.make_image(2, 300, 500)                     # needs the x2 resize -> float32 blob, identical to the oracle's
    blob2, scales2 = _get_image_blob(im2)
    want, s = net_cpu.image_blob(im2)
    assert scales2[0] == s == 2.0
    np.testing.assert_array_equal(blob2, want)
    im3 = synth.make_image(3, 500, 1500)                    # long side capped at MAX_SIZE
    _, scales3 = _get_image_blob(im3)
    assert scales3[0] == pytest.approx(1000.0 / 1500.0)


@pytest.mark.parametrize("mode", ["H", "O"])
def test_text_detector_host_code_matches_reference_goldens(mode, monkeypatch):
    import lib.text_connector.detectors as det
    from lib.fast_rcnn.config import cfg
    monkeypatch.setattr(cfg.TEST, "DETECT_MODE", mode, raising=False)
    cfg.TEST.DETECT_MODE = mode
    try:
        for seed in range(4):
            tp, sc = synth.make_text_proposals(seed)
            recs = det.TextDetector().detect(tp, sc, (600, 900))
            np.testing.assert_array_equal(recs, G["text_%s_%d" % (mode, seed)])
    finally:
        cfg.TEST.DETECT_MODE = "H"


def test_text_detector_empty_and_single():
    import lib.text_connector.detectors as det
    d = det.TextDetector()                                  # the whole of detect() is host code: no GPU needed
    out = d.detect(np.zeros((0, 4), np.float32), np.zeros((0, 1), np.float32), (600, 900))
    assert out.shape == (0, 9)
    out = d.detect(np.array([[16, 10, 32, 40]], np.float32), np.array([[0.99]], np.float32), (600, 900))
    assert out.shape == (0, 9)                          # a lone proposal forms no line


def test_connector_stages_match_python_graph_builder():
    """The C++ grouping (ctpn_text_groups_host) against the vectorised Python graph builder kept in the mirror, and the
    host NMS (ctpn_text_filter_nms_host) against the oracle's NMS, on the synthetic layouts."""
    from ctpn_b200 import textlines
    from lib.text_connector.text_proposal_graph_builder import TextProposalGraphBuilder
    for seed in range(30):
        tp, sc = synth.make_text_proposals(300 + seed, n_lines=2 + seed % 9, n_noise=40 + 7 * seed)
        keep = textlines.filter_nms(tp, sc)
        sel = np.where(sc.ravel() > 0.7)[0]
        order = sel[np.argsort(-sc.ravel()[sel], kind="stable")]
        want = order[postproc.nms(np.hstack((tp[order], sc[order])), 0.2)]
        np.testing.assert_array_equal(keep, want)
        got = textlines.groups(tp[keep], sc[keep], (600, 900))
        ref = TextProposalGraphBuilder().build_graph(tp[keep], sc[keep], (600, 900)).sub_graphs_connected()
        assert got == [list(map(int, c)) for c in ref]


def test_draw_boxes_writes_reference_format(tmp_path, monkeypatch):
    from ctpn import demo
    monkeypatch.setattr(demo, "RESULTS_DIR", str(tmp_path))
    img = np.zeros((120, 200, 3), np.uint8)
    boxes = np.array([[10.7, 20.2, 150.9, 20.2, 10.7, 60.8, 150.9, 60.8, 0.95],
                      [1, 1, 3, 1, 1, 3, 3, 3, 0.99]], np.float64)       # second box is skipped (too small, demo.py:32)
    demo.draw_boxes(img, "some/dir/pic_01.jpg", boxes, 2.0)
    txt = open(tmp_path / "res_pic_01.txt", "rb").read()
    assert txt == b"5,10,75,30\r\n"
    assert (tmp_path / "pic_01.jpg").exists()


def test_generate_anchors_table():
    from lib.rpn_msr.generate_anchors import generate_anchors
    np.testing.assert_array_equal(generate_anchors(), G["anchors"])
    np.testing.assert_array_equal(generate_anchors(py2=True), postproc.anchors(py2=True))


def test_product_and_oracle_synthetic_generators_agree():
    """bench.py's GPU arm draws its random-init weights from ctpn_b200.synthetic (the product never imports oracle/);
    the tests and the CPU baseline use oracle.synth.  Same seeds -> same tensors."""
    from ctpn_b200 import synthetic
    a, b = synthetic.make_weights(0), synth.make_weights(0)
    assert sorted(a) == sorted(b)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    np.testing.assert_array_equal(synthetic.make_image(3, 40, 50), synth.make_image(3, 40, 50))


def test_clip_boxes_in_place_x_then_y():
    from lib.text_connector.other import clip_boxes
    b = np.array([[-3.0, -2.0, 950.0, 700.0], [10.5, 599.5, 899.5, 20.0], [1, 2, 3, 4, 5, 6, 2000, -7]][:2], np.float32)
    q = np.array([[1, 2, 3, 4, 5, 6, 2000, -7]], np.float64)             # 8-column quadrilateral rows are clipped the same way
    out = clip_boxes(b, (600, 900))
    assert out is b
    assert np.array_equal(b, np.array([[0, 0, 899, 599], [10.5, 599, 899, 20]], np.float32))
    assert np.array_equal(clip_boxes(q, (600, 900)), [[1, 2, 3, 4, 5, 6, 899, 0]])
