"""CPU oracle for the CTPN detection hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference algorithm
(eragonruan/text-detection-ctpn @ c04a571e).  It exists to check the CUDA
path; nothing in the product package (``text-detection-ctpn_b200/``) imports
it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.

Pinning status (see DESIGN.md "Oracle"):
  * post-processing half (anchors, bbox decode/clip/filter, sort, NMS,
    TextDetector H/O): PINNED against outputs of the reference's own numpy
    code, imported unmodified from /root/reference by
    ``tests/golden/make_golden.py`` (fixtures in ``tests/golden/*.npz``).
  * network half (VGG16 / BiLSTM / heads / softmax): PARITY UNPINNED.  The
    arithmetic lives in TensorFlow 1.3 (requirements.txt:2), which is neither
    vendored nor installable here; ``net_cpu.py`` restates the documented
    TF 1.3 semantics on torch-CPU.  Narrowed, not closed: ``net_alt.py`` is a
    second, independently written restatement (numpy im2col convolutions,
    torch.nn.LSTM with permuted gate columns) and tests/test_oracle_net_cpu.py
    asserts the two agree to 1e-11 in float64.  The graph WIRING is pinned:
    tests/golden/make_golden_net.py runs the reference's own VGGnet_test /
    network.py / test_ctpn / py_func proposal layer, unmodified, on a numpy
    stand-in for the TensorFlow functions they call (tests/golden/tf1_stub),
    and net_cpu.py + postproc.py match its tensors and proposals to float32
    rounding; the per-op TF semantics restated in that stub stay unpinned.
  * image front-end (``resize.py``): uint8 INTER_LINEAR PINNED against
    cv2.resize; float32 INTER_LINEAR PINNED against OpenCV's own code path
    (cv2 with IPP disabled; IPP-dispatching builds differ, see the docstring).
  * training-side target ops (``train_targets.py``: bbox_overlaps,
    bbox_intersections, bbox_transform, anchor_target_layer): PINNED against
    outputs of the reference's own operator run on its own Cython bbox module
    (re-cythonized into oracle/_ref by oracle/Makefile),
    ``tests/golden/make_golden_train.py``.
  * ``quant.py``: this repo's own F16F8 operand format (no counterpart in the
    reference), checked against torch's float16 / float8_e4m3fn conversions.
"""
