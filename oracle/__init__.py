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
    TF 1.3 semantics on torch-CPU and is cross-checked only against an
    independent float64 evaluation of itself.
"""
