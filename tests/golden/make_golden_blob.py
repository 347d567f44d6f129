#!/usr/bin/env python
"""Generate tests/golden/reference_image_blob.npz with the REFERENCE's own _get_image_blob
(lib/fast_rcnn/test.py:7-31, imported unmodified with the shims of make_golden.py) on small seeded images, with
cfg.TEST.SCALES / MAX_SIZE reduced to (60,) / 100 so that every branch is hit while the fixture stays small:
identity scale, up-scale, the MAX_SIZE cap, a cap that lands on scale 1.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

CASES = [(60, 90), (45, 60), (30, 100), (80, 200), (61, 97)]


def main():
    cfg = make_golden.load_reference()[0]
    from lib.fast_rcnn import test as ref_test
    assert os.path.realpath(ref_test.__file__).startswith(make_golden.REF)
    cfg.TEST.SCALES, cfg.TEST.MAX_SIZE = (60,), 100
    out = {"scales": np.array(cfg.TEST.SCALES), "max_size": np.array(cfg.TEST.MAX_SIZE)}
    for k, (h, w) in enumerate(CASES):
        im = np.random.RandomState(300 + k).randint(0, 256, (h, w, 3)).astype(np.uint8)
        blob, factors = ref_test._get_image_blob(im)
        out["shape_%d" % k] = np.array([h, w])
        out["blob_%d" % k] = blob
        out["factors_%d" % k] = factors
        print((h, w), "->", blob.shape, blob.dtype, factors)
    np.savez_compressed(os.path.join(HERE, "reference_image_blob.npz"), **out)


if __name__ == "__main__":
    main()
