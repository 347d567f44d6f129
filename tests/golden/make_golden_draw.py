#!/usr/bin/env python
"""Generate tests/golden/reference_draw_boxes.npz with the REFERENCE's own draw_boxes / resize_im (ctpn/demo.py:21-52).

ctpn/demo.py imports tensorflow at module level, so the two functions are taken out of the unmodified source file
with `ast` (by name) and executed as they stand in a namespace that provides np, cv2 and os.  Run in the build
container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_draw.py
Inputs: the reference TextDetector outputs already frozen in reference_postproc.npz plus seeded synthetic images.
"""
import ast
import os
import tempfile

import cv2
import numpy as np

REF_DEMO = "/root/reference/ctpn/demo.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_functions(names):
    src = open(REF_DEMO).read()
    tree = ast.parse(src)
    ns = {"np": np, "cv2": cv2, "os": os}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF_DEMO, "exec"), ns)
    return [ns[n] for n in names]


def main():
    draw_boxes, resize_im = reference_functions(["draw_boxes", "resize_im"])
    gold = np.load(os.path.join(HERE, "reference_postproc.npz"))
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        os.makedirs("data/results")
        try:
            for k, (mode, seed, scale) in enumerate([("H", 0, 1.25), ("O", 1, 0.75), ("H", 2, 1.0), ("O", 3, 600.0 / 410.0)]):
                boxes = gold["text_%s_%d" % (mode, seed)].copy()
                boxes[0, 1] = boxes[0, 0] + 1.0                     # trips the |box[0] - box[1]| < 5 skip of demo.py:32
                img = np.full((600, 900, 3), 40 * (k + 1), np.uint8)     # flat image: the fixture stays small
                img[::50] = 255 - 40 * (k + 1)                            # a few rows of structure for the final resize
                name = "some/dir/pic_%d.png" % k
                draw_boxes(img, name, boxes, scale)                  # draws into img in place, writes txt + image
                out["boxes_%d" % k] = boxes
                out["scale_%d" % k] = np.float64(scale)
                out["res_%d" % k] = np.frombuffer(open("data/results/res_pic_%d.txt" % k, "rb").read(), np.uint8)
                out["image_%d" % k] = cv2.imread("data/results/pic_%d.png" % k)
            # resize_im with reduced targets (120 / 240 instead of 600 / 1200) so that random images stay small
            for k, (h, w) in enumerate([(60, 90), (240, 320), (50, 240), (97, 211), (33, 29)]):
                im = np.random.RandomState(90 + k).randint(0, 256, (h, w, 3)).astype(np.uint8)
                rim, f = resize_im(im, scale=120, max_scale=240)
                out["resize_shape_%d" % k] = np.array([h, w])
                out["resize_f_%d" % k] = np.float64(f)
                out["resize_out_%d" % k] = rim
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "reference_draw_boxes.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_draw_boxes.npz"), {k: v.shape for k, v in out.items() if k.startswith(("res_", "image_"))})


if __name__ == "__main__":
    main()
