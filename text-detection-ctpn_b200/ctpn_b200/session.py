"""Duck-typed stand-ins for the two TensorFlow objects the reference's call sites touch
(ctpn/demo.py:79-105, lib/fast_rcnn/test.py:40-58): a Session with .run(fetches, feed_dict)
and the placeholders / output handle of VGGnet_test (lib/networks/VGGnet_test.py:7-14)."""
import numpy as np

from .engine import Engine


class Placeholder:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "<ctpn_b200 placeholder %s>" % self.name


class OutputHandle:
    def __init__(self, name):
        self.name = name


class Session:
    """Replaces tf.Session for the test path.  `restore(weights)` replaces
    tf.train.Saver().restore (demo.py:85-90)."""

    def __init__(self, weights=None, planes=2, device=0, config=None, engine=None):
        self.engine = engine or Engine(weights, planes=planes, device=device)

    def restore(self, weights):
        self.engine.load_weights(weights)

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        data = info = None
        for k, v in (feed_dict or {}).items():
            if getattr(k, "name", None) == "data":
                data = v
            elif getattr(k, "name", None) == "im_info":
                info = v
        if data is None or info is None:
            raise ValueError("feed_dict must provide net.data and net.im_info")
        data = np.asarray(data)
        info = np.asarray(info, np.float32).reshape(-1, 3)
        # batch-1 only, like the reference (proposal_layer_tf.py:51-52)
        assert data.shape[0] == 1, "Only single item batches are supported"
        rois = self.engine.rois_batch(data, info)[0]
        out = []
        for f in fl:
            name = getattr(f, "name", f)
            if name in ("rois", "rpn_rois"):
                out.append(rois)
            else:
                raise KeyError("cannot fetch %r from the ctpn_b200 session" % (name,))
        return out[0] if single else out
