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


# tensor names demo_pb.py fetches from the frozen graph (ctpn/demo_pb.py:73-75)
PB_INPUT = "Placeholder:0"
PB_CLS_PROB = "Reshape_2:0"                      # rpn_cls_prob_reshape [1,H,W,20] (pair softmax applied)
PB_BOX_PRED = "rpn_bbox_pred/Reshape_1:0"        # rpn_bbox_pred [1,H,W,40]


class _Graph:
    """sess.graph.get_tensor_by_name for the three tensors the frozen-graph demo uses."""

    def get_tensor_by_name(self, name):
        if name not in (PB_INPUT, PB_CLS_PROB, PB_BOX_PRED):
            raise KeyError("The name %r refers to a Tensor which does not exist in the ctpn_b200 graph" % (name,))
        return Placeholder(name) if name == PB_INPUT else OutputHandle(name)


class Session:
    """Replaces tf.Session for the test path.  `restore(weights)` replaces
    tf.train.Saver().restore (demo.py:85-90) and tf.import_graph_def of a frozen .pb (demo_pb.py:66-70)."""
    graph = _Graph()

    def __init__(self, weights=None, planes=2, device=0, config=None, engine=None):
        self.engine = engine or Engine(weights, planes=planes, device=device)

    def restore(self, weights):
        self.engine.load_weights(weights)

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        data = info = None
        for k, v in (feed_dict or {}).items():
            if getattr(k, "name", None) in ("data", PB_INPUT):
                data = v
            elif getattr(k, "name", None) == "im_info":
                info = v
        names = [getattr(f, "name", f) for f in fl]
        if data is not None and all(n in (PB_CLS_PROB, PB_BOX_PRED) for n in names):
            # the frozen-graph boundary (demo_pb.py:90): head tensors out, proposal_layer is called by the caller
            import torch
            data = np.asarray(data)
            assert data.shape[0] == 1, "Only single item batches are supported"
            dt = torch.uint8 if data.dtype == np.uint8 else torch.float32
            dev = torch.from_numpy(np.ascontiguousarray(data, np.uint8 if dt == torch.uint8 else np.float32)).to(self.engine.device)
            cls, box = self.engine.forward_heads(dev)
            prob = torch.softmax(cls.reshape(-1, 2), dim=1).reshape(cls.shape)      # spatial softmax over (bg, fg) pairs
            vals = {PB_CLS_PROB: prob.cpu().numpy(), PB_BOX_PRED: box.cpu().numpy()}
            out = [vals[n] for n in names]
            return out[0] if single else out
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
