"""proposal_layer(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, cfg_key, ...) -> (blob, deltas):
the operator interface of lib/rpn_msr/proposal_layer_tf.py:14 (called through tf.py_func at
lib/networks/network.py:214 and directly at ctpn/demo_pb.py:92), running on the device via
ctpn_proposals (decode + clip + filter + sort + NMS + top-N fused on the GPU)."""
import numpy as np
import torch

from lib.fast_rcnn.config import cfg

_engine_cache = {}


def _proposal_engine():
    # a weight-less Engine: only its workspace management and ctpn_proposals binding are used
    from ctpn_b200.engine import Engine
    dev = cfg.GPU_ID
    if dev not in _engine_cache:
        _engine_cache[dev] = Engine(None, device=dev)
    return _engine_cache[dev]


def proposal_layer(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, cfg_key, _feat_stride=[16, ], anchor_scales=[16, ]):
    if isinstance(cfg_key, bytes):
        cfg_key = cfg_key.decode('ascii')
    if cfg_key != 'TEST':
        # the training graph (anchor targets, losses) is out of scope of this inference engine
        raise NotImplementedError("proposal_layer: only cfg_key='TEST' is supported (got %r)" % (cfg_key,))
    cls_prob = np.ascontiguousarray(rpn_cls_prob_reshape, np.float32)
    bbox = np.ascontiguousarray(rpn_bbox_pred, np.float32)
    assert cls_prob.shape[0] == 1, 'Only single item batches are supported'
    c = cfg[cfg_key]
    eng = _proposal_engine()
    over = dict(RPN_PRE_NMS_TOP_N=c.RPN_PRE_NMS_TOP_N, RPN_POST_NMS_TOP_N=c.RPN_POST_NMS_TOP_N,
                RPN_NMS_THRESH=c.RPN_NMS_THRESH, RPN_MIN_SIZE=c.RPN_MIN_SIZE, FEAT_STRIDE=int(_feat_stride[0]))
    info = torch.from_numpy(np.asarray(im_info, np.float32).reshape(1, 3))
    rois, index, count = eng.proposals(torch.from_numpy(cls_prob).to(eng.device), torch.from_numpy(bbox).to(eng.device),
                                       info, cls_is_logit=False, cfg=over)
    n = int(count[0])
    blob = rois[0, :n].cpu().numpy()
    idx = index[0, :n].cpu().numpy().astype(np.int64)
    deltas = bbox.reshape(-1, 4)[idx]
    return blob, deltas
