"""nms(dets, thresh) -- dispatch of lib/fast_rcnn/nms_wrapper.py:11-20.  This build has one
back-end: the CUDA path (gpu_nms -> ctpn_nms_host).  The reference's cython / pure-python CPU
fallbacks are deliberately absent: a missing library is an ImportError, not a slow path."""
from lib.utils.gpu_nms import gpu_nms
from .config import cfg


def nms(dets, thresh):
    if dets.shape[0] == 0:
        return []
    return gpu_nms(dets, thresh, device_id=cfg.GPU_ID)
