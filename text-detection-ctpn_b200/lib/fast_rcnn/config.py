"""Global config object, same surface as the reference's lib/fast_rcnn/config.py for the keys
the detection path reads (config.py:7-16, 147-183, 200, 256-316).  easydict is not a
dependency: a small attribute-dict is defined here."""
import os.path as osp

import numpy as np


class edict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, edict):
            v = edict(v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


__C = edict()
cfg = __C

__C.GPU_ID = 0
__C.IS_RPN = True
__C.ANCHOR_SCALES = [16]
__C.NCLASSES = 2
__C.USE_GPU_NMS = True
__C.IS_MULTISCALE = False
__C.IS_EXTRAPOLATING = True
__C.REGION_PROPOSAL = 'RPN'
__C.NET_NAME = 'VGGnet'
__C.SUBCLS_NAME = 'voxel_exemplars'
__C.EXP_DIR = 'default'
__C.LOG_DIR = 'default'

# text.yml carries a TRAIN block; the keys are accepted.  The inference engine ignores them; the RPN_* / DONTCARE /
# PRECLUDE keys are read by lib/rpn_msr/anchor_target_layer_tf.py (config.py:112-140)
__C.TRAIN = edict(dict(
    restore=0, max_steps=100000, SOLVER='Momentum', OHEM=False, RPN_BATCHSIZE=256, BATCH_SIZE=128,
    LOG_IMAGE_ITERS=100, DISPLAY=10, SNAPSHOT_ITERS=5000, HAS_RPN=False, LEARNING_RATE=0.001, MOMENTUM=0.9,
    GAMMA=0.1, STEPSIZE=50000, IMS_PER_BATCH=2, BBOX_NORMALIZE_TARGETS_PRECOMPUTED=False,
    RPN_POSITIVE_OVERLAP=0.7, PROPOSAL_METHOD='selective_search', BG_THRESH_LO=0.1,
    PRECLUDE_HARD_SAMPLES=True, BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
    RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0), RPN_POSITIVE_WEIGHT=-1.0, FG_FRACTION=0.25,
    WEIGHT_DECAY=0.0005, RPN_NEGATIVE_OVERLAP=0.3, RPN_CLOBBER_POSITIVES=False, RPN_FG_FRACTION=0.5,
    DONTCARE_AREA_INTERSECTION_HI=0.5))

__C.TEST = edict()
__C.TEST.checkpoints_path = "checkpoints/"
__C.TEST.DETECT_MODE = "H"          # H/O for horizontal/oriented mode
__C.TEST.SCALES = (600,)
__C.TEST.MAX_SIZE = 1000
__C.TEST.NMS = 0.3
__C.TEST.SVM = False
__C.TEST.BBOX_REG = True
__C.TEST.HAS_RPN = True
__C.TEST.PROPOSAL_METHOD = 'selective_search'
__C.TEST.RPN_NMS_THRESH = 0.7
__C.TEST.RPN_PRE_NMS_TOP_N = 12000
__C.TEST.RPN_POST_NMS_TOP_N = 1000
__C.TEST.RPN_MIN_SIZE = 8

__C.DEDUP_BOXES = 1. / 16.
__C.PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])
__C.RNG_SEED = 3
__C.EPS = 1e-14
__C.ROOT_DIR = osp.abspath(osp.join(osp.dirname(__file__), '..', '..'))
__C.DATA_DIR = osp.abspath(osp.join(__C.ROOT_DIR, 'data'))


def _merge_a_into_b(a, b, lenient=()):
    """config.py:256-286: keys of a must exist in b with the same type."""
    if not isinstance(a, dict):
        return
    for k, v in a.items():
        if k not in b:
            raise KeyError('{} is not a valid config key'.format(k))
        old_type = type(b[k])
        if old_type is not type(v):
            if isinstance(b[k], np.ndarray):
                v = np.array(v, dtype=b[k].dtype)
            elif isinstance(b[k], (tuple, list)) and isinstance(v, (tuple, list)):
                v = old_type(v)
            elif isinstance(b[k], dict) and isinstance(v, dict):
                pass
            elif isinstance(b[k], float) and isinstance(v, int):
                v = float(v)
            else:
                raise ValueError('Type mismatch ({} vs. {}) for config key: {}'.format(type(b[k]), type(v), k))
        if isinstance(v, dict):
            try:
                _merge_a_into_b(v, b[k])
            except Exception:
                print('Error under config key: {}'.format(k))
                raise
        else:
            b[k] = v


def cfg_from_file(filename):
    """config.py:288-294 (yaml.safe_load: PyYAML >= 6 rejects the Loader-less yaml.load)."""
    import yaml
    with open(filename, 'r') as f:
        yaml_cfg = edict(yaml.safe_load(f))
    _merge_a_into_b(yaml_cfg, __C)


def cfg_from_list(cfg_list):
    """config.py:296-316."""
    from ast import literal_eval
    assert len(cfg_list) % 2 == 0
    for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
        key_list = k.split('.')
        d = __C
        for subkey in key_list[:-1]:
            assert subkey in d
            d = d[subkey]
        subkey = key_list[-1]
        assert subkey in d
        try:
            value = literal_eval(v)
        except Exception:
            value = v
        assert type(value) == type(d[subkey]), 'type {} does not match original type {}'.format(type(value), type(d[subkey]))
        d[subkey] = value
