"""im_list_to_blob (lib/utils/blob.py:6-19): zero-pad a list of HxWx3 float images into one
NHWC float32 blob."""
import numpy as np


def im_list_to_blob(ims):
    max_shape = np.array([im.shape for im in ims]).max(axis=0)
    blob = np.zeros((len(ims), max_shape[0], max_shape[1], 3), dtype=np.float32)
    for i, im in enumerate(ims):
        blob[i, 0:im.shape[0], 0:im.shape[1], :] = im
    return blob
