"""get_network(name) (lib/networks/factory.py:4-14).  'VGGnet_test' returns the handle object
the test path needs (placeholders + get_output); the graph itself is the engine's fixed kernel
sequence.  'VGGnet_train' is out of scope for this inference engine."""
from ctpn_b200.session import OutputHandle, Placeholder


class VGGnet_test(object):
    """Mirror of the attributes lib/fast_rcnn/test.py:49-51 touches (VGGnet_test.py:7-14)."""

    def __init__(self, trainable=True):
        self.data = Placeholder("data")
        self.im_info = Placeholder("im_info")
        self.keep_prob = Placeholder("keep_prob")
        self.layers = {"data": self.data, "im_info": self.im_info,
                       "rois": (OutputHandle("rois"), OutputHandle("rpn_targets"))}
        self.trainable = trainable

    def get_output(self, layer):
        try:
            return self.layers[layer]
        except KeyError:
            print(list(self.layers.keys()))
            raise KeyError('Unknown layer name fed: %s' % layer)


def get_network(name):
    """Get a network by name."""
    if name.split('_')[0] == 'VGGnet':
        if name.split('_')[1] == 'test':
            return VGGnet_test()
        elif name.split('_')[1] == 'train':
            raise NotImplementedError("VGGnet_train: training is outside this engine's scope")
        else:
            raise KeyError('Unknown dataset: {}'.format(name))
    else:
        raise KeyError('Unknown dataset: {}'.format(name))
