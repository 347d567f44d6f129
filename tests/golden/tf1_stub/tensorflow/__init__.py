"""Test infrastructure: a numpy stand-in for the small part of the TensorFlow 1.x graph API that the reference's inference
code touches, so that the reference's OWN graph-building code (lib/networks/network.py, VGGnet_test.py, lib/fast_rcnn/test.py,
lib/rpn_msr/proposal_layer_tf.py) can be imported unmodified and executed in the build container, where TensorFlow 1.3
(requirements.txt:2) cannot be installed.  tests/golden/make_golden_net.py puts this directory on sys.path.

What this pins and what it does not: the graph's WIRING is the reference's (layer order, variable scopes and names, the
[N*H, W, C] row sequences of Bilstm, fw/bw concatenation, the spatial reshape / pair softmax, the py_func call of the
proposal layer, test_ctpn's blob and im_info handling).  The SEMANTICS OF EACH OP below are a restatement of TensorFlow 1.3's
documented behaviour, written here and not checked against TensorFlow:
  conv2d        NHWC x HWIO cross-correlation; SAME: out = ceil(in / stride), pad_total = max((out-1)*stride + k - in, 0),
                pad_before = pad_total // 2 (the extra pixel goes after)
  max_pool      VALID: out = floor((in - k) / stride) + 1, windows that do not fit are dropped
  LSTMCell      tf.contrib.rnn.LSTMCell (no peepholes / projection / clipping): variables <scope>/lstm_cell/kernel
                [input + units, 4 units] and .../bias [4 units]; z = concat([x, h]) . kernel + bias; i, j, f, o = split(z, 4);
                c' = sigmoid(f + forget_bias) * c + sigmoid(i) * tanh(j); h' = sigmoid(o) * tanh(c'); forget_bias = 1.0;
                zero initial state
  bidirectional_dynamic_rnn   scopes <scope>/bidirectional_rnn/fw and .../bw; the backward cell reads the sequence reversed
                along time and its outputs are reversed back; returns ((out_fw, out_bw), states)
  softmax       over the last axis
  py_func       calls the Python function on the evaluated inputs.  A Python str argument is handed over as str: that is
                what the reference's author ran (Python 2); TensorFlow under Python 3 hands it over as bytes and the
                reference then fails with KeyError
Tensors are lazy nodes evaluated by Session.run(fetches, feed_dict) in float32 numpy.
"""
import contextlib
import types

import numpy as np

float32 = np.float32
int32 = np.int32


class StaticShape:
    def __init__(self, dims):
        self.dims = None if dims is None else list(dims)

    @property
    def ndims(self):
        return None if self.dims is None else len(self.dims)

    def as_list(self):
        return list(self.dims)

    def __getitem__(self, i):
        r = self.dims[i]
        return StaticShape(r) if isinstance(i, slice) else r


class Tensor:
    def __init__(self, fn, inputs=(), static=None, name=None):
        self.fn, self.inputs, self.static, self.name = fn, tuple(inputs), static, name

    def get_shape(self):
        return StaticShape(self.static)

    def set_shape(self, dims):
        self.static = list(dims)

    def __getitem__(self, idx):
        return Tensor(lambda v: v[idx], [self])

    def _binary(self, other, op, swap=False):
        return Tensor((lambda a, b: op(b, a)) if swap else op, [self, other])

    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, np.multiply, True)
    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, np.add, True)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __truediv__(self, o): return self._binary(o, np.true_divide)
    def __rtruediv__(self, o): return self._binary(o, np.true_divide, True)


def _value(x, env):
    if isinstance(x, Tensor):
        key = id(x)
        if key not in env.memo:
            env.memo[key] = x.fn(*[_value(i, env) for i in x.inputs])
        return env.memo[key]
    if isinstance(x, (list, tuple)):
        return type(x)(_value(i, env) for i in x)
    return x


# ---- sessions, placeholders, variables -----------------------------------------------------------------------------------
class ConfigProto:
    def __init__(self, **kwargs):
        self.kwargs = kwargs


class Session:
    current = None

    def __init__(self, config=None):
        self.variables = {}
        self.memo = {}
        self.feeds = {}
        Session.current = self

    def load_variables(self, variables):          # stub-only: stands for tf.train.Saver().restore(sess, checkpoint)
        self.variables = {k: np.asarray(v, np.float32) for k, v in variables.items()}

    def run(self, fetches, feed_dict=None):
        self.memo = {}
        self.feeds = {id(k): np.asarray(v) for k, v in (feed_dict or {}).items()}
        return _value(fetches, self)


def placeholder(dtype, shape=None, name=None):
    node = Tensor(None, static=shape, name=name)

    def read():
        feeds = Session.current.feeds
        if id(node) not in feeds:
            raise RuntimeError("placeholder %r was not fed" % (name,))
        return feeds[id(node)].astype(dtype, copy=False)
    node.fn = read
    return node


_scope_stack = []
requested_variables = []          # full names in creation order (make_golden_net.py checks them against SURVEY App. A.2)


class _Scope:
    def __init__(self, name):
        self.name = name


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _scope_stack.append(name)
    try:
        yield _Scope("/".join(_scope_stack))
    finally:
        _scope_stack.pop()


def _variable(full_name, shape=None):
    requested_variables.append(full_name)

    def read():
        store = Session.current.variables
        if full_name not in store:
            raise KeyError("variable %s is not in the restored checkpoint" % full_name)
        v = store[full_name]
        if shape is not None and all(isinstance(d, (int, np.integer)) for d in shape) and tuple(v.shape) != tuple(int(d) for d in shape):
            raise ValueError("variable %s has shape %s, the graph asks for %s" % (full_name, v.shape, tuple(shape)))
        return v
    return Tensor(read, static=None if shape is None else list(shape), name=full_name)


def get_variable(name, shape=None, initializer=None, trainable=True, regularizer=None):
    return _variable("/".join(_scope_stack + [name]), shape)


def truncated_normal_initializer(mean=0.0, stddev=1.0):
    return ("truncated_normal", mean, stddev)


def constant_initializer(value=0.0):
    return ("constant", value)


# ---- array ops -------------------------------------------------------------------------------------------------------------
def shape(x, name=None):
    return Tensor(lambda v: np.array(v.shape, np.int32), [x])


def reshape(x, new_shape, name=None):
    return Tensor(lambda v, s: np.reshape(v, [int(d) for d in s]), [x, list(new_shape) if isinstance(new_shape, (list, tuple)) else new_shape])


def transpose(x, perm, name=None):
    return Tensor(lambda v: np.transpose(v, perm), [x])


def concat(values, axis, name=None):
    return Tensor(lambda vs: np.concatenate(list(vs), axis=axis), [tuple(values)])


def cast(x, dtype):
    return Tensor(lambda v: np.asarray(v).astype(dtype), [x])


def convert_to_tensor(x, name=None, dtype=None):
    return x


def matmul(a, b):
    return Tensor(lambda u, v: np.matmul(u, v), [a, b])


def py_func(func, inp, Tout):
    call = Tensor(lambda args: func(*args), [tuple(inp)])
    return [Tensor(lambda res, k=k, t=t: np.asarray(res[k]).astype(t, copy=False), [call]) for k, t in enumerate(Tout)]


# ---- nn ----------------------------------------------------------------------------------------------------------------------
def _same_padding(size, k, stride):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2, total - total // 2


def _conv2d(x, w, strides, padding):
    n, h, wd, c = x.shape
    kh, kw, ci, co = w.shape
    assert ci == c and strides[0] == 1 and strides[3] == 1
    sh, sw = strides[1], strides[2]
    if padding == "SAME":
        oh, pt, pb = _same_padding(h, kh, sh)
        ow, pl, pr = _same_padding(wd, kw, sw)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    else:
        oh, ow = (h - kh) // sh + 1, (wd - kw) // sw + 1
    out = np.zeros((n, oh, ow, co), np.float32)
    for ky in range(kh):
        for kx in range(kw):
            patch = x[:, ky:ky + (oh - 1) * sh + 1:sh, kx:kx + (ow - 1) * sw + 1:sw, :]
            out += np.matmul(patch.reshape(-1, c), w[ky, kx]).reshape(n, oh, ow, co)
    return out


def _max_pool(x, ksize, strides, padding):
    assert padding == "VALID" and ksize[0] == ksize[3] == 1 and strides[0] == strides[3] == 1
    kh, kw, sh, sw = ksize[1], ksize[2], strides[1], strides[2]
    n, h, w, c = x.shape
    oh, ow = (h - kh) // sh + 1, (w - kw) // sw + 1
    out = np.full((n, oh, ow, c), -np.inf, np.float32)
    for ky in range(kh):
        for kx in range(kw):
            out = np.maximum(out, x[:, ky:ky + (oh - 1) * sh + 1:sh, kx:kx + (ow - 1) * sw + 1:sw, :])
    return out


def _softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)


def _sigmoid(x):
    return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)


class LSTMCell:
    def __init__(self, num_units, state_is_tuple=True, forget_bias=1.0):
        assert state_is_tuple
        self.num_units, self.forget_bias = int(num_units), np.float32(forget_bias)

    def run(self, x, kernel, bias, reverse):
        """x [batch, time, depth] -> h of every step [batch, time, units]."""
        b, t, _ = x.shape
        u = self.num_units
        h = np.zeros((b, u), np.float32)
        c = np.zeros((b, u), np.float32)
        out = np.zeros((b, t, u), np.float32)
        seq = x[:, ::-1] if reverse else x                  # array_ops.reverse_sequence over the full length
        for s in range(t):
            z = np.matmul(np.concatenate([seq[:, s], h], axis=1), kernel) + bias
            i, j, f, o = np.split(z, 4, axis=1)
            c = _sigmoid(f + self.forget_bias) * c + _sigmoid(i) * np.tanh(j)
            h = _sigmoid(o) * np.tanh(c)
            out[:, s] = h
        return out[:, ::-1] if reverse else out


def _bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, dtype=None, sequence_length=None):
    assert sequence_length is None
    outs = []
    for direction, cell, reverse in (("fw", cell_fw, False), ("bw", cell_bw, True)):
        prefix = "/".join(_scope_stack + ["bidirectional_rnn", direction, "lstm_cell"])
        kernel, bias = _variable(prefix + "/kernel"), _variable(prefix + "/bias")
        outs.append(Tensor(lambda x, k, b, cell=cell, reverse=reverse: cell.run(x.astype(np.float32), k, b, reverse), [inputs, kernel, bias]))
    return (outs[0], outs[1]), None


nn = types.SimpleNamespace(
    conv2d=lambda input, filter, strides, padding, name=None: Tensor(
        lambda x, w: _conv2d(x, w, strides, padding), [input, filter],
        static=[None, None, None, filter.static[-1] if filter.static else None]),
    bias_add=lambda value, bias, name=None: Tensor(lambda v, b: v + b, [value, bias], static=value.static),
    relu=lambda features, name=None: Tensor(lambda v: np.maximum(v, np.float32(0)), [features], static=features.static),
    max_pool=lambda value, ksize, strides, padding, name=None: Tensor(
        lambda v: _max_pool(v, ksize, strides, padding), [value], static=[None, None, None, value.static[-1] if value.static else None]),
    softmax=lambda logits, name=None: Tensor(_softmax, [logits]),
    bidirectional_dynamic_rnn=_bidirectional_dynamic_rnn,
)
contrib = types.SimpleNamespace(rnn=types.SimpleNamespace(LSTMCell=LSTMCell))
