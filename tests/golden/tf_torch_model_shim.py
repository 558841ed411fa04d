"""A torch-backed stand-in for the `tensorflow` / Keras symbols that the reference's MODEL code touches
(detr_tf/networks/{detr,resnet_backbone,transformer,custom_layers,position_embeddings}.py).

TensorFlow is not installable in the build container (no wheel, no network).  With this shim installed as
``sys.modules['tensorflow']`` the reference's model files are imported from /root/reference and EXECUTED UNMODIFIED --
`get_detr_model()` builds its layers, creates its variables and runs its `call()` methods line by line -- on float32 torch
tensors, to produce the golden vectors of tests/golden/model_golden.npz (see make_golden_model.py).  What the shim supplies
are Keras *library* semantics only (Conv2D / ZeroPadding2D / MaxPool2D / LayerNormalization / Dropout(training=False), the
tf.* array functions, variable naming by layer-name path); every line of model wiring is the reference's own.

The functional API (`tf.keras.Input`, `tf.keras.Model(inputs, outputs)`) is executed eagerly: `Input()` returns the concrete
image tensor set with `set_input()`, so `get_detr_model` computes real activations while it "builds the graph".

Test infrastructure only.
"""
import inspect
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

STATE = {"input": None, "weights": {}, "created": [], "outputs": {}, "layers": [], "autograd": False, "variables": {}}


class T(torch.Tensor):
    """TensorFlow tensors are immutable: `x += y` rebinds the name to a new tensor.  The reference relies on it (e.g.
    transformer.py:169 `source += ...` after `source` was fed to the attention); torch's in-place `+=` would overwrite values
    that autograd saved for the backward pass, so the in-place operators are made out-of-place."""

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other


def as_tf(x):
    return x.as_subclass(T)


def set_input(x):
    STATE["input"] = as_tf(x)
    return STATE["input"]


def set_weights(named, autograd=False):
    """name path (e.g. 'backbone/layer1/0/conv1/kernel') -> float32 torch tensor; every add_weight() must find its entry.
    autograd=True: trainable variables become leaves that require grad (STATE['variables'][name]) so that the gradient of a
    loss computed by the reference's code can be taken with torch.autograd (stand-in for tf.GradientTape)."""
    STATE["weights"] = dict(named)
    STATE["autograd"] = bool(autograd)
    STATE["variables"] = {}
    STATE["created"] = []
    STATE["outputs"] = {}
    STATE["layers"] = []
    Layer._counters = {}                    # a fresh Keras session: auto-names restart at 'dense', 'dense_1', ...


def _relink():
    """parents from attribute containment (lists included: the reference appends bottlenecks to a list after assigning it)"""
    for L in STATE["layers"]:
        for k, v in list(vars(L).items()):
            if k == "_parent":
                continue
            for e in (v if isinstance(v, (list, tuple)) else (v,)):
                if isinstance(e, Layer) and e is not L and e._parent is None:
                    object.__setattr__(e, "_parent", L)


class Layer:
    _counters = {}

    def __init__(self, name=None, **kwargs):
        assert not kwargs or set(kwargs) <= {"dtype", "trainable"}, kwargs
        if name is None:
            base = type(self).__name__.lower()
            n = Layer._counters.get(base, 0)
            Layer._counters[base] = n + 1
            name = base if n == 0 else f"{base}_{n}"
        object.__setattr__(self, "_parent", None)
        self.name = name
        self._built = False
        STATE["layers"].append(self)

    # ---- variable naming: the path of layer names below the root model, as Keras scopes variable names
    def _path(self):
        names, node = [], self
        while node is not None:
            names.append((node.name, type(node).__name__))
            node = node._parent
        if names[-1][1] == "DETR":          # the layers of the DETR container are re-used at the top level of the functional
            names.pop()                     # model (detr.py:149-177): their variables carry no 'detr/' prefix
        return "/".join(n for n, _ in reversed(names))

    def add_weight(self, name=None, shape=None, initializer=None, dtype=None, trainable=True):
        full = self._path() + "/" + name
        assert full in STATE["weights"], f"no injected value for variable {full!r}"
        w = STATE["weights"][full]
        assert tuple(w.shape) == tuple(int(s) for s in shape), (full, tuple(w.shape), tuple(shape))
        STATE["created"].append((full, tuple(w.shape), bool(trainable)))
        v = as_tf(w.detach().clone())
        if STATE["autograd"] and trainable:
            v.requires_grad_(True)
        STATE["variables"][full] = v
        return v

    def build(self, input_shape):
        pass

    def __call__(self, *args, **kwargs):
        if not self._built:
            _relink()
            x = args[0] if args else None
            if isinstance(x, torch.Tensor):
                shape = tuple(x.shape)
            elif isinstance(x, (list, tuple)):
                shape = [tuple(t.shape) for t in x]
            else:
                shape = None
            self.build(shape)
            self._built = True
        params = inspect.signature(self.call).parameters
        if "training" in kwargs and "training" not in params:        # Keras drops the argument for layers that do not take it
            kwargs.pop("training")
        out = self.call(*args, **kwargs)
        STATE["outputs"].setdefault(self._path() or self.name, out)
        return out


class Model(Layer):
    def __init__(self, *args, **kwargs):
        if len(args) == 2:                                            # functional: Model(inputs, outputs, name=...)
            super().__init__(name=kwargs.pop("name", None))
            self._inputs, self._outputs = args
            return
        assert not args
        super().__init__(**kwargs)

    def call(self, x, **kwargs):                                      # functional models only: same concrete input -> stored output
        assert x is self._inputs
        return self._outputs

    def get_layer(self, name):
        for v in vars(self).values():
            if isinstance(v, Layer) and v.name == name:
                return v
        raise ValueError(name)


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        Layer.__init__(self, name=name)
        self.seq = list(layers or [])

    def call(self, x):
        for layer in self.seq:
            x = layer(x)
        return x


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, dilation_rate=1, **kw):
        super().__init__(**kw)
        assert padding == "valid"
        self.filters, self.k, self.s, self.use_bias, self.d = filters, kernel_size, strides, use_bias, dilation_rate

    def build(self, input_shape):
        self.kernel = self.add_weight(name="kernel", shape=(self.k, self.k, input_shape[-1], self.filters))
        self.bias = self.add_weight(name="bias", shape=(self.filters,)) if self.use_bias else None

    def call(self, x):
        return _nhwc(F.conv2d(_nchw(x), self.kernel.permute(3, 2, 0, 1), self.bias, stride=self.s, padding=0, dilation=self.d))


class ZeroPadding2D(Layer):
    def __init__(self, padding=1, **kw):
        super().__init__(**kw)
        self.p = padding

    def call(self, x):
        return F.pad(x, (0, 0, self.p, self.p, self.p, self.p))


class ReLU(Layer):
    def call(self, x):
        return torch.relu(x)


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        assert activation == "relu"

    def call(self, x):
        return torch.relu(x)


class MaxPool2D(Layer):
    def __init__(self, pool_size=2, strides=None, padding="valid", **kw):
        super().__init__(**kw)
        assert padding == "valid"
        self.k, self.s = pool_size, strides or pool_size

    def call(self, x):
        return _nhwc(F.max_pool2d(_nchw(x), self.k, self.s))


class Dropout(Layer):
    def __init__(self, rate=0.0, **kw):
        super().__init__(**kw)

    def call(self, x, training=False):
        assert not training, "the golden vectors are taken in inference mode (TF's dropout RNG cannot be reproduced)"
        return x


class LayerNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kw):
        super().__init__(**kw)
        self.eps = epsilon

    def build(self, input_shape):
        self.gamma = self.add_weight(name="gamma", shape=(input_shape[-1],))
        self.beta = self.add_weight(name="beta", shape=(input_shape[-1],))

    def call(self, x):
        mean = x.mean(-1, keepdim=True)
        var = ((x - mean) ** 2).mean(-1, keepdim=True)
        return (x - mean) * torch.rsqrt(var + self.eps) * self.gamma + self.beta


class Dense(Layer):
    def __init__(self, units, activation=None, **kw):
        super().__init__(**kw)
        self.units, self.act = units, activation

    def build(self, input_shape):
        self.kernel = self.add_weight(name="kernel", shape=(input_shape[-1], self.units))
        self.bias = self.add_weight(name="bias", shape=(self.units,))

    def call(self, x):
        y = x @ self.kernel + self.bias
        return torch.relu(y) if self.act == "relu" else torch.sigmoid(y) if self.act == "sigmoid" else y


def _ints(shape):
    return [int(s) for s in shape]


def build():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32, tf.int64, tf.bool = torch.float32, torch.int32, torch.int64, torch.bool
    tf.Tensor = torch.Tensor
    tf.newaxis = None
    tf.shape = lambda x: list(x.shape)
    tf.Variable = lambda v, **kw: v                                   # training_config.py:66-68 wraps the learning rates
    tf.reshape = lambda x, shape: x.reshape(_ints(shape)).clone()
    tf.transpose = lambda x, perm: x.permute(*perm).contiguous()
    tf.matmul = lambda a, b, transpose_b=False: a @ (b.transpose(-1, -2) if transpose_b else b)
    tf.zeros = lambda shape, dtype=torch.float32: torch.zeros(_ints(shape), dtype=dtype)
    tf.zeros_like = lambda x: torch.zeros_like(x)
    tf.expand_dims = lambda x, axis: x.unsqueeze(axis)
    tf.squeeze = lambda x, axis=None: x.squeeze() if axis is None else x.squeeze(axis)
    tf.stack = lambda xs, axis=0: torch.stack(list(xs), dim=axis)
    tf.concat = lambda xs, axis: torch.cat(list(xs), dim=axis)
    tf.cast = lambda x, dt: x.to(dt)
    tf.tile = lambda x, m: x.repeat(*_ints(m))
    tf.sigmoid = torch.sigmoid
    tf.range = lambda n, dtype=torch.int32: torch.arange(int(n), dtype=dtype)
    tf.reduce_mean = lambda x, axis=None: x.mean() if axis is None else x.mean(axis)
    tf.math = types.SimpleNamespace(sin=torch.sin, cos=torch.cos, rsqrt=torch.rsqrt, cumsum=lambda x, axis=0: torch.cumsum(x, dim=axis))
    tf.nn = types.SimpleNamespace(softmax=lambda x, axis=-1: torch.softmax(x, dim=axis))

    # ---- the ops of the loss path (detr_tf/loss/*.py, detr_tf/bbox.py), differentiable where TF's are
    _t = lambda x: x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
    tf.cast = lambda x, dt: _t(x).to(dt)
    tf.abs = lambda x: _t(x).abs()
    tf.clip_by_value = lambda x, lo, hi: _t(x).clamp(float(lo), float(hi))
    tf.reduce_sum = lambda x, axis=None: _t(x).sum() if axis is None else _t(x).sum(axis)
    tf.reduce_max = lambda x, axis=None: _t(x).max() if axis is None else _t(x).max(axis).values
    tf.argmax = lambda x, axis=None: _t(x).argmax(axis)
    tf.where = lambda c: torch.nonzero(_t(c))                         # single-argument form: coordinates of the true elements
    tf.constant = lambda v, dtype=None: torch.as_tensor(v, dtype=dtype)
    tf.gather = lambda params, indices, axis=0: torch.index_select(_t(params), axis, _t(indices).long().reshape(-1))

    def slice_(x, begin, size):
        idx = []
        for b, n in zip(begin, size):
            b, n = int(b), int(n)
            idx.append(slice(b, None) if n == -1 else slice(b, b + n))
        return x[tuple(idx)]
    tf.slice = slice_

    def norm(x, ord=2, axis=None):
        assert ord == 1
        return _t(x).abs().sum(axis)
    tf.norm = norm

    def numpy_function(fn, inp, Tout):
        res = fn(*[_t(i).detach().cpu().numpy() for i in inp])
        return [torch.as_tensor(np.asarray(r)).to(t) for r, t in zip(res, Tout)]
    tf.numpy_function = numpy_function
    tf.nn.relu = torch.relu
    tf.nn.sparse_softmax_cross_entropy_with_logits = lambda labels, logits: F.cross_entropy(logits, _t(labels).long(), reduction="none")
    tf.math.minimum, tf.math.maximum = torch.minimum, torch.maximum
    tf.math.abs, tf.math.log = torch.abs, torch.log
    tf.linalg = types.SimpleNamespace(diag_part=lambda x: torch.diagonal(x))
    if not hasattr(np, "bool"):                                       # hungarian_matching.py:37,41 uses the removed alias
        np.bool = bool

    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    for cls in (Layer, Conv2D, ZeroPadding2D, ReLU, Activation, MaxPool2D, Dropout, LayerNormalization, Dense):
        setattr(layers, cls.__name__, cls)
    keras.layers = layers
    keras.Model = Model
    keras.Input = lambda shape=None, **kw: STATE["input"]
    keras.initializers = types.SimpleNamespace(GlorotUniform=lambda *a, **k: None)
    keras.models = types.SimpleNamespace(Sequential=Sequential)
    keras.applications = types.SimpleNamespace(ResNet50=None)
    tf.keras = keras
    return tf, keras, layers


def install():
    tf, keras, layers = build()
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.keras"] = keras
    sys.modules["tensorflow.keras.layers"] = layers
    for name in ("matplotlib", "matplotlib.pyplot"):                  # imported at module top by bbox.py / detr.py, unused on this path
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot") and "matplotlib.pyplot" in sys.modules:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    return tf


__all__ = ["install", "set_input", "set_weights", "STATE", "np"]
