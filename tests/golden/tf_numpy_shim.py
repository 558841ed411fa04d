"""A numpy-backed stand-in for the handful of `tensorflow` symbols that the reference's
loss path touches (detr_tf/loss/loss.py, detr_tf/loss/hungarian_matching.py, detr_tf/bbox.py).

TensorFlow is not installable in the build container (no wheel, no network), so the
reference's *model* cannot be run.  Its loss/matcher code however is ~40 elementwise /
gather / reduce ops with unambiguous semantics; with this shim installed as
``sys.modules['tensorflow']`` the reference files are imported from /root/reference and
executed UNMODIFIED to produce the golden vectors under tests/golden/ (see make_golden.py).

Test infrastructure only.  float32 in -> float32 arithmetic, like TF eager on CPU.
"""
import sys
import types

import numpy as np


def _np(x):
    return np.asarray(x)


def build():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64 = np.float32, np.float64
    tf.int32, tf.int64, tf.bool = np.int32, np.int64, np.bool_
    tf.Tensor = np.ndarray
    tf.newaxis = None

    tf.shape = lambda x: np.array(_np(x).shape, dtype=np.int32)
    tf.cast = lambda x, dt: _np(x).astype(dt)
    tf.zeros = lambda shape, dtype=np.float32: np.zeros(tuple(int(s) for s in np.atleast_1d(shape)), dtype=dtype)
    tf.zeros_like = lambda x: np.zeros_like(_np(x))
    tf.expand_dims = lambda x, axis: np.expand_dims(_np(x), axis)
    tf.squeeze = lambda x, axis=None: np.squeeze(_np(x), axis=axis)
    tf.tile = lambda x, m: np.tile(_np(x), tuple(int(v) for v in m))
    tf.concat = lambda xs, axis: np.concatenate([_np(x) for x in xs], axis=axis)
    tf.clip_by_value = lambda x, lo, hi: np.clip(_np(x), np.float32(lo), np.float32(hi))
    tf.abs = lambda x: np.abs(_np(x))
    tf.reduce_sum = lambda x, axis=None: np.sum(_np(x), axis=axis, dtype=_np(x).dtype)
    tf.reduce_mean = lambda x, axis=None: np.mean(_np(x), axis=axis, dtype=_np(x).dtype)
    tf.argmax = lambda x, axis=None: np.argmax(_np(x), axis=axis).astype(np.int64)
    tf.where = lambda c: np.argwhere(_np(c)).astype(np.int64)
    tf.reduce_max = lambda x, axis=None: np.max(_np(x), axis=axis)
    tf.constant = lambda v, dtype=None: np.array(v, dtype=dtype)

    def pad(x, paddings, mode="CONSTANT", constant_values=0):
        assert mode == "CONSTANT"
        pw = [(int(a), int(b)) for a, b in paddings]
        assert all(a >= 0 and b >= 0 for a, b in pw), "tf.pad rejects negative paddings"
        return np.pad(_np(x), pw, mode="constant", constant_values=constant_values)
    tf.pad = pad

    def slice_(x, begin, size):
        x = _np(x)
        idx = []
        for d, (b, s) in enumerate(zip(begin, size)):
            b, s = int(b), int(s)
            idx.append(slice(b, None) if s == -1 else slice(b, b + s))
        return x[tuple(idx)]
    tf.slice = slice_

    def gather(params, indices, axis=0):
        return np.take(_np(params), _np(indices).astype(np.int64), axis=axis)
    tf.gather = gather

    def norm(x, ord=2, axis=None):
        assert ord == 1
        return np.sum(np.abs(_np(x)), axis=axis, dtype=_np(x).dtype)
    tf.norm = norm

    def numpy_function(fn, inp, Tout):
        res = fn(*[_np(i) for i in inp])
        return [np.asarray(r).astype(t) for r, t in zip(res, Tout)]
    tf.numpy_function = numpy_function

    def softmax(x, axis=-1):
        x = _np(x)
        e = np.exp(x - np.max(x, axis=axis, keepdims=True))
        return (e / np.sum(e, axis=axis, keepdims=True, dtype=x.dtype)).astype(x.dtype)

    def sparse_ce(labels, logits):
        logits = _np(logits)
        labels = _np(labels).astype(np.int64)
        m = np.max(logits, axis=-1, keepdims=True)
        lse = np.log(np.sum(np.exp(logits - m), axis=-1, dtype=logits.dtype)) + m[..., 0]
        return (lse - np.take_along_axis(logits, labels[:, None], axis=-1)[:, 0]).astype(logits.dtype)

    nn = types.ModuleType("tensorflow.nn")
    nn.relu = lambda x: np.maximum(_np(x), 0)
    nn.softmax = softmax
    nn.sparse_softmax_cross_entropy_with_logits = sparse_ce
    tf.nn = nn

    math = types.ModuleType("tensorflow.math")
    math.minimum = lambda a, b: np.minimum(_np(a), _np(b))
    math.maximum = lambda a, b: np.maximum(_np(a), _np(b))
    math.abs = tf.abs
    math.log = lambda x: np.log(_np(x))
    tf.math = math

    linalg = types.ModuleType("tensorflow.linalg")
    linalg.diag_part = lambda x: np.diagonal(_np(x)).copy()
    tf.linalg = linalg
    return tf


def install():
    """Install the shim + stubs for the plotting imports of bbox.py (matplotlib is missing)."""
    tf = build()
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.nn"] = tf.nn
    sys.modules["tensorflow.math"] = tf.math
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if not hasattr(np, "bool"):      # hungarian_matching.py:37,41 uses the removed alias
        np.bool = bool
    return tf
