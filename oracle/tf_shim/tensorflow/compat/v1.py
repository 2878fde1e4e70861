"""Eager numpy implementation of the `tensorflow.compat.v1` ops that
`/root/reference/Main_Functions.py:157-335, 463-494` calls on the decode path.

TEST INFRASTRUCTURE ONLY (oracle).  Semantics that matter and match TF 2.4:
  * everything is float32 (`to_float`, `zeros`, `ones`); python scalars are weak
    (NumPy >= 2 promotion) so `0.0001 * f32_tensor` stays float32 like TF;
  * `round` is round-half-to-even (`np.rint`), as `tf.round`;
  * `sign(0) == 0`;
  * `matmul` is a float32 GEMM (the connection matrices are 0/1).
Graph-mode only entry points (`placeholder`, `Session`, `get_variable`, the
optimizer) are intentionally absent: the oracle never runs the loss/optimizer
block (`sampling_type == 2` skips it, Main_Functions.py:338).
"""
import numpy as np

float32 = np.float32
int64 = np.int64


def disable_v2_behavior():
    return None


def _f(x):
    return np.asarray(x, dtype=np.float32)


def to_float(x):
    return np.asarray(x).astype(np.float32)


def transpose(x, perm=None):
    return np.transpose(x, perm)


def multiply(a, b):
    return np.multiply(a, b)


def add(a, b):
    return np.add(a, b)


def reshape(x, shape, name=None):
    return np.reshape(x, shape)


def matmul(a, b):
    return np.matmul(_f(a), _f(b))


def tile(x, multiples):
    return np.tile(x, multiples)


def abs(x):  # noqa: A001 - mirrors tf.abs
    return np.abs(x)


def sign(x):
    return np.sign(x)


def reduce_prod(x, axis=None, reduction_indices=None):
    ax = axis if axis is not None else reduction_indices
    return np.prod(x, axis=ax, dtype=np.float32)


def reduce_min(x, axis=None):
    return np.min(x, axis=axis)


def zeros(shape, dtype=np.float32):
    return np.zeros(shape, dtype=dtype)


def ones(shape, dtype=np.float32):
    return np.ones(shape, dtype=dtype)


def clip_by_value(x, clip_value_min, clip_value_max):
    return np.clip(x, np.float32(clip_value_min), np.float32(clip_value_max))


def round(x):  # noqa: A001 - mirrors tf.round (half to even)
    return np.rint(x)


def stop_gradient(x):
    return x


def concat(values, axis):
    return np.concatenate(values, axis=axis)


def tanh(x):
    return np.tanh(x)


def atanh(x):
    return np.arctanh(x)
