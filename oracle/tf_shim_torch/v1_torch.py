"""torch-backed stand-in for `tensorflow.compat.v1`, used ONLY to obtain the gradients of the reference's own
loss (Main_Functions.py:337-378) by automatic differentiation.  TEST INFRASTRUCTURE ONLY (oracle).

`oracle/ref_grad.py` loads a private copy of /root/reference/Main_Functions.py and rebinds its global `tf` to
this module, so `build_neural_network` runs UNMODIFIED on torch tensors and torch.autograd differentiates it.
The per-op gradient rules that matter agree between TF 2.4 and torch:
  * reduce_min / amin: the gradient is split evenly among tied minima;
  * clip_by_value / clamp: passes where min <= x <= max (inclusive);
  * abs: sign(x) (0 at 0);   sign, round, comparisons: no gradient;   stop_gradient = detach.
`tf.train.AdamOptimizer(...).minimize(loss, var_list)` only records (loss, var_list): the caller differentiates.
"""
import numpy as np
import torch

float32 = torch.float32
int64 = torch.int64
_CONST = {}


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        key = (id(x), x.shape)
        if key not in _CONST or _CONST[key][0] is not x:
            _CONST[key] = (x, torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)))
        return _CONST[key][1]
    return torch.as_tensor(x, dtype=torch.float32)


def disable_v2_behavior():
    return None


def to_float(x):
    return _t(x).to(torch.float32)


def transpose(x, perm=None):
    return _t(x).permute(*perm)


def multiply(a, b):
    return _t(a) * _t(b)


def add(a, b):
    return _t(a) + _t(b)


def reshape(x, shape, name=None):
    return _t(x).reshape(*[int(s) for s in shape])


def matmul(a, b):
    return torch.matmul(_t(a), _t(b))


def tile(x, multiples):
    return _t(x).repeat(*[int(m) for m in multiples])


def abs(x):  # noqa: A001
    return torch.abs(_t(x))


def sign(x):
    return torch.sign(_t(x)).detach()      # tf.sign has no gradient


def exp(x):
    return torch.exp(_t(x))


def reduce_prod(x, axis=None, reduction_indices=None):
    ax = axis if axis is not None else reduction_indices
    return torch.prod(_t(x), dim=ax)


def reduce_min(x, axis=None):
    return torch.amin(_t(x), dim=axis)


def reduce_mean(x, name=None):
    return torch.mean(_t(x))


def zeros(shape, dtype=torch.float32):
    return torch.zeros(*[int(s) for s in shape], dtype=torch.float32)


def ones(shape, dtype=torch.float32):
    return torch.ones(*[int(s) for s in shape], dtype=torch.float32)


def clip_by_value(x, clip_value_min, clip_value_max):
    return torch.clamp(_t(x), float(np.float32(clip_value_min)), float(np.float32(clip_value_max)))


def round(x):  # noqa: A001  (half to even, like tf.round; only ever used under stop_gradient)
    return torch.round(_t(x)).detach()


def stop_gradient(x):
    return _t(x).detach()


def concat(values, axis):
    return torch.cat([_t(v) for v in values], dim=axis)


def tanh(x):
    return torch.tanh(_t(x))


def atanh(x):
    return torch.atanh(_t(x))


class nn:  # noqa: N801
    @staticmethod
    def sigmoid_cross_entropy_with_logits(labels=None, logits=None):
        # TF's implementation (nn_impl.py): where(x >= 0, x, 0) - x * z + log1p(exp(where(x >= 0, -x, x))); with the
        # selects (not clamp / abs) the gradient at x == 0 is sigmoid(0) - z, and quantised APPs are often exactly 0
        x, z = _t(logits), _t(labels).to(torch.float32)
        cond = x >= 0
        zero = torch.zeros_like(x)
        return torch.where(cond, x, zero) - x * z + torch.log1p(torch.exp(torch.where(cond, -x, x)))


class math:  # noqa: N801
    @staticmethod
    def sigmoid(x):
        return torch.sigmoid(_t(x))


class _Adam:
    last = None

    def __init__(self, learning_rate=None):
        self.learning_rate = learning_rate

    def minimize(self, loss, var_list=None):
        _Adam.last = (loss, list(var_list))
        return ("train_step", loss, list(var_list))


class train:  # noqa: N801
    AdamOptimizer = _Adam
