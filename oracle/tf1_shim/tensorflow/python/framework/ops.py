"""tensorflow.python.framework.ops subset used by utils/amsgrad.py."""
import torch as _torch
import tensorflow as _tf


def control_dependencies(*a, **k): return _tf._Ctx()
def colocate_with(*a, **k): return _tf._Ctx()
def name_scope(*a, **k): return _tf._Ctx()


def convert_to_tensor(x, name=None):
    if callable(x):
        x = x()
    x = _tf._u(x)
    return x if isinstance(x, _torch.Tensor) else _torch.tensor(float(x), dtype=_tf.state.dtype)


class Tensor:  # isinstance checks only
    pass
