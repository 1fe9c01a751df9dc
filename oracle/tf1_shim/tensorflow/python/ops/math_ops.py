import torch as _torch
import tensorflow as _tf


def cast(x, dtype):
    return _tf.cast(x, dtype)


def sqrt(x):
    return _torch.sqrt(_tf._u(x))
