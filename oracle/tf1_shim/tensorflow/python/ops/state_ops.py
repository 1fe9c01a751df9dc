"""tf state ops on shim Variables: assign returns the variable (its new value is what later reads see)."""
import torch as _torch
import tensorflow as _tf


def assign(ref, value, use_locking=None):
    return ref.assign(value)


def assign_sub(ref, value, use_locking=None):
    with _torch.no_grad():
        ref.t.sub_(_tf._u(value))
    return ref


def scatter_add(ref, indices, updates, use_locking=None):
    with _torch.no_grad():
        ref.t.index_add_(0, _tf._u(indices).long(), _tf._u(updates))
    return ref
