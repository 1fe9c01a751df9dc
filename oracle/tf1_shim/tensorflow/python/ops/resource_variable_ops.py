"""resource_variable_ops subset: scatter-add into a shim Variable (duplicate indices accumulate)."""
import torch as _torch
import tensorflow as _tf


def resource_scatter_add(handle, indices, updates):
    with _torch.no_grad():
        handle.t.index_add_(0, _tf._u(indices).long(), _tf._u(updates))
    return handle
