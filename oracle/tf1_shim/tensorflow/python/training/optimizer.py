"""tf.train.Optimizer base-class subset: slots, compute_gradients (torch autograd), apply_gradients driving the
subclass's _create_slots/_prepare/_resource_apply_dense/_finish exactly like TF's apply_gradients does for
dense gradients of resource variables."""
import numpy as _np
import torch as _torch
import tensorflow as _tf


class Optimizer(object):
    def __init__(self, use_locking, name):
        self._use_locking = use_locking
        self._name = name
        self._slots = {}

    def _zeros_slot(self, var, slot_name, op_name):
        key = (var.op_name, slot_name)
        if key not in self._slots:
            full = "%s/%s/%s" % (var.op_name, op_name, slot_name)
            init = _tf.state.init_values.get(full, _np.zeros(tuple(var.t.shape)))
            v = _tf.Variable(full, init, trainable=False)
            _tf.state.variables[full] = v
            self._slots[key] = v
        return self._slots[key]

    def get_slot(self, var, name):
        return self._slots.get((var.op_name, name))

    def _call_if_callable(self, p):
        return p() if callable(p) else p

    def compute_gradients(self, loss):
        vs = _tf.trainable_variables()
        gs = _torch.autograd.grad(loss, [v.t for v in vs], allow_unused=True)
        _tf.state.last_gradients = {v.op_name: (None if g is None else g.detach().clone()) for v, g in zip(vs, gs)}
        return list(zip(gs, vs))

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        grads_and_vars = [(g, v) for g, v in grads_and_vars if g is not None]
        var_list = [v for _, v in grads_and_vars]
        with _torch.no_grad():
            _tf.run_update_ops()          # control_dependencies(update_ops) at models.py:196-197
            self._create_slots(var_list)
            self._prepare()
            update_ops = [self._resource_apply_dense(_tf._u(g).detach(), v) for g, v in grads_and_vars]
            return self._finish(update_ops, name or self._name)
