"""tf.train.Optimizer base-class subset: slots, compute_gradients (torch autograd), apply_gradients driving the
subclass's _create_slots/_prepare/_resource_apply_dense/_resource_apply_sparse_duplicate_indices/_finish the way
TF-1.14's apply_gradients does for resource variables (optimizer._DenseResourceVariableProcessor.update_op: an
IndexedSlices gradient goes to _resource_apply_sparse_duplicate_indices, a Tensor to _resource_apply_dense)."""
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
        gs = list(_torch.autograd.grad(loss, [v.t for v in vs], allow_unused=True, retain_graph=True))
        # which variables are read ONLY through embedding_lookup?  Probe with the gather backward blocked.
        _tf.state.block_gather = True
        try:
            other = _torch.autograd.grad(loss, [v.t for v in vs], allow_unused=True)
        finally:
            _tf.state.block_gather = False
        _tf.state.last_gradients = {v.op_name: (None if g is None else g.detach().clone()) for v, g in zip(vs, gs)}
        _tf.state.last_sparse = {}
        for i, v in enumerate(vs):
            recs = [r for r in _tf.state.gather_tape if r["var"] is v and "values" in r]
            if recs and other[i] is None:
                sl = _tf.IndexedSlices(_torch.cat([r["values"] for r in recs]), _torch.cat([r["ids"] for r in recs]),
                                       tuple(v.t.shape))
                gs[i] = sl
                _tf.state.last_sparse[v.op_name] = (sl.values.detach().clone(), sl.indices.clone())
        _tf.state.gather_tape = []
        return list(zip(gs, vs))

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        grads_and_vars = [(g, v) for g, v in grads_and_vars if g is not None]
        var_list = [v for _, v in grads_and_vars]
        with _torch.no_grad():
            _tf.run_update_ops()          # control_dependencies(update_ops) at models.py:196-197
            self._create_slots(var_list)
            self._prepare()
            update_ops = []
            for g, v in grads_and_vars:
                if isinstance(g, _tf.IndexedSlices):
                    update_ops.append(self._resource_apply_sparse_duplicate_indices(g.values.detach(), v, g.indices))
                else:
                    update_ops.append(self._resource_apply_dense(_tf._u(g).detach(), v))
            return self._finish(update_ops, name or self._name)
