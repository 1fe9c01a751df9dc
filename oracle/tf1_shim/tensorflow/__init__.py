"""TEST INFRASTRUCTURE — a minimal eager, torch-CPU emulation of the TensorFlow-1.14 API surface that the
reference's ``qa_cpg/models.py``, ``qa_cpg/utils/amsgrad.py`` and ``qa_cpg/metrics.py`` touch.

Purpose: TensorFlow 1.14 cannot be installed in this environment (no network; no Python-3.12 build), so the
reference cannot run as shipped.  With this package first on ``sys.path`` (``oracle/gen_golden.py`` does that),
the reference's OWN, UNMODIFIED source files import and execute: the graph wiring, op order, shapes,
reshape/flatten order, loss reduction, clipping and the optimizer's update code are the reference's; only the
numerical semantics of each individual TF op are restated here (from TF's documented behaviour — ASSUMED).
"Graph construction" executes eagerly, so building ``ConvE(...)`` runs one forward (+ one train step).

Never imported by the product (``coper_b200``); never shipped; not a TensorFlow replacement.
"""
import math as _math

import numpy as _np
import torch as _torch

float32 = _torch.float32
float64 = _torch.float64
int64 = _torch.int64
int32 = _torch.int32
string = "string"
AUTO_REUSE = "AUTO_REUSE"


class _State:
    """Everything a 'session.run' would feed, set by the caller before building the model."""
    def __init__(self):
        self.reset()

    def reset(self):
        self.dtype = _torch.float32
        self.batch = None              # dict e1,e2,rel,e2_multi,lookup_values (numpy)
        self.is_train = False
        self.init_values = {}          # variable name -> numpy initial value
        self.dropout_masks = []        # one entry (numpy keep-mask or None) per tf.nn.dropout call, in call order
        self.dropout_calls = 0
        self.variables = {}            # name -> Variable (trainable and not)
        self.trainable = []            # creation order
        self.update_ops = []
        self.rng = _np.random.default_rng(0)
        self.scope = []
        self.last_gradients = None
        self.gather_tape = []          # embedding_lookup records of trainable variables (for IndexedSlices grads)
        self.block_gather = False      # backward of embedding_lookup returns no gradient (usage probing)


state = _State()


class DType:
    def __init__(self, t):
        self.base_dtype = t
        self.t = t


class Variable:
    """A resource variable: a named torch leaf tensor with TF-style assign ops and arithmetic."""
    graph = "the-graph"

    def __init__(self, name, value, trainable=True):
        self.name = name + ":0"
        self.op_name = name
        self.t = _torch.tensor(_np.asarray(value), dtype=state.dtype)
        self.t.requires_grad_(bool(trainable))
        self.trainable = trainable
        self.handle = self
        self.dtype = DType(state.dtype)

    # -- TF variable API used by the reference
    def assign(self, value, use_locking=None):
        with _torch.no_grad():
            self.t.copy_(_u(value))
        return self

    def value(self):
        return self.t

    def get_shape(self):
        return tuple(self.t.shape)

    @property
    def shape(self):
        return tuple(self.t.shape)

    def numpy(self):
        return self.t.detach().numpy().copy()

    # -- arithmetic (returns plain tensors, like reading the variable)
    def __add__(self, o): return self.t + _u(o)
    def __radd__(self, o): return _u(o) + self.t
    def __sub__(self, o): return self.t - _u(o)
    def __rsub__(self, o): return _u(o) - self.t
    def __mul__(self, o): return self.t * _u(o)
    def __rmul__(self, o): return _u(o) * self.t
    def __truediv__(self, o): return self.t / _u(o)
    def __rtruediv__(self, o): return _u(o) / self.t
    def __neg__(self): return -self.t
    def __getitem__(self, k): return self.t[k]


def _u(x):
    """unwrap Variables / python scalars / numpy into torch tensors (or python numbers)."""
    if isinstance(x, Variable):
        return x.t
    if isinstance(x, _np.ndarray):
        return _torch.as_tensor(x)
    return x


def _full_name(name):
    return name


# ---------------------------------------------------------------- graph plumbing (no-ops eagerly)
class _Ctx:
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


def device(*a, **k): return _Ctx()
def variable_scope(*a, **k): return _Ctx()
def name_scope(*a, **k): return _Ctx()
def control_dependencies(*a, **k): return _Ctx()


class GraphKeys:
    UPDATE_OPS = "update_ops"


def get_collection(key):
    return list(state.update_ops) if key == GraphKeys.UPDATE_OPS else []


def placeholder(dtype, shape=None, name=None):
    return "placeholder:" + str(name)


def placeholder_with_default(default, shape=None, name=None):
    if name == "is_train":
        return bool(state.is_train)
    return default


class _Iterator:
    def __init__(self, output_types):
        self.types = output_types

    def get_next(self):
        b = state.batch
        out = {}
        for k, t in self.types.items():
            v = b[k]
            out[k] = _torch.as_tensor(_np.asarray(v)).to(state.dtype if t == float32 else t)
        return out


class _IteratorNS:
    @staticmethod
    def from_string_handle(handle, output_types, output_shapes=None):
        return _Iterator(output_types)


class data:
    Iterator = _IteratorNS


# ---------------------------------------------------------------- variables / initialisers
class _Xavier:
    def __call__(self, shape):
        shape = tuple(int(s) for s in shape)
        if len(shape) == 1:
            fi = fo = shape[0]
        else:
            rec = int(_np.prod(shape[:-2])) if len(shape) > 2 else 1
            fi, fo = shape[-2] * rec, shape[-1] * rec
        lim = _math.sqrt(6.0 / (fi + fo))
        return state.rng.uniform(-lim, lim, size=shape)


class _ContribLayers:
    @staticmethod
    def xavier_initializer():
        return _Xavier()


class contrib:
    layers = _ContribLayers


def zeros_initializer():
    return lambda shape: _np.zeros(tuple(int(s) for s in shape))


def ones_initializer():
    return lambda shape: _np.ones(tuple(int(s) for s in shape))


def get_variable(name, dtype=None, shape=None, initializer=None, trainable=True):
    if name in state.variables:
        return state.variables[name]
    if name in state.init_values:
        val = _np.asarray(state.init_values[name]).reshape(tuple(int(s) for s in shape))
    else:
        val = initializer(shape)
    v = Variable(name, val, trainable)
    state.variables[name] = v
    if trainable:
        state.trainable.append(v)
    return v


def trainable_variables():
    return list(state.trainable)


# ---------------------------------------------------------------- math
def cast(x, dtype):
    x = _u(x)
    if isinstance(dtype, DType):
        dtype = dtype.t
    if dtype == float32:
        dtype = state.dtype            # the whole emulation runs in state.dtype (fp32 or fp64)
    if isinstance(x, (bool, int, float)):
        return _torch.tensor(float(x) if dtype in (_torch.float32, _torch.float64) else x, dtype=dtype)
    return x.to(dtype)


def matmul(a, b, name=None):
    return _torch.matmul(_u(a), _u(b))


def reshape(x, shape):
    shape = [int(s) if not isinstance(s, _torch.Tensor) else int(s.item()) for s in shape]
    return _u(x).reshape(shape)


def transpose(x, perm=None):
    x = _u(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*perm)


def concat(values, axis):
    return _torch.cat([_u(v) for v in values], dim=axis)


def gather(params, indices):
    return nn.embedding_lookup(params, indices)        # same ResourceGather op / IndexedSlices gradient in TF


def shape(x):
    return list(_u(x).shape)


def zeros(shape, dtype=None):
    return _torch.zeros(tuple(shape), dtype=state.dtype)


def sqrt(x): return _torch.sqrt(_u(x))
def square(x): return _u(x) ** 2
def maximum(a, b): return _torch.maximum(_u(a), _u(b))
def reduce_mean(x): return _u(x).mean()
def reduce_max(x): return _u(x).max()
def reduce_min(x): return _u(x).min()


def reduce_sum(x, name=None):
    return _u(x).sum()


def map_fn(fn, elems):
    outs = [fn(tuple(e[i] for e in elems)) for i in range(elems[0].shape[0])]
    return tuple(_torch.stack([o[j] for o in outs]) for j in range(len(outs[0])))


class IndexedSlices:
    """tf.IndexedSlices: what TF-1.14 returns as the gradient of a variable that is only read through
    tf.nn.embedding_lookup / tf.gather (ResourceGather's gradient).  A variable that ALSO has a dense gradient
    gets the sum as a dense tensor (gradients_util._AggregateIndexedSlicesGradients: "if any gradient is a Tensor,
    add_n them")."""

    def __init__(self, values, indices, dense_shape=None):
        self.values, self.indices, self.dense_shape = values, indices, dense_shape

    def to_dense(self):
        out = _torch.zeros(self.dense_shape, dtype=self.values.dtype)
        return out.index_add_(0, self.indices, self.values)


def clip_by_global_norm(t_list, clip_norm):
    """tf.clip_by_global_norm: t * clip_norm / max(global_norm, clip_norm); None entries pass through.
    For an IndexedSlices the norm is taken over its `values` (clip_ops.global_norm uses l2_loss(t.values)):
    slices that share an index are NOT summed first."""
    def vals(t):
        return t.values if isinstance(t, IndexedSlices) else _u(t)
    norm = _torch.sqrt(sum((vals(t) ** 2).sum() for t in t_list if t is not None))
    scale = clip_norm / _torch.maximum(norm, _torch.tensor(float(clip_norm), dtype=norm.dtype))
    out = []
    for t in t_list:
        if t is None:
            out.append(None)
        elif isinstance(t, IndexedSlices):
            out.append(IndexedSlices(t.values * scale, t.indices, t.dense_shape))
        else:
            out.append(_u(t) * scale)
    return out, norm


class _GatherFn(_torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, ids, rec):
        ctx.rec, ctx.shape = rec, params.shape
        ctx.save_for_backward(ids)
        return params[ids]

    @staticmethod
    def backward(ctx, go):
        (ids,) = ctx.saved_tensors
        flat = go.reshape(-1, *ctx.shape[1:])
        ctx.rec["values"] = flat.detach().clone()
        if state.block_gather:
            return None, None, None
        dense = _torch.zeros(ctx.shape, dtype=go.dtype).index_add_(0, ids.reshape(-1), flat)
        return dense, None, None


class nn:
    @staticmethod
    def embedding_lookup(params, ids, name=None):
        ids = _u(ids).long()
        if isinstance(params, Variable) and params.trainable:
            rec = {"var": params, "ids": ids.reshape(-1)}
            state.gather_tape.append(rec)
            return _GatherFn.apply(params.t, ids, rec)
        return _u(params)[ids]

    @staticmethod
    def relu(x):
        return _torch.relu(_u(x))

    @staticmethod
    def conv2d(input, filter, strides, padding):
        assert padding == "VALID" and list(strides) == [1, 1, 1, 1]
        x = _u(input).permute(0, 3, 1, 2)                       # NHWC -> NCHW
        w = _u(filter).permute(3, 2, 0, 1)                      # HWIO -> OIHW
        return _torch.nn.functional.conv2d(x, w).permute(0, 2, 3, 1)

    @staticmethod
    def dropout(x, keep_prob):
        """x * mask / keep_prob; the Bernoulli(keep) mask of call #i comes from state.dropout_masks[i]."""
        i = state.dropout_calls
        state.dropout_calls += 1
        keep = float(_u(keep_prob))
        if keep >= 1.0:
            return _u(x)
        mask = state.dropout_masks[i]
        return _u(x) * _torch.as_tensor(_np.asarray(mask)).to(state.dtype).reshape(_u(x).shape) / keep


class layers:
    @staticmethod
    def batch_normalization(x, momentum=0.99, epsilon=1e-3, reuse=None, training=False, fused=None, name=None):
        """axis=-1; gamma=1, beta=0, moving_mean=0, moving_variance=1 at init; training -> batch mean and BIASED
        variance normalise, moving <- moving*momentum + batch*(1-momentum) (Bessel-corrected variance on the
        fused 4-D path only)."""
        x = _u(x)
        C = x.shape[-1]
        g = get_variable(name + "/gamma", shape=[C], initializer=ones_initializer())
        b = get_variable(name + "/beta", shape=[C], initializer=zeros_initializer())
        mm = get_variable(name + "/moving_mean", shape=[C], initializer=zeros_initializer(), trainable=False)
        mv = get_variable(name + "/moving_variance", shape=[C], initializer=ones_initializer(), trainable=False)
        if bool(training):
            red = tuple(range(x.dim() - 1))
            mean = x.mean(dim=red)
            var = ((x - mean) ** 2).mean(dim=red)
            n = x.numel() // C
            fused_path = bool(fused) and x.dim() == 4
            with _torch.no_grad():
                vm = var * (n / max(n - 1, 1)) if fused_path else var
                new_mm = mm.t * momentum + mean * (1 - momentum)
                new_mv = mv.t * momentum + vm * (1 - momentum)
            state.update_ops.append((mm, new_mm.detach(), mv, new_mv.detach()))
        else:
            mean, var = mm.t, mv.t
        inv = _torch.rsqrt(var + epsilon)
        return (x - mean) * inv * g.t + b.t


class losses:
    @staticmethod
    def sigmoid_cross_entropy(multi_class_labels, logits):
        """mean over all elements of max(s,0) - s*z + log1p(exp(-|s|)) (SUM_BY_NONZERO_WEIGHTS, weight 1)."""
        s, z = _u(logits), _u(multi_class_labels)
        el = _torch.clamp(s, min=0) - s * z + _torch.log1p(_torch.exp(-_torch.abs(s)))
        return el.sum() / el.numel()


class summary:
    @staticmethod
    def scalar(*a, **k): return None
    @staticmethod
    def histogram(*a, **k): return None
    @staticmethod
    def merge_all(): return None


class errors:
    class OutOfRangeError(Exception):
        pass


def run_update_ops():
    """Apply the BN moving-average updates collected in UPDATE_OPS (they run with train_op, models.py:196-197)."""
    for mm, new_mm, mv, new_mv in state.update_ops:
        mm.assign(new_mm)
        mv.assign(new_mv)
    state.update_ops = []
