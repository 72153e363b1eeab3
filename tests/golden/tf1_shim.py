"""A minimal EAGER stand-in for the TensorFlow-1 API surface that the reference's TCAR path touches, backed by
torch float64 + autograd.  It exists for ONE purpose: to execute the reference's own graph-building code
(/root/reference/model_combine.py, modules.py, util.py) in this TF-less container so that golden vectors can be
generated from the reference source itself (tests/golden/make_golden.py).  It is test tooling, not product code.

How it works: `tf.placeholder(name=...)` returns the value registered in FEED under that name, so constructing
`Seq2SeqAttNN(args)` evaluates the whole graph eagerly, including `compute_gradients` (torch.autograd),
`clip_by_norm` and `AdamOptimizer.apply_gradients` (one optimiser step).  Semantics of the TF kernels follow the
TF 1.x documentation:
  embedding_lookup(max_norm)  -> clip_by_norm over the embedding axis (zero rows pass through)
  sparse_softmax_cross_entropy_with_logits -> logsumexp(logits) - logits[label]
  clip_by_norm(t, c)          -> t * c / max(||t||, c); for an IndexedSlices gradient ||t|| is the norm of the
                                 UN-AGGREGATED slice values (clip_ops.clip_by_norm reads t.values): the gradient of a
                                 variable that is only ever read through embedding_lookup (dec_pos, month, day, week,
                                 hour, minute, duration) is the concatenation of one slice per lookup, duplicates not
                                 summed (gradients_util._AggregatedGrads); item_emb also feeds the dense `[1:]` slice
                                 of the candidate matrix, so its gradient is aggregated to a dense tensor (TF >= 1.14
                                 backprop.aggregate_indexed_slices_gradients: "if any gradient is a Tensor, add_n").
                                 The optimizer sums duplicate slices before the update (_apply_sparse_duplicate_indices)
                                 so only the clip FACTOR differs from the dense reading.
  AdamOptimizer               -> lr_t = lr sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
                                 var -= lr_t m / (sqrt(v) + eps)      (training_ops ApplyAdam)
"""
import contextlib
import math
import sys
import types

import numpy as np
import torch

DT = torch.float64
FEED = {}            # placeholder name -> numpy array / python value
STATE = {"vars": [], "trainable": [], "grads": None, "init": {}, "adam": None, "rng": None, "rn_calls": 0,
         "lookups": {}, "dense_use": set(), "slices_sq": {}}

bool = "bool"        # noqa: A001  (tf.bool)
int32 = "int32"
float32 = "float32"


def reset(seed=2020):
    FEED.clear()
    STATE.update(vars=[], trainable=[], grads=None, init={}, adam=None, rng=seed, rn_calls=0, lookups={},
                 dense_use=set(), slices_sq={})


def rn_values(seed, call, shape, stddev=1.0, mean=0.0):
    """The k-th tf.random_normal draw of a graph built under `seed`: legacy NumPy RandomState, so that tests can
    regenerate large weight matrices instead of storing them (TF's own stream is not reproducible outside TF)."""
    return np.random.RandomState(seed * 1000 + call).normal(mean, stddev, [int(s) for s in shape])


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    a = np.asarray(x)
    if a.dtype.kind in "iub":
        return torch.tensor(a, dtype=torch.long)
    return torch.tensor(a, dtype=DT)


torch.Tensor.get_shape = lambda self: tuple(self.shape)   # used only by the reference's debug prints


def Variable(initial_value, dtype=None, trainable=True, name=None):
    if callable(initial_value):
        initial_value = initial_value()
    v = _t(initial_value).clone()
    if dtype == float32 or (dtype is None and v.dtype == DT):
        v = v.to(torch.float32).to(DT)        # TF variables are float32; values are rounded, maths stays float64
    if dtype == int32:
        v = v.long()
    v.tf_name = name
    if trainable and v.dtype == DT:
        v.requires_grad_(True)
        STATE["trainable"].append(v)
        STATE["init"][id(v)] = v.detach().clone()
    STATE["vars"].append(v)
    return v


def placeholder(dtype, shape=None, name=None):
    if name not in FEED:
        raise KeyError("placeholder %r has no value in FEED" % name)
    return _t(FEED[name]) if dtype != bool else FEED[name]


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    yield


def random_normal(shape, stddev=1.0, mean=0.0):
    r = torch.tensor(rn_values(STATE["rng"], STATE["rn_calls"], shape, stddev, mean), dtype=DT)
    STATE["rn_calls"] += 1
    return r.to(torch.float32).to(DT)


def set_random_seed(seed):
    STATE["rng"] = seed


def trainable_variables():
    return list(STATE["trainable"])


def shape(x):
    return list(x.shape)


def reshape(x, shp):
    return x.reshape([int(s) for s in shp])


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def range(n):                      # noqa: A001  (tf.range)
    return torch.arange(int(n))


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def matmul(a, b, transpose_b=False):
    return torch.matmul(a, b.transpose(-1, -2) if transpose_b else b)


def reduce_sum(x, axis=None, keep_dims=False, keepdims=False):
    if axis is None:
        return x.sum()
    return x.sum(dim=axis, keepdim=keep_dims or keepdims)


def exp(x):
    return torch.exp(x)


def log(x):
    return torch.log(x)


def sigmoid(x):
    return torch.sigmoid(x)


class SlicedGrad:
    """Stand-in for a tf.IndexedSlices gradient: `dense` = the slices summed into the variable's shape (what the
    optimizer applies), `values_sq` = sum of squares of the un-aggregated slice values (what clip_by_norm norms)."""

    def __init__(self, dense, values_sq):
        self.dense, self.values_sq = dense, values_sq


def clip_by_norm(t, clip_norm, axes=None):
    if isinstance(t, SlicedGrad):
        n = torch.sqrt(torch.as_tensor(t.values_sq, dtype=DT))
        return t.dense * clip_norm / torch.clamp(n, min=float(clip_norm))
    if axes is None:
        n = torch.sqrt((t * t).sum())
        return t * clip_norm / torch.clamp(n, min=float(clip_norm))
    sq = (t * t).sum(dim=axes, keepdim=True)
    safe = torch.where(sq > 0, sq, torch.ones_like(sq))
    norm = torch.where(sq > 0, safe.sqrt(), sq)
    return (t * clip_norm) / torch.clamp(norm, min=float(clip_norm))


class _NN(types.SimpleNamespace):
    @staticmethod
    def embedding_lookup(params, ids, max_norm=None):
        rows = params[_t(ids).long()]
        if any(params is v for v in STATE["trainable"]):
            # one IndexedSlices gradient per lookup
            STATE["lookups"].setdefault(id(params), []).append((rows, _t(ids).long()))
        return clip_by_norm(rows, max_norm, axes=-1) if max_norm is not None else rows

    tanh = staticmethod(torch.tanh)
    sigmoid = staticmethod(torch.sigmoid)
    relu = staticmethod(torch.relu)

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(logits=None, labels=None):
        lab = _t(labels).long()
        return torch.logsumexp(logits, dim=-1) - logits.gather(-1, lab[:, None]).squeeze(-1)


nn = _NN()


class _Adam:
    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = learning_rate, beta1, beta2, epsilon, 0
        self.m, self.v = {}, {}
        STATE["adam"] = self

    def compute_gradients(self, loss, var_list):
        taps = [r for v in var_list for r, _ in STATE["lookups"].get(id(v), [])]
        all_g = torch.autograd.grad(loss.sum(), list(var_list) + taps, allow_unused=True)
        grads, tap_g = all_g[: len(var_list)], list(all_g[len(var_list):])
        STATE["grads"] = [None if g is None else g.detach().clone() for g in grads]
        out = []
        for g, v in zip(STATE["grads"], var_list):
            looks = STATE["lookups"].get(id(v), [])
            sq, recon = 0.0, torch.zeros_like(v)
            for _, ids in looks:
                tg = tap_g.pop(0)
                if tg is not None:
                    sq += float((tg.detach() ** 2).sum())
                    recon.index_add_(0, ids.reshape(-1), tg.detach().reshape(-1, v.shape[1]))
            # a variable read ONLY through embedding_lookup gets an IndexedSlices gradient (slices concatenated, not
            # summed); any other use (item_emb[1:] in the candidate matrix) shows up as gradient the lookups do not
            # explain, and makes TF aggregate to a dense tensor
            sliced = len(looks) > 0 and g is not None and float((g - recon).abs().max()) <= 1e-12 * (1 + float(g.abs().max()))
            STATE["slices_sq"][id(v)] = sq if sliced else None
            out.append((SlicedGrad(g, sq) if sliced else g, v))
        return out

    def apply_gradients(self, grads_and_vars, global_step=None):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        grads_and_vars = [(g.dense if isinstance(g, SlicedGrad) else g, v) for g, v in grads_and_vars]
        STATE["capped"] = [g.detach().clone() for g, _ in grads_and_vars]
        with torch.no_grad():
            for g, v in grads_and_vars:
                m = self.m.setdefault(id(v), torch.zeros_like(v))
                s = self.v.setdefault(id(v), torch.zeros_like(v))
                m += (g - m) * (1 - self.b1)
                s += (g * g - s) * (1 - self.b2)
                v -= lr_t * m / (s.sqrt() + self.eps)
            if global_step is not None:
                global_step += 1
        return "train_op"


train = types.SimpleNamespace(AdamOptimizer=_Adam)


def install():
    """Register this module as `tensorflow` so `import tensorflow as tf` in the reference resolves to it."""
    sys.modules["tensorflow"] = sys.modules[__name__]
