"""Mint golden vectors from the REFERENCE's own source files (test infrastructure; run in the build container,
where /root/reference exists — the GPU box never needs it, it reads the committed .npz files).

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz
    python oracle/gen_golden.py concat_train concat_eval      # only the named model cases

What runs:
  * ``/root/reference/CoPER_ConvE/qa_cpg/metrics.py`` — UNMODIFIED, with real NumPy; the fake ``tensorflow``
    only provides ``tf.errors.OutOfRangeError`` and the ``session`` is a stub that replays fixed
    (e1, e2, rel, e2_multi, pred) batches.                              -> tests/golden/ref_metrics_*.npz
  * ``/root/reference/CoPER_ConvE/qa_cpg/models.py`` + ``utils/amsgrad.py`` — UNMODIFIED, on ``oracle/tf1_shim``
    (torch-CPU emulation of the TF-1 API; TF op semantics restated/assumed, graph wiring = the reference's).
    Building ``ConvE(model_descriptors)`` executes forward, loss, compute_gradients, clip_by_global_norm(5.0) and
    the AMSGrad apply once.                                              -> tests/golden/ref_model_*.npz
The reference is imported straight from /root/reference (nothing is copied into this repo).
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/CoPER_ConvE"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "tf1_shim"))
    sys.path.insert(0, REF)
    import tensorflow as tf                     # the shim
    assert "tf1_shim" in tf.__file__
    models = importlib.import_module("qa_cpg.models")
    metrics = importlib.import_module("qa_cpg.metrics")
    assert models.__file__.startswith(REF) and metrics.__file__.startswith(REF)
    return tf, models, metrics


# ------------------------------------------------------------------------------------------------ metrics.py
class _ReplaySession:
    def __init__(self, tf, batches):
        self.tf, self.batches, self.i = tf, batches, 0

    def run(self, fetches, feed_dict=None):
        if self.i >= len(self.batches):
            raise self.tf.errors.OutOfRangeError()
        b = self.batches[self.i]
        self.i += 1
        return (b["e1"], b["e2"], b["rel"], b["e2_multi"].copy(), b["pred"].copy())


class _ModelStub:
    e1 = e2 = rel = e2_multi = predictions_all = input_iterator_handle = None


def gen_metrics(tf, metrics, name, B_list, N, seed, filt_p):
    rng = np.random.default_rng(seed)
    batches = []
    for B in B_list:
        pred = rng.normal(size=(B, N)).astype(np.float32)
        e2 = rng.integers(0, N, B).astype(np.int64)
        filt = (rng.random((B, N)) < filt_p).astype(np.float32)
        filt[np.arange(B), e2] = 1.0               # the gold tail is always in the known-true set (data.py:494)
        batches.append({"e1": np.zeros(B, np.int64), "e2": e2, "rel": np.zeros(B, np.int64), "e2_multi": filt,
                        "pred": pred})
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        mr, mrr, hits = metrics.ranking_and_hits(_ModelStub(), td, "handle", name, _ReplaySession(tf, batches))
        # per-query ranks: a one-query evaluation's mean rank IS that query's rank
        ranks = []
        for b in batches:
            for i in range(len(b["e2"])):
                one = {k: v[i:i + 1] for k, v in b.items()}
                r, _, _ = metrics.ranking_and_hits(_ModelStub(), td, "handle", name, _ReplaySession(tf, [one]))
                ranks.append(int(r))
    out = {"pred": np.concatenate([b["pred"] for b in batches]), "e2": np.concatenate([b["e2"] for b in batches]),
           "e2_multi": np.concatenate([b["e2_multi"] for b in batches]), "batch_sizes": np.asarray(B_list),
           "mr": mr, "mrr": mrr, "hits_levels": np.asarray(sorted(hits)), "ranks": np.asarray(ranks, np.int64),
           "hits_values": np.asarray([hits[k] for k in sorted(hits)])}
    np.savez_compressed(os.path.join(GOLD, "ref_metrics_%s.npz" % name), **out)
    print("metrics golden", name, "mr", mr, "mrr", mrr)


# ------------------------------------------------------------------------------------------------ models.py
# d / C are kept small for the multi-step cases so the fixtures stay a few hundred KB (conv_num_channels is a
# model_descriptors option of the reference, models.py:111)
CASES = {
    "glinear_eval": dict(ctx=[], bn_train=False, usebn=False, drop=(0.0, 0.0, 0.0), is_train=False, B=7, d=40, C=32),
    "glinear_train": dict(ctx=[], bn_train=True, usebn=False, drop=(0.3, 0.2, 0.0), is_train=True, B=7, d=30, C=8),
    "glinear_train_movingstats": dict(ctx=[], bn_train=False, usebn=False, drop=(0.3, 0.2, 0.0), is_train=True, B=7,
                                      d=30, C=8),
    "gmlp_train": dict(ctx=[6], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=True, B=9, d=30, C=8),
    "gmlp_eval": dict(ctx=[6], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=False, B=9, d=30, C=8),
    # sampled labels (use_negative_sampling, models.py:438-443): L = 12 ids per query, positives + random entities,
    # with repeats inside a row and across rows (the reference's samplers allow both, data.py:236-238)
    "glinear_train_sampled": dict(ctx=[], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=True, B=9, d=30,
                                  C=8, sampled=12),
    # the other two shipped model types: plain ConvE (*_plain.yaml: relation image stacked under the entity image,
    # shared FC weights; rel_emb_size == ent_emb_size) and per-relation parameter tables (*_param_lookup.yaml)
    "plain_train": dict(ctx=None, bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=True, B=9, d=30, C=8,
                        variant="plain"),
    "plain_eval": dict(ctx=None, bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=False, B=7, d=30, C=8,
                       variant="plain"),
    "lookup_train": dict(ctx=[], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=True, B=9, d=30, C=8,
                         variant="param_lookup"),
    "lookup_eval": dict(ctx=[], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=False, B=7, d=30, C=8,
                        variant="param_lookup"),
    # conv filter and bias generated per query too (context_rel_conv, models.py:216-241, 375-380)
    "cpgconv_train": dict(ctx=[6], ctx_conv=[5], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=True, B=9,
                          d=30, C=8),
    "cpgconv_eval": dict(ctx=[], ctx_conv=[], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=False, B=7,
                         d=30, C=8),
    # concat_rel (models.py:270-271, 406-407): rel_emb appended to the flattened conv features, for the generated-FC
    # and the shared-FC model types
    "concat_train": dict(ctx=[6], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=True, B=9, d=30, C=8,
                         concat=True),
    "concat_eval": dict(ctx=[], bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=False, B=7, d=30, C=8,
                        concat=True),
    "concat_plain_train": dict(ctx=None, bn_train=True, usebn=True, drop=(0.3, 0.2, 0.2), is_train=True, B=9, d=30, C=8,
                               variant="plain", concat=True),
}


def gen_model(tf, models, name, case, n_steps=1):
    from oracle import conve_oracle as O
    import torch
    ctx, B = case["ctx"], case["B"]
    d, C = case["d"], case["C"]
    variant = case.get("variant", "cpg")
    dr = d if variant == "plain" else 5
    ctx_conv = case.get("ctx_conv")
    cfg = O.OracleConfig(num_ent=97, num_rel=6, ent_emb_size=d, rel_emb_size=dr, context_rel_out=ctx, variant=variant,
                         context_rel_conv=ctx_conv, conv_num_channels=C, hidden_dropout=case["drop"][0], output_dropout=case["drop"][1],
                         concat_rel=bool(case.get("concat")),
                         context_rel_dropout=case["drop"][2], context_rel_use_batch_norm=case["usebn"],
                         batch_norm_train_stats=case["bn_train"], batch_norm_momentum=0.9)
    p = O.init_params(cfg, seed=11, bias_noise=0.1)
    save = {}
    init = {"ent_emb": p["ent_emb"], "pred_bias": p["pred_bias"]}
    if ctx_conv is None:
        init["conv1_weights"], init["conv1_bias"] = p["conv1_weights"], p["conv1_bias"]
    if variant != "param_lookup":
        init["rel_emb"] = p["rel_emb"]
    F = cfg.fc_input_size
    for which in ("fc_weights", "fc_bias") + (("conv1_weights", "conv1_bias") if ctx_conv is not None else ()):
        if variant == "plain":            # plain tf variables fc_weights [F, d], fc_bias [d] (models.py:334-340)
            init[which] = p[which + "_proj"][0].reshape((F, d) if which == "fc_weights" else (d,))
            continue
        if variant == "param_lookup":     # ParameterLookup tables, named after the generator (models.py:86)
            init[which] = p[which + "_proj"][0]
            continue
        for i, a in enumerate(p[which + "_proj"]):
            init["%s/CPG/Projection%d" % (which, i)] = a
        for i, bn in enumerate(p[which + "_bn"]):
            for k, tfk in (("gamma", "gamma"), ("beta", "beta"), ("moving_mean", "moving_mean"),
                           ("moving_var", "moving_variance")):
                init["%s/CPG/Projection%d/BatchNorm/%s" % (which, i, tfk)] = bn[k]
    for nm in ("Conv1BN", "FCBN"):
        for k, tfk in (("gamma", "gamma"), ("beta", "beta"), ("moving_mean", "moving_mean"),
                       ("moving_var", "moving_variance")):
            init["%s/%s" % (nm, tfk)] = p[nm][k]
    for k, v in init.items():
        save["init/" + k] = np.asarray(v)
    OH, OW = cfg.conv_out_hw
    for step in range(n_steps):
        e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, B, seed=100 + step)
        dense = O.csr_to_dense(rowptr, col, cfg.num_ent)
        # keep-masks = exactly what the CUDA kernels draw for ConvE(seed=0) at this step (oracle/dropout_hash.py)
        from oracle import dropout_hash as DH
        sd = DH.model_seed0(0) + step + 1
        m_fm = DH.keep_mask(B * OH * OW * C, 1 - case["drop"][0], sd, DH.SALT_FEATURE_MAP).reshape(B, OH, OW, C)
        m_out = DH.keep_mask(B * d, 1 - case["drop"][1], sd, DH.SALT_OUTPUT).reshape(B, d)
        m_cw = [DH.keep_mask(B * n, 1 - case["drop"][2], sd, DH.ctx_salt(0, i)).reshape(B, n)
                for i, n in enumerate(ctx or [])]
        m_cb = [DH.keep_mask(B * n, 1 - case["drop"][2], sd, DH.ctx_salt(1, i)).reshape(B, n)
                for i, n in enumerate(ctx or [])]
        m_ccw = [DH.keep_mask(B * n, 1 - case["drop"][2], sd, DH.ctx_salt(2, i)).reshape(B, n)
                 for i, n in enumerate(ctx_conv or [])]
        m_ccb = [DH.keep_mask(B * n, 1 - case["drop"][2], sd, DH.ctx_salt(3, i)).reshape(B, n)
                 for i, n in enumerate(ctx_conv or [])]
        # tf.nn.dropout call order in models.py: [conv1_weights CPG hidden, conv1_bias CPG hidden (:367),] conv1
        # (:390), fc_weights CPG hidden, fc_bias CPG hidden, fc (:414)
        tf.state.reset() if step == 0 else None
        st = tf.state
        if step > 0:
            st.variables, st.trainable, st.update_ops = {}, [], []
        st.dtype = torch.float32
        st.init_values = init
        st.is_train = case["is_train"]
        st.batch = {"e1": e1, "e2": e2, "rel": rel, "e2_multi": dense, "lookup_values": np.zeros((B, 0), np.int32)}
        L = case.get("sampled")
        if L:
            rng = np.random.default_rng(500 + step)
            lookup = rng.integers(0, cfg.num_ent, (B, L)).astype(np.int32)
            for i in range(B):                                   # up to 3 positives first, then random entities
                pos = col[rowptr[i]:rowptr[i + 1]][:3]
                lookup[i, :len(pos)] = pos
            lookup[:, -1] = lookup[:, 0]                         # a repeated id inside every row
            labels = dense[np.arange(B)[:, None], lookup].astype(np.float32)
            st.batch["e2_multi"], st.batch["lookup_values"] = labels, lookup
        st.dropout_masks = m_ccw + m_ccb + [m_fm] + m_cw + m_cb + [m_out]
        st.dropout_calls = 0
        model = models.ConvE(model_descriptors={
            "use_negative_sampling": bool(case.get("sampled")), "label_smoothing_epsilon": 0.1, "num_ent": cfg.num_ent,
            "num_rel": cfg.num_rel, "ent_emb_size": d, "rel_emb_size": dr, "concat_rel": bool(case.get("concat")),
            "conv_num_channels": C,
            "context_rel_conv": ctx_conv, "context_rel_out": ctx, "context_rel_dropout": case["drop"][2],
            "context_rel_use_batch_norm": case["usebn"], "input_dropout": 0.2, "hidden_dropout": case["drop"][0],
            "output_dropout": case["drop"][1], "learning_rate": 1e-2, "batch_size": B, "add_loss_summaries": False,
            "add_variable_summaries": False, "add_tensor_summaries": False, "batch_norm_momentum": 0.9,
            "batch_norm_train_stats": case["bn_train"], "do_parameter_lookup": variant == "param_lookup"})
        pre = "step%d/" % step
        save[pre + "e1"], save[pre + "rel"], save[pre + "e2"] = e1, rel, e2
        save[pre + "rowptr"], save[pre + "col"] = rowptr, col
        save[pre + "mask_fm"], save[pre + "mask_out"] = m_fm, m_out
        for i, (a, b) in enumerate(zip(m_cw, m_cb)):
            save[pre + "mask_cw%d" % i], save[pre + "mask_cb%d" % i] = a, b
        for i, (a, b) in enumerate(zip(m_ccw, m_ccb)):
            save[pre + "mask_ccw%d" % i], save[pre + "mask_ccb%d" % i] = a, b
        if L:
            save[pre + "lookup"], save[pre + "labels"] = lookup, labels
            save[pre + "predictions_lookup"] = model.predictions_lookup.detach().numpy()
        save[pre + "loss"] = float(model.loss.detach())
        save[pre + "predictions_all"] = model.predictions_all.detach().numpy()
        save[pre + "predicted_e2_emb"] = model.predicted_e2_emb.detach().numpy()
        for k, g in st.last_gradients.items():
            if g is not None:
                save[pre + "grad/" + k] = g.numpy()
        # state after train_op (variables, BN moving stats, optimizer slots, beta powers) -> next step's init
        init = {k: v.numpy() for k, v in st.variables.items()}
        for k, v in init.items():
            save[pre + "after/" + k] = v
    np.savez_compressed(os.path.join(GOLD, "ref_model_%s.npz" % name), **save)
    print("model golden", name, "loss", [save["step%d/loss" % s] for s in range(n_steps)])


def main():
    os.makedirs(GOLD, exist_ok=True)
    tf, models, metrics = _import_reference()
    only = set(sys.argv[1:])
    if not only:
        gen_metrics(tf, metrics, "small", [16, 16, 5], 211, seed=1, filt_p=0.05)
        gen_metrics(tf, metrics, "dense_filter", [8], 64, seed=2, filt_p=0.5)
        gen_metrics(tf, metrics, "single", [1], 10, seed=3, filt_p=0.2)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        gen_model(tf, models, name, case, n_steps=3 if case["is_train"] else 1)


if __name__ == "__main__":
    main()
