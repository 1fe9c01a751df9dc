"""CPU tests of the oracle itself: analytic backward vs autograd of the literal (materialising) torch port,
count-rank vs literal argsort-rank, AMSGrad restatement.  (No GPU, no CUDA library.)"""
import numpy as np
import pytest
import torch

from oracle import conve_oracle as O
from oracle.torch_port import TorchPort


def _case(ctx, bn_train, usebn, B=7, seed=1):
    cfg = O.OracleConfig(num_ent=97, num_rel=6, ent_emb_size=40, rel_emb_size=5, context_rel_out=ctx,
                         hidden_dropout=0.3, output_dropout=0.2, context_rel_dropout=0.2,
                         context_rel_use_batch_norm=usebn, batch_norm_train_stats=bn_train, batch_norm_momentum=0.9)
    p = O.init_params(cfg, seed, bias_noise=0.1)
    e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, B, seed + 2)
    z = O.csr_to_dense(rowptr, col, cfg.num_ent)
    rng = np.random.default_rng(seed + 4)
    OH, OW = cfg.conv_out_hw
    masks = {"feature_map": rng.random((B, OH, OW, 32)) < 0.7, "output": rng.random((B, 40)) < 0.8,
             "ctx_w": [rng.random((B, n)) < 0.8 for n in ctx], "ctx_b": [rng.random((B, n)) < 0.8 for n in ctx]}
    return cfg, p, e1, rel, e2, z, masks


def _rel(a, b):
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-6)


@pytest.mark.parametrize("ctx,bn_train,usebn", [([], True, False), ([6], True, True), ([6], False, True),
                                                 ([], False, False)])
def test_analytic_backward_matches_autograd(ctx, bn_train, usebn):
    cfg, p, e1, rel, e2, z, masks = _case(ctx, bn_train, usebn)
    out = O.forward(p, cfg, e1, rel, True, masks, z, np.float64)
    g = O.backward(out, cfg)
    port = TorchPort(p, cfg, torch.float64)
    L, tg = port.loss_and_grads(e1, rel, z, True, masks)
    assert abs(out["loss"] - L) < 1e-12
    assert _rel(g["ent_emb"], tg["ent_emb"]) < 1e-9
    assert _rel(g["rel_emb"], tg["rel_emb"]) < 1e-9
    assert _rel(g["conv1_weights"], tg["conv1_weights"]) < 1e-9
    assert _rel(g["conv1_bias"], tg["conv1_bias"]) < 1e-9
    assert _rel(g["pred_bias"], tg["pred_bias"]) < 1e-9
    for i, a in enumerate(g["fc_weights_proj"]):
        assert _rel(a, tg["fc_weights_proj.%d" % i]) < 1e-9
    for i, a in enumerate(g["fc_bias_proj"]):
        assert _rel(a, tg["fc_bias_proj.%d" % i]) < 1e-9
    for nm in ("FCBN", "Conv1BN"):
        assert _rel(g[nm]["gamma"], tg[nm + ".gamma"]) < 1e-9
        assert _rel(g[nm]["beta"], tg[nm + ".beta"]) < 1e-9
    if usebn and ctx:
        assert _rel(g["fc_weights_bn"][0]["gamma"], tg["fc_weights_bn.0.gamma"]) < 1e-9


def test_fused_equals_materialised_fp32():
    """(c (x) f).P^ == bmm(f, reshape(c.P)) to fp32 round-off (SURVEY §8c cross-check)."""
    cfg, p, e1, rel, e2, z, _ = _case([], False, False)
    out = O.forward(p, cfg, e1, rel, False, None, z, np.float32)
    port = TorchPort(p, cfg, torch.float32)
    with torch.no_grad():
        S, _ = port.predict(e1, rel, False)
    assert np.abs(out["scores"] - S.numpy()).max() < 5e-6


def test_rank_count_equals_literal():
    rng = np.random.default_rng(0)
    B, N = 64, 501
    pred = rng.normal(size=(B, N)).astype(np.float32)
    e2 = rng.integers(0, N, B)
    filt = (rng.random((B, N)) < 0.05).astype(np.float32)
    filt[np.arange(B), e2] = 1.0
    rl = O.rank_literal(pred, e2, filt)
    rc, ne = O.rank_count(pred, e2, filt)
    assert ne.sum() == 0
    assert (rl == rc).all()
    mr, mrr, hits = O.summarize_ranks(rl)
    assert mr == np.mean(rl) and abs(mrr - np.mean(1.0 / rl)) < 1e-15
    assert hits[1] == np.mean(rl <= 1)


def test_rank_ties_reported():
    pred = np.zeros((2, 10), np.float32)
    e2 = np.array([3, 4])
    filt = np.zeros((2, 10), np.float32)
    rc, ne = O.rank_count(pred, e2, filt)
    assert (rc == 1).all() and (ne == 9).all()


def test_amsgrad_oracle_matches_port_and_bug():
    cfg, p, e1, rel, e2, z, _ = _case([], True, False)
    port = TorchPort(p, cfg, torch.float64, lr=1e-2)
    l0 = port.train_step(e1, rel, z)
    l1 = port.train_step(e1, rel, z)
    assert l1 < l0
    # compat mode: slots m, v stay exactly zero (amsgrad.py:142-151)
    assert all(float(s["m"].abs().max()) == 0 and float(s["v"].abs().max()) == 0 for s in port.slots.values())
    opt = O.AMSGradOracle(1e-2)
    th = np.ones(5)
    g = np.linspace(-1, 1, 5)
    opt.apply({"x": (th, g)})
    lr_t = 1e-2 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = 1 - lr_t * 0.1 * g / (np.sqrt(0.001 * g * g) + 1e-8)
    assert np.allclose(th, exp)


@pytest.mark.parametrize("variant,kw", [
    ("plain", dict(rel_emb_size=40, context_rel_out=None)),
    ("param_lookup", dict(rel_emb_size=5, context_rel_out=[])),
    ("cpg", dict(rel_emb_size=5, context_rel_out=[6], context_rel_conv=[7], context_rel_use_batch_norm=True)),
])
def test_other_model_types_backward_matches_finite_differences(variant, kw):
    """plain ConvE, ParameterLookup tables and generated conv filters: the analytic backward of the oracle against
    central differences of its own fp64 forward (dropout masks fixed, batch-statistics BN)."""
    cfg = O.OracleConfig(num_ent=31, num_rel=4, ent_emb_size=40, conv_num_channels=4, batch_norm_train_stats=True,
                         hidden_dropout=0.3, output_dropout=0.2, context_rel_dropout=0.2, variant=variant, **kw)
    p = O.cast_params(O.init_params(cfg, 1, 0.1), np.float64)
    B = 6
    e1, rel, e2, rp, col = O.synthetic_batch(cfg, B, 2)
    dense = O.csr_to_dense(rp, col, cfg.num_ent, np.float64)
    rng = np.random.default_rng(0)
    OH, OW = cfg.conv_out_hw
    masks = {"feature_map": rng.random((B, OH, OW, 4)) < 0.7, "output": rng.random((B, 40)) < 0.8,
             "ctx_w": [rng.random((B, 6)) < 0.8], "ctx_b": [rng.random((B, 6)) < 0.8],
             "ctx_cw": [rng.random((B, 7)) < 0.8], "ctx_cb": [rng.random((B, 7)) < 0.8]}

    def loss_of():
        return O.forward(p, cfg, e1, rel, True, masks, dense, np.float64)["loss"]
    g = O.backward(O.forward(p, cfg, e1, rel, True, masks, dense, np.float64), cfg)
    checks = [(p["ent_emb"], g["ent_emb"]), (p["fc_weights_proj"][-1], g["fc_weights_proj"][-1]),
              (p["fc_bias_proj"][-1], g["fc_bias_proj"][-1]), (p["FCBN"]["gamma"], g["FCBN"]["gamma"])]
    if variant != "param_lookup":
        checks.append((p["rel_emb"], g["rel_emb"]))
    if cfg.context_rel_conv is not None:
        for i in range(2):
            checks.append((p["conv1_weights_proj"][i], g["conv1_weights_proj"][i]))
            checks.append((p["conv1_bias_proj"][i], g["conv1_bias_proj"][i]))
        checks.append((p["conv1_weights_bn"][0]["gamma"], g["conv1_weights_bn"][0]["gamma"]))
    else:
        checks.append((p["conv1_weights"], g["conv1_weights"]))
    for arr, grad in checks:
        idxs = np.argwhere(np.abs(grad) > 0.1 * np.abs(grad).max())
        for idx in idxs[rng.choice(len(idxs), 3)]:
            idx, h = tuple(idx), 1e-5
            old = arr[idx]
            arr[idx] = old + h
            lp = loss_of()
            arr[idx] = old - h
            lm = loss_of()
            arr[idx] = old
            assert abs((lp - lm) / (2 * h) - grad[idx]) < 1e-6 * abs(grad[idx]) + 1e-12
    if variant == "param_lookup":       # the IndexedSlices of the tables sum to the dense gradients
        vals, idx = g["_sparse"]["fc_weights"][0]
        dense_g = np.zeros_like(g["fc_weights_proj"][0])
        np.add.at(dense_g, idx, vals)
        assert np.abs(dense_g - g["fc_weights_proj"][0]).max() < 1e-15
