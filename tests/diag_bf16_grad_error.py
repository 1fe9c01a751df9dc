"""Diagnostic (not collected by pytest): how far the bf16 engine's downstream gradients sit from the fp64 oracle.
Lives under tests/ because it uses the oracle (only tests/, smoke() and bench.py's CPU legs may)."""
import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from oracle import conve_oracle as O
import test_gpu_model as T
for name in ("toy_glinear_batch_stats", "ragged_mid", "d256_16x16"):
    kw, B = T.CASES[name]
    cfg = O.OracleConfig(**kw)
    params = O.init_params(cfg, seed=3, bias_noise=0.05)
    e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, B, seed=5)
    dense = O.csr_to_dense(rowptr, col, cfg.num_ent)
    for prec in ("bf16",):
        m = T.make(cfg, params, prec=prec)
        m.train_step(T.batch_of(e1, rel, e2, rowptr, col), apply_update=False)
        masks = T.export_masks(m, cfg, B)
        out = O.forward(params, cfg, e1, rel, True, masks, dense, np.float64)
        g = O.backward(out, cfg)
        b = m._bufs[B]
        print(name, prec, "dq %.2e dy %.2e df %.2e dE %.2e" % (T.relerr(b.dq.cpu().numpy(), g["_dq"]), T.relerr(b.dy.cpu().numpy(), g["_dy"]),
              T.relerr(b.df.cpu().numpy(), g["_df"]), T.relerr(m.grads["ent_emb"].cpu().numpy(), g["ent_emb"])))
