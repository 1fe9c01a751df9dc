"""Synthetic workloads of the BASELINE.json shapes (no real training split is available offline).

Value distributions follow SURVEY §8(d): Xavier-uniform tables, uniform head entities / relations,
positives per row ~ 1 + Geometric (mean ~3) drawn without replacement; eval gold ``e2`` is one of the
positives and the filter set is the positive set.
"""
from __future__ import annotations

import numpy as np

# name -> model_descriptors overrides + batch size (reference configs: qa_cpg/configs/config_<name>_cpg.yaml)
SHAPES = {
    "fb15k-237": dict(num_ent=14541, num_rel=474, ent_emb_size=200, rel_emb_size=32, batch=512, H=10,
                      bn_momentum=0.99),
    "wn18rr": dict(num_ent=40943, num_rel=22, ent_emb_size=200, rel_emb_size=8, batch=512, H=10, bn_momentum=0.1),
    "nell-995": dict(num_ent=75492, num_rel=400, ent_emb_size=200, rel_emb_size=32, batch=512, H=10,
                     bn_momentum=0.1),
    "yago3-10": dict(num_ent=123182, num_rel=74, ent_emb_size=200, rel_emb_size=37, batch=128, H=10,
                     bn_momentum=0.1),
    "synth-10m": dict(num_ent=10_000_000, num_rel=2000, ent_emb_size=256, rel_emb_size=32, batch=512, H=16,
                      bn_momentum=0.1),
    "toy": dict(num_ent=997, num_rel=6, ent_emb_size=40, rel_emb_size=5, batch=32, H=10, bn_momentum=0.1),
}


def descriptors(shape: str, dropout: bool = True, **over):
    """``model_descriptors`` (run_cpg.py:115-137) of the <shape>_cpg config with full 1-N labels."""
    s = SHAPES[shape]
    md = {"use_negative_sampling": False, "label_smoothing_epsilon": 0.1, "num_ent": s["num_ent"],
          "num_rel": s["num_rel"], "ent_emb_size": s["ent_emb_size"], "rel_emb_size": s["rel_emb_size"],
          "concat_rel": False, "context_rel_conv": None, "context_rel_out": [], "context_rel_dropout": 0.2,
          "context_rel_use_batch_norm": True, "input_dropout": 0.2, "hidden_dropout": 0.3 if dropout else 0.0,
          "output_dropout": 0.2 if dropout else 0.0, "learning_rate": 0.001, "batch_size": s["batch"],
          "add_loss_summaries": False, "add_variable_summaries": False, "add_tensor_summaries": False,
          "batch_norm_momentum": s["bn_momentum"], "batch_norm_train_stats": True, "do_parameter_lookup": False}
    md.update(over)
    return md


def make_batches(num_ent: int, num_rel: int, B: int, n_batches: int, seed: int = 0, mean_pos: float = 3.0,
                 max_pos: int = 1000):
    """List of host batches {e1, rel, e2 int64 [B]; e2_multi_rowptr int32 [B+1]; e2_multi_col int32 [nnz]}."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_batches):
        e1 = rng.integers(0, num_ent, B, dtype=np.int64)
        rel = rng.integers(0, num_rel, B, dtype=np.int64)
        k = np.minimum(rng.geometric(1.0 / mean_pos, B), min(max_pos, num_ent)).astype(np.int64)
        rowptr = np.zeros(B + 1, np.int32)
        rowptr[1:] = np.cumsum(k)
        # sampling with replacement then de-duplicating per row keeps this O(nnz) at 10M entities
        col = rng.integers(0, num_ent, int(rowptr[-1]), dtype=np.int64).astype(np.int32)
        e2 = col[rowptr[:-1]].astype(np.int64)
        out.append({"e1": e1, "rel": rel, "e2": e2, "e2_multi_rowptr": rowptr, "e2_multi_col": col})
    return out


def to_sampled(batch, num_ent: int, num_labels: int, seed: int = 0, prop_negatives: float = 10.0):
    """Sampled-label form of a 1-N batch (data.py:228-277): [B, L] lookup ids = the query's positives (at most
    L / (1 + prop_negatives) of them) followed by random entities, and their labels."""
    rng = np.random.default_rng(seed)
    B = len(batch["e1"])
    rp, col = batch["e2_multi_rowptr"], batch["e2_multi_col"]
    lookup = rng.integers(0, num_ent, (B, num_labels), dtype=np.int64)
    labels = np.zeros((B, num_labels), np.float32)
    n_pos_needed = max(1, int(1.0 / (1.0 + prop_negatives) * num_labels))
    for i in range(B):
        pos = col[rp[i]:rp[i + 1]][:n_pos_needed]
        lookup[i, :len(pos)] = pos
        labels[i] = np.isin(lookup[i], col[rp[i]:rp[i + 1]])
    return {"e1": batch["e1"], "rel": batch["rel"], "e2": batch["e2"], "e2_multi": labels,
            "lookup_values": lookup.astype(np.int32)}
