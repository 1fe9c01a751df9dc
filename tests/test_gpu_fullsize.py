"""BASELINE.json's named configurations at FULL size on one GPU, through the public API.  The NumPy oracle would need
minutes (and tens of GB for the materialised [B, F, d] weights) at these sizes, so the checks are the size-independent
properties of the path: a step is bit-reproducible, the loss of a random-init model sits at the value the label
smoothing dictates, and the fused filtered rank equals the reference's counting rule (metrics.py:44-51) applied to the
logits the same model writes out — bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(shape, prec):
    from coper_b200 import synthetic
    from coper_b200.models import ConvE
    s = synthetic.SHAPES[shape]
    return ConvE(synthetic.descriptors(shape, dropout=True), seed=0, prec=prec, conv_in_height=s["H"])


@pytest.mark.parametrize("prec", ["tf32x3", "fp16x3", "bf16"])
@pytest.mark.parametrize("shape", ["fb15k-237", "wn18rr", "nell-995", "yago3-10"])
def test_full_size_step_is_reproducible_and_ranks_follow_the_counting_rule(shape, prec):
    from coper_b200 import synthetic
    s = synthetic.SHAPES[shape]
    N, B = s["num_ent"], s["batch"]
    host = synthetic.make_batches(N, s["num_rel"], B, 2, seed=3)
    m1, m2 = _model(shape, prec), _model(shape, prec)
    losses = []
    for m in (m1, m2):
        losses.append([float(m.train_step(host[0]).item()) for _ in range(3)])     # eager, capture, replay
    assert losses[0] == losses[1] and all(np.isfinite(losses[0]))
    assert torch.equal(m1.ent_emb, m2.ent_emb) and torch.equal(m1.rel_emb, m2.rel_emb)
    assert torch.equal(m1.fc_weights.projections[-1], m2.fc_weights.projections[-1])
    # random init: logits ~ 0 -> BCE ~ ln 2 for (almost) every one of the B*N labels (models.py:448-453)
    assert abs(losses[0][0] - np.log(2.0)) < 0.02
    # evaluation: ranks straight from the scorer's accumulators == counting rule on the logits written to HBM
    hb = host[1]
    S = m1.predict_all(hb).clone()
    rank, n_equal = m1.filtered_ranks(hb)
    e2 = torch.as_tensor(hb["e2"]).cuda()
    rowptr = torch.as_tensor(hb["e2_multi_rowptr"].astype(np.int64)).cuda()
    col = torch.as_tensor(hb["e2_multi_col"].astype(np.int64)).cuda()
    rows = torch.repeat_interleave(torch.arange(B, device="cuda"), rowptr[1:] - rowptr[:-1])
    valid = torch.ones(B, N, dtype=torch.bool, device="cuda")
    valid[rows, col] = False                                 # every known true tail is filtered ...
    ar = torch.arange(B, device="cuda")
    valid[ar, e2] = False                                    # ... and the gold entity is never compared with itself
    gold = S[ar, e2]
    n_greater = ((S > gold[:, None]) & valid).sum(1).to(torch.int32)
    n_eq_ref = ((S == gold[:, None]) & valid).sum(1).to(torch.int32)
    assert torch.equal(rank, n_greater + 1) and torch.equal(n_equal, n_eq_ref)
    assert int(rank.min()) >= 1 and int(rank.max()) <= N
