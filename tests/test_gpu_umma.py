"""tcgen05/TMEM engine parity (through the C ABI).

Tolerances:
  * COPER_PREC_TF32X3 (3-term error-compensated tf32): max|S - S64| <= 1e-5 * max|S64|  (fp32-class; the BASELINE bar)
  * COPER_PREC_BF16: vs an fp64 product of the bf16-ROUNDED operands <= 1e-5 (the tensor pipe accumulates in fp32);
    vs the unrounded fp64 product <= 1e-2 (the stated tolerance of the bf16 path).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PREC = {"bf16": 1, "tf32x3": 2, "fp16x3": 3}
FP32_CLASS = ("tf32x3", "fp16x3")      # 3-term compensated engines: fp32-class accuracy on the tensor pipe


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def bf16_round(x):
    return torch.as_tensor(x, dtype=torch.float32).to(torch.bfloat16).to(torch.float64).numpy()


@pytest.mark.parametrize("prec", ["bf16", "tf32x3", "fp16x3"])
@pytest.mark.parametrize("B,N,d", [(512, 40943, 200), (7, 97, 40), (130, 1003, 200), (33, 5000, 256), (512, 70001, 256),
                                   (128, 256, 64), (1, 1, 8)])
def test_score1n_fwd_tensor_pipe(prec, B, N, d):
    from coper_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(B + N + d)
    q = np.maximum(rng.normal(size=(B, d)), 0).astype(np.float32)
    E = rng.uniform(-0.05, 0.05, size=(N, d)).astype(np.float32)
    bias = (rng.normal(size=N) * 0.1).astype(np.float32)
    ld = -(-N // 32) * 32
    tq, tE, tb = torch.as_tensor(q).cuda(), torch.as_tensor(E).cuda(), torch.as_tensor(bias).cuda()
    S = torch.full((B, ld), float("nan"), device="cuda")
    p = PREC[prec]
    ws = torch.empty(lib.coper_score1n_workspace_bytes(B, N, d, p), dtype=torch.uint8, device="cuda")
    L.call("coper_score1n_fwd", L.ptr(tq), L.ptr(tE), L.ptr(tb), B, N, d, L.ptr(S), ld, L.ptr(ws), ws.numel(), p)
    got = S[:, :N].cpu().numpy()
    assert np.isfinite(got).all()
    exact = q.astype(np.float64) @ E.astype(np.float64).T + bias
    if prec in FP32_CLASS:
        assert relerr(got, exact) < 1e-5
    else:
        rounded = bf16_round(q) @ bf16_round(E).T + bias
        assert relerr(got, rounded) < 1e-5
        assert relerr(got, exact) < 1e-2
    # prepared-operand entry point gives bit-identical results
    qp = torch.empty(lib.coper_prepared_bytes(B, d, p), dtype=torch.uint8, device="cuda")
    Ep = torch.empty(lib.coper_prepared_bytes(N, d, p), dtype=torch.uint8, device="cuda")
    L.call("coper_prepare_operand", L.ptr(tq), B, d, d, p, L.ptr(qp))
    L.call("coper_prepare_operand", L.ptr(tE), N, d, d, p, L.ptr(Ep))
    S2 = torch.zeros(B, ld, device="cuda")
    L.call("coper_score1n_fwd_prepared", L.ptr(qp), L.ptr(Ep), L.ptr(tb), B, N, d, L.ptr(S2), ld, p)
    assert torch.equal(S2[:, :N], S[:, :N])


@pytest.mark.parametrize("prec", ["bf16", "tf32x3", "fp16x3"])
@pytest.mark.parametrize("M,N,K", [(130, 77, 301), (512, 200, 4608), (128, 200, 512), (300, 201, 100), (5, 8, 8),
                                   (1000, 256, 64)])
def test_tc_gemm_all_layouts(prec, M, N, K):
    """C = op(A).op(B) with K-major and MN-major operands (transA / transB) on the tensor pipe."""
    from coper_b200 import _lib as L
    lib = L.load()
    p = PREC[prec]
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.normal(size=(M, K)).astype(np.float32)
    Bm = rng.normal(size=(K, N)).astype(np.float32)
    exact = A.astype(np.float64) @ Bm.astype(np.float64)
    rounded = bf16_round(A) @ bf16_round(Bm)
    ws = torch.empty(lib.coper_tc_gemm_workspace_bytes(M, N, K, p), dtype=torch.uint8, device="cuda")
    for ta in (0, 1):
        for tb in (0, 1):
            a = torch.as_tensor(np.ascontiguousarray(A.T if ta else A)).cuda()
            b = torch.as_tensor(np.ascontiguousarray(Bm.T if tb else Bm)).cuda()
            c = torch.full((M, N), float("nan"), device="cuda")
            L.call("coper_tc_gemm", ta, tb, M, N, K, L.ptr(a), a.shape[1], L.ptr(b), b.shape[1], L.ptr(c), N, p,
                   L.ptr(ws), ws.numel())
            got = c.cpu().numpy()
            assert np.isfinite(got).all(), (ta, tb)
            if prec in FP32_CLASS:
                assert relerr(got, exact) < 1e-5, (ta, tb, relerr(got, exact))
            else:
                assert relerr(got, rounded) < 1e-5, (ta, tb, relerr(got, rounded))


@pytest.mark.parametrize("prec", ["bf16", "tf32x3", "fp16x3"])
@pytest.mark.parametrize("B,N,d", [(512, 40943, 200), (7, 97, 40), (130, 1003, 200), (33, 5000, 256), (128, 70001, 256),
                                   (4096, 5118, 200),      # one rank's share of 8-way data-parallel WN18RR
                                   # BASELINE.json configs at full size: FB15k-237, NELL-995, YAGO3-10 (B = 128)
                                   (512, 14541, 200), (512, 75492, 200), (128, 123182, 200)])
def test_score1n_bce_fwd_bwd_tensor_pipe(prec, B, N, d):
    """scorer + label-smoothed BCE + gradient on tcgen05 (models.py:433-437,448-453,198): loss, dq, dE, dbias.
    tf32x3: fp32-class (loss 1e-5 rel, gradients 2e-5 of max); bf16: vs fp64 arithmetic on bf16-rounded q, E
    (loss 1e-5; gradients additionally carry the bf16 rounding of G -> 1e-2 of max)."""
    from coper_b200 import _lib as L
    from oracle import conve_oracle as O
    lib = L.load()
    p = PREC[prec]
    rng = np.random.default_rng(B * 3 + N + d)
    q = np.maximum(rng.normal(size=(B, d)), 0).astype(np.float32)
    E = rng.uniform(-0.05, 0.05, size=(N, d)).astype(np.float32)
    bias = (rng.normal(size=N) * 0.1).astype(np.float32)
    cfg = O.OracleConfig(num_ent=N, num_rel=2, ent_emb_size=d, rel_emb_size=2, conv_in_height=d // 4 if d % 10 else 10)
    _, _, _, rowptr, col = O.synthetic_batch(cfg, B, seed=9, mean_pos=4.0)
    words, ld = -(-N // 32), -(-N // 32) * 32
    bits = torch.full((N, -(-B // 32)), -1, dtype=torch.int32, device="cuda")     # entity-major label bits
    trp, tcol = torch.as_tensor(rowptr.astype(np.int32)).cuda(), torch.as_tensor(col.astype(np.int32)).cuda()
    L.call("coper_csr_to_bits_t", L.ptr(trp), L.ptr(tcol), B, 0, N, L.ptr(bits))
    z = O.csr_to_dense(rowptr, col, N, np.float64)
    pos, neg = np.float32(np.float32(0.9) + np.float32(1.0 / N)), np.float32(1.0 / N)
    zs = np.where(z > 0, np.float64(pos), np.float64(neg))
    inv = 1.0 / (B * N)
    tq, tE, tb = torch.as_tensor(q).cuda(), torch.as_tensor(E).cuda(), torch.as_tensor(bias).cuda()
    ws = torch.empty(lib.coper_score1n_bce_workspace_bytes(B, N, d, p), dtype=torch.uint8, device="cuda")
    G = torch.empty(lib.coper_score1n_bce_G_bytes(B, N, p), dtype=torch.uint8, device="cuda")
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    nan = float("nan")
    dq, dE, db = (torch.full(s, nan, device="cuda") for s in ((B, d), (N, d), (N,)))
    L.call("coper_score1n_bce_fwd_bwd", L.ptr(tq), L.ptr(tE), None, L.ptr(tb), L.ptr(bits), B, N, d, float(pos), float(neg),
           inv, L.ptr(loss), L.ptr(G), ld, L.ptr(dq), L.ptr(dE), L.ptr(db), L.ptr(ws), ws.numel(), p)
    torch.cuda.synchronize()
    if prec in FP32_CLASS:
        qq, EE = q.astype(np.float64), E.astype(np.float64)
    else:
        qq, EE = bf16_round(q), bf16_round(E)
    Sr = qq @ EE.T + bias
    el = np.maximum(Sr, 0) - Sr * zs + np.log1p(np.exp(-np.abs(Sr)))
    Gr = (O.sigmoid(Sr) - zs) * inv
    assert abs(loss.item() - el.sum()) < 1e-5 * el.sum()
    gtol = 2e-5 if prec in FP32_CLASS else 1e-2
    for got, ref in ((dq, Gr @ EE), (dE, Gr.T @ qq), (db, Gr.sum(0))):
        got = got.cpu().numpy()
        assert np.isfinite(got).all()
        assert relerr(got, ref) < gtol
    # determinism: a second call is bit-identical
    dq2, dE2, db2 = (torch.zeros_like(t) for t in (dq, dE, db))
    loss2 = torch.zeros_like(loss)
    L.call("coper_score1n_bce_fwd_bwd", L.ptr(tq), L.ptr(tE), None, L.ptr(tb), L.ptr(bits), B, N, d, float(pos), float(neg),
           inv, L.ptr(loss2), L.ptr(G), ld, L.ptr(dq2), L.ptr(dE2), L.ptr(db2), L.ptr(ws), ws.numel(), p)
    assert torch.equal(dq, dq2) and torch.equal(dE, dE2) and torch.equal(db, db2) and torch.equal(loss, loss2)


@pytest.mark.parametrize("prec", ["bf16", "tf32x3", "fp16x3"])
@pytest.mark.parametrize("B,N,d,lo", [(512, 40943, 200, 0), (7, 97, 40, 0), (130, 1003, 200, 0), (33, 5000, 256, 0),
                                      (64, 70001, 256, 123456), (300, 33, 64, 0), (4096, 5118, 200, 5118),
                                      (512, 14541, 200, 0), (512, 75492, 200, 0), (128, 123182, 200, 0)])
def test_fused_score_rank_equals_two_pass(prec, B, N, d, lo):
    """scorer + filtered rank fused (logits stay in TMEM) == scoring into HBM + coper_filtered_rank, bit for bit:
    the gold logits from the gather+diagonal pass equal the stored logits, and the integer counts agree — including
    exact ties (duplicated entity rows), filtered entities, and gold entities owned by another shard (ent_lo > 0)."""
    from coper_b200 import _lib as L
    from oracle import conve_oracle as O
    lib = L.load()
    p = PREC[prec]
    rng = np.random.default_rng(B + N + d)
    q = np.maximum(rng.normal(size=(B, d)), 0).astype(np.float32)
    E = rng.uniform(-0.05, 0.05, size=(N, d)).astype(np.float32)
    if N > 64:
        E[N // 2:N // 2 + N // 8] = E[:N // 8]                  # duplicated entities -> exact score ties
    bias = (rng.normal(size=N) * 0.1).astype(np.float32)
    if N > 64:
        bias[N // 2:N // 2 + N // 8] = bias[:N // 8]
    e2 = rng.integers(lo, lo + N, B)
    if lo:
        e2[::3] = rng.integers(0, lo, len(e2[::3]))             # gold owned by another shard
    filt = rng.random((B, N)) < 0.05
    own = (e2 >= lo) & (e2 < lo + N)
    filt[np.arange(B)[own], (e2 - lo)[own]] = True
    words, ld = -(-N // 32), -(-N // 32) * 32
    packed = np.zeros((B, words * 32), bool)
    packed[:, :N] = filt
    bits = np.packbits(packed.reshape(B, words, 32), axis=2, bitorder="little").view(np.uint32).reshape(B, words)
    tq, tE, tb = torch.as_tensor(q).cuda(), torch.as_tensor(E).cuda(), torch.as_tensor(bias).cuda()
    te2, tbits = torch.as_tensor(e2).cuda(), torch.as_tensor(bits.view(np.int32)).cuda()
    wordsB = -(-B // 32)
    packedT = np.zeros((N, wordsB * 32), bool)
    packedT[:, :B] = filt.T
    bitsT = np.packbits(packedT.reshape(N, wordsB, 32), axis=2, bitorder="little").view(np.uint32).reshape(N, wordsB)
    tbitsT = torch.as_tensor(bitsT.view(np.int32)).cuda()
    # the library's own builders give the same matrix
    nz = [np.flatnonzero(filt[b]) + lo for b in range(B)]
    rp = np.zeros(B + 1, np.int32); rp[1:] = np.cumsum([len(x) for x in nz])
    cl = np.concatenate(nz).astype(np.int32) if rp[-1] else np.zeros(0, np.int32)
    chk = torch.full((N, wordsB), -1, dtype=torch.int32, device="cuda")
    trp = torch.as_tensor(rp).cuda()                      # keep references: the calls are asynchronous
    tcl = torch.as_tensor(cl).cuda() if len(cl) else torch.zeros(1, dtype=torch.int32, device="cuda")
    L.call("coper_csr_to_bits_t", L.ptr(trp), L.ptr(tcl), B, lo, lo + N, L.ptr(chk))
    torch.cuda.synchronize()
    assert torch.equal(chk, tbitsT)
    chk2 = torch.zeros((N, wordsB), dtype=torch.int32, device="cuda")
    tdense = torch.as_tensor(filt.astype(np.float32)).cuda()
    L.call("coper_dense_to_bits_t", L.ptr(tdense), B, N, N, L.ptr(chk2))
    torch.cuda.synchronize()
    assert torch.equal(chk2, tbitsT)
    qp = torch.empty(lib.coper_prepared_bytes(B, d, p), dtype=torch.uint8, device="cuda")
    Ep = torch.empty(lib.coper_prepared_bytes(N, d, p), dtype=torch.uint8, device="cuda")
    L.call("coper_prepare_operand", L.ptr(tq), B, d, d, p, L.ptr(qp))
    L.call("coper_prepare_operand", L.ptr(tE), N, d, d, p, L.ptr(Ep))
    # two-pass
    S = torch.zeros(B, ld, device="cuda")
    L.call("coper_score1n_fwd_prepared", L.ptr(qp), L.ptr(Ep), L.ptr(tb), B, N, d, L.ptr(S), ld, p)
    gold2 = torch.zeros(B, device="cuda")
    L.call("coper_gold_scores", L.ptr(S), ld, B, N, L.ptr(te2), lo, L.ptr(gold2))
    ng2, ne2 = torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")
    L.call("coper_filtered_rank", L.ptr(S), ld, B, N, L.ptr(te2), lo, L.ptr(gold2), L.ptr(tbits), L.ptr(ng2), L.ptr(ne2))
    # fused
    ws = torch.empty(max(256, lib.coper_score1n_rank_workspace_bytes(B, d, p)), dtype=torch.uint8, device="cuda")
    gold = torch.full((B,), float("nan"), device="cuda")
    L.call("coper_score1n_gold_prepared", L.ptr(qp), L.ptr(Ep), L.ptr(tb), B, N, d, L.ptr(te2), lo, L.ptr(gold),
           L.ptr(ws), ws.numel(), p)
    assert torch.equal(gold, gold2)                             # bit-identical gold logits
    if lo:
        gold = gold + torch.as_tensor((~own).astype(np.float32) * 0.0123).cuda()   # "other shard's" gold logit
        ng2.zero_(); ne2.zero_()
        L.call("coper_filtered_rank", L.ptr(S), ld, B, N, L.ptr(te2), lo, L.ptr(gold), L.ptr(tbits), L.ptr(ng2), L.ptr(ne2))
    ng, ne = torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")
    L.call("coper_score1n_rank_prepared", L.ptr(qp), L.ptr(Ep), L.ptr(tb), B, N, d, L.ptr(gold), L.ptr(tbitsT),
           L.ptr(ng), L.ptr(ne), p)
    assert torch.equal(ng, ng2) and torch.equal(ne, ne2)
    if N > 64:
        assert int(ne.sum().item()) > 0                         # the tie path was exercised
    # and against the oracle's counting rank on the stored logits
    if not lo:
        cnt, eq = O.rank_count(S[:, :N].cpu().numpy(), e2, filt.astype(np.float32))
        assert np.array_equal(ng.cpu().numpy() + 1, cnt) and np.array_equal(ne.cpu().numpy(), eq)
