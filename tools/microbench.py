"""Isolated CUDA-event timings of single C-ABI calls (dev tool; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coper_b200 import _lib as L

lib = L.load()


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def score(B, N, d):
    q = torch.randn(B, d, device="cuda").clamp_(min=0)
    E = (torch.rand(N, d, device="cuda") - 0.5) * 0.1
    bias = torch.zeros(N, device="cuda")
    ld = -(-N // 32) * 32
    S = torch.empty(B, ld, device="cuda")
    fl = 2.0 * B * N * d
    t = timeit(lambda: L.call("coper_score1n_fwd", L.ptr(q), L.ptr(E), L.ptr(bias), B, N, d, L.ptr(S), ld, None, 0, 0))
    print("score1n_fwd  B=%d N=%d d=%d  simt-fp32 %.3f ms %.1f TF/s" % (B, N, d, t, fl / t / 1e9))
    for name, p in (("bf16", 1), ("tf32x3", 2)):
        qp = torch.empty(lib.coper_prepared_bytes(B, d, p), dtype=torch.uint8, device="cuda")
        Ep = torch.empty(lib.coper_prepared_bytes(N, d, p), dtype=torch.uint8, device="cuda")
        tp = timeit(lambda: L.call("coper_prepare_operand", L.ptr(E), N, d, d, p, L.ptr(Ep)))
        L.call("coper_prepare_operand", L.ptr(q), B, d, d, p, L.ptr(qp))
        t = timeit(lambda: L.call("coper_score1n_fwd_prepared", L.ptr(qp), L.ptr(Ep), L.ptr(bias), B, N, d, L.ptr(S), ld, p))
        print("   %-7s prepared %.3f ms %.1f TF/s (prepare E %.3f ms; out write %.0f GB/s)" % (
            name, t, fl / t / 1e9, tp, B * N * 4 / t / 1e6))


def tcgemm(M, N, K, ta, tb):
    A = torch.randn((K, M) if ta else (M, K), device="cuda")
    Bm = torch.randn((N, K) if tb else (K, N), device="cuda")
    C = torch.empty(M, N, device="cuda")
    fl = 2.0 * M * N * K
    for name, p in (("bf16", 1), ("tf32x3", 2)):
        ws = torch.empty(lib.coper_tc_gemm_workspace_bytes(M, N, K, p), dtype=torch.uint8, device="cuda")
        t = timeit(lambda: L.call("coper_tc_gemm", ta, tb, M, N, K, L.ptr(A), A.shape[1], L.ptr(Bm), Bm.shape[1],
                                  L.ptr(C), N, p, L.ptr(ws), ws.numel()))
        print("tc_gemm M=%d N=%d K=%d ta=%d tb=%d %-7s %.3f ms %.1f TF/s (incl. operand prep)" % (M, N, K, ta, tb, name, t, fl / t / 1e9))
    t = timeit(lambda: L.call("coper_sgemm", ta, tb, M, N, K, L.ptr(A), A.shape[1], L.ptr(Bm), Bm.shape[1], L.ptr(C), N, 0))
    print("   simt sgemm %.3f ms %.1f TF/s" % (t, fl / t / 1e9))


def bce(B, N, d, precs=("fp32", "bf16", "tf32x3")):
    q = torch.randn(B, d, device="cuda").clamp_(min=0)
    E = (torch.rand(N, d, device="cuda") - 0.5) * 0.1
    bias = torch.zeros(N, device="cuda")
    ld = -(-N // 32) * 32
    bits_q = torch.zeros(B, ld // 32, dtype=torch.int32, device="cuda")
    bits_q[:, ::97] = 5
    bits_t = torch.zeros(N, -(-B // 32), dtype=torch.int32, device="cuda")
    bits_t[::97] = 5
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    dq, dE, db = torch.zeros(B, d, device="cuda"), torch.zeros(N, d, device="cuda"), torch.zeros(N, device="cuda")
    fl = 6.0 * B * N * d
    for name in precs:
        p = L.PREC[name]
        ws = torch.empty(lib.coper_score1n_bce_workspace_bytes(B, N, d, p), dtype=torch.uint8, device="cuda")
        G = torch.empty(lib.coper_score1n_bce_G_bytes(B, N, p), dtype=torch.uint8, device="cuda")
        bits = bits_q if p == 0 else bits_t
        t = timeit(lambda: L.call("coper_score1n_bce_fwd_bwd", L.ptr(q), L.ptr(E), None, L.ptr(bias), L.ptr(bits), B, N, d,
                                  0.9, 1.0 / N, 1.0 / (B * N), L.ptr(loss), L.ptr(G), ld, L.ptr(dq), L.ptr(dE),
                                  L.ptr(db), L.ptr(ws), ws.numel(), p))
        print("score1n_bce_fwd_bwd B=%d N=%d d=%d %-7s %.3f ms %.1f TF/s" % (B, N, d, name, t, fl / t / 1e9))


def rankf(B, N, d, precs=("bf16", "tf32x3", "fp16x3")):
    q = torch.randn(B, d, device="cuda").clamp_(min=0)
    E = (torch.rand(N, d, device="cuda") - 0.5) * 0.1
    bias = torch.zeros(N, device="cuda")
    ld = -(-N // 32) * 32
    bits = torch.zeros(N, -(-B // 32), dtype=torch.int32, device="cuda")
    e2 = torch.randint(0, N, (B,), device="cuda")
    gold = torch.zeros(B, device="cuda")
    ng, ne = torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")
    fl = 2.0 * B * N * d
    for name in precs:
        p = L.PREC[name]
        qp = torch.empty(lib.coper_prepared_bytes(B, d, p), dtype=torch.uint8, device="cuda")
        Ep = torch.empty(lib.coper_prepared_bytes(N, d, p), dtype=torch.uint8, device="cuda")
        ws = torch.empty(max(256, lib.coper_score1n_rank_workspace_bytes(B, d, p)), dtype=torch.uint8, device="cuda")
        L.call("coper_prepare_operand", L.ptr(E), N, d, d, p, L.ptr(Ep))
        L.call("coper_prepare_operand", L.ptr(q), B, d, d, p, L.ptr(qp))
        t0 = timeit(lambda: L.call("coper_score1n_gold_prepared", L.ptr(qp), L.ptr(Ep), L.ptr(bias), B, N, d, L.ptr(e2), 0,
                                   L.ptr(gold), L.ptr(ws), ws.numel(), p))
        t = timeit(lambda: L.call("coper_score1n_rank_prepared", L.ptr(qp), L.ptr(Ep), L.ptr(bias), B, N, d,
                                  L.ptr(gold), L.ptr(bits), L.ptr(ng), L.ptr(ne), p))
        print("score1n_rank_fused B=%d N=%d d=%d %-7s gold %.3f ms, rank %.3f ms %.1f TF/s" % (B, N, d, name, t0, t, fl / t / 1e9))


def cpg(B, dc, F, d, precs=("fp32", "bf16", "tf32x3", "fp16x3")):
    c, f = torch.randn(B, dc, device="cuda"), torch.randn(B, F, device="cuda").clamp_(min=0)
    P, Pb = torch.randn(dc, F * d, device="cuda") * 0.01, torch.randn(dc, d, device="cuda")
    dy = torch.randn(B, d, device="cuda")
    y, df, dcw, dcb = (torch.zeros(s, device="cuda") for s in ((B, d), (B, F), (B, dc), (B, dc)))
    dP, dPb = torch.zeros_like(P), torch.zeros_like(Pb)
    fl = 2.0 * B * dc * F * d
    for name in precs:
        p = L.PREC[name]
        ws = torch.empty(max(lib.coper_cpg_fc_fwd_workspace_bytes(B, dc, F, d, p),
                             lib.coper_cpg_fc_bwd_workspace_bytes(B, dc, F, d, p)), dtype=torch.uint8, device="cuda")
        t1 = timeit(lambda: L.call("coper_cpg_fc_fwd", L.ptr(c), L.ptr(f), L.ptr(P), None, L.ptr(c), L.ptr(Pb), B, dc, F, d, dc,
                                   1.0, None, 0, L.ptr(y), L.ptr(ws), ws.numel(), p))
        t2 = timeit(lambda: L.call("coper_cpg_fc_bwd", L.ptr(c), L.ptr(f), L.ptr(P), None, L.ptr(c), L.ptr(Pb), L.ptr(dy), B, dc,
                                   F, d, dc, L.ptr(dP), L.ptr(dPb), L.ptr(df), L.ptr(dcw), L.ptr(dcb), L.ptr(ws),
                                   ws.numel(), p, int(p != 0)))
        print("cpg_fc B=%d dc=%d F=%d d=%d %-7s fwd %.3f ms %.1f TF/s | bwd %.3f ms %.1f TF/s" % (
            B, dc, F, d, name, t1, fl / t1 / 1e9, t2, 2 * fl / t2 / 1e9))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cpg":
        cpg(512, 8, 4608, 200)
        cpg(512, 32, 4608, 200)
        cpg(512, 32, 6272, 256)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "prof":
        bce(512, 40943, 200, precs=(sys.argv[2],))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rank":
        precs = tuple(sys.argv[2:]) or ("bf16", "tf32x3", "fp16x3")
        rankf(512, 40943, 200, precs)
        rankf(512, 1000000, 256, precs)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bcebig":
        bce(512, 1250000, 256, precs=(sys.argv[2],))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bce":
        bce(512, 40943, 200)
        bce(512, 1250000, 256, precs=("bf16", "tf32x3"))
        sys.exit(0)
    score(512, 40943, 200)
    score(512, 1000000, 256)
    tcgemm(40943, 200, 512, 1, 0)     # dE-like
    tcgemm(4608, 200, 512, 1, 0)      # dP-like
