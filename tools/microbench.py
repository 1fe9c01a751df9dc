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


if __name__ == "__main__":
    score(512, 40943, 200)
    score(512, 1000000, 256)
    tcgemm(40943, 200, 512, 1, 0)     # dE-like
    tcgemm(4608, 200, 512, 1, 0)      # dP-like
