"""Dev tool: pipeline timeline of CTA 0 of the fused scorer kernel (clock64 stamps per entity tile)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coper_b200 import _lib as L
lib = L.load()
B, N, d = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 1250000, int(sys.argv[2]) if len(sys.argv) > 2 else 256
q = torch.randn(B, d, device="cuda").clamp_(min=0)
E = (torch.rand(N, d, device="cuda") - 0.5) * 0.1
bias = torch.zeros(N, device="cuda")
ld = -(-N // 32) * 32
bits_t = torch.zeros(N, -(-B // 32), dtype=torch.int32, device="cuda")
loss = torch.zeros(1, dtype=torch.float64, device="cuda")
dq, dE, db = torch.zeros(B, d, device="cuda"), torch.zeros(N, d, device="cuda"), torch.zeros(N, device="cuda")
p = 1
ws = torch.empty(lib.coper_score1n_bce_workspace_bytes(B, N, d, p), dtype=torch.uint8, device="cuda")
G = torch.empty(lib.coper_score1n_bce_G_bytes(B, N, p), dtype=torch.uint8, device="cuda")
Ep = torch.empty(lib.coper_prepared_bytes(N, d, p), dtype=torch.uint8, device="cuda")
L.call("coper_prepare_operand", L.ptr(E), N, d, d, p, L.ptr(Ep))
tiles = 512
trace = torch.zeros(tiles, 8, dtype=torch.int64, device="cuda")
lib.coper_debug_set_fused_trace.argtypes = [ctypes.c_void_p]
lib.coper_debug_set_fused_trace.restype = None
def run():
    L.call("coper_score1n_bce_fwd_bwd", L.ptr(q), L.ptr(E), L.ptr(Ep), L.ptr(bias), L.ptr(bits_t), B, N, d,
           0.9, 1.0 / N, 1.0 / (B * N), L.ptr(loss), L.ptr(G), ld, L.ptr(dq), L.ptr(dE), L.ptr(db), L.ptr(ws), ws.numel(), p)
run(); run(); torch.cuda.synchronize()
lib.coper_debug_set_fused_trace(trace.data_ptr())
run(); torch.cuda.synchronize()
lib.coper_debug_set_fused_trace(None)
t = trace.cpu().numpy()
names = ["S_issue", "S_issued", "G_seen", "DQ_issued", "E_free_seen", "S_seen(epi)", "epi_math_done", "G_written"]
t0 = t[0][0]
print("tile " + " ".join("%13s" % n for n in names))
for i in list(range(0, 12)) + list(range(40, 46)):
    if t[i][0] == 0:
        break
    print("%4d " % i + " ".join("%13d" % (x - t0 if x else -1) for x in t[i]))
nz = [i for i in range(tiles) if t[i][0]]
if len(nz) > 10:
    per = (t[nz[-1]][0] - t[nz[5]][0]) / (nz[-1] - nz[5])
    print("steady-state period: %.0f clk per tile over tiles %d..%d" % (per, nz[5], nz[-1]))
    import numpy as np
    a = t[nz[5]:nz[-1]]
    print("mean S issue->issued %.0f | S issued->epi sees S %.0f | epi: S seen->math done %.0f | math done->G written %.0f | "
          "G written->MMA sees G %.0f | G seen->dq issued %.0f" % (
              np.mean(a[:, 1] - a[:, 0]), np.mean(a[:, 5] - a[:, 1]), np.mean(a[:, 6] - a[:, 5]), np.mean(a[:, 7] - a[:, 6]),
              np.mean(a[:, 2] - a[:, 7]), np.mean(a[:, 3] - a[:, 2])))
    print("mean E buffer free seen (tile i) - DQ issued (tile i-2): %.0f" % np.mean(a[2:, 4] - a[:-2, 3]))
    print("mean S issue (tile i) - E free seen (tile i): %.0f" % np.mean(a[:, 0] - a[:, 4]))
