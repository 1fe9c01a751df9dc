"""The fused generate-and-apply contraction as a ``torch.autograd.Function`` (last clause).

CoPER-MINERVA — the reference's PyTorch half — applies generated per-example weights with
``torch.einsum('ij,ijk->ik', X, pg_weights(Q)) + pg_bias(Q)`` (``CoPER_MINERVA/src/rl/graph_search/pn.py:125,132``,
``src/lstm_pg.py:166-169``, ``src/emb/fact_network.py:376-380,428``), where ``pg_weights(Q) = reshape(Q . P, [B, F, d])``
is materialised first (``B*F*d`` floats).  :func:`generate_and_apply` computes the same value and the same gradients
through ``coper_cpg_fc_fwd`` / ``coper_cpg_fc_bwd`` (include/coper.h) — the kernels of the ConvE path — without ever
forming the ``[B, F, d]`` weights:

    y[b] = sum_k context[b, k] * (x[b] . P[k])  +  context_bias[b] . P_bias          P[k] = P viewed [dc, F, d][k]

Engines: ``tf32x3`` / ``bf16`` (tcgen05; needs ``F % 32 == 0`` and ``d <= 256``) or ``fp32`` (CUDA cores, any shape);
``prec="auto"`` takes tf32x3 where the shape allows it.  There is no non-CUDA path.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr

PREC = _lib.PREC          # include/coper.h: COPER_PREC_FP32 = 0, COPER_PREC_BF16 = 1, COPER_PREC_TF32X3 = 2


def _workspace(nbytes, dev):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


class _GenerateAndApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, context, x, P, context_bias, P_bias, prec):
        lib = _lib.load()
        B, dc = context.shape
        F = x.shape[1]
        d = P.shape[1] // F
        dcb = context_bias.shape[1]
        y = torch.empty(B, d, dtype=torch.float32, device=x.device)
        ws = _workspace(lib.coper_cpg_fc_fwd_workspace_bytes(B, dc, F, d, prec), x.device)
        call("coper_cpg_fc_fwd", ptr(context), ptr(x), ptr(P), None, ptr(context_bias), ptr(P_bias), B, dc, F, d, dcb,
             1.0, None, 0, ptr(y), ptr(ws), ws.numel(), prec)
        ctx.save_for_backward(context, x, P, context_bias, P_bias)
        ctx.prec = prec
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        context, x, P, context_bias, P_bias = ctx.saved_tensors
        B, dc = context.shape
        F = x.shape[1]
        d = P.shape[1] // F
        dcb = context_bias.shape[1]
        dy = dy.contiguous().float()
        dP, dPb = torch.empty_like(P), torch.empty_like(P_bias)
        dx, dcontext, dcontext_b = torch.empty_like(x), torch.empty_like(context), torch.empty_like(context_bias)
        ws = _workspace(lib.coper_cpg_fc_bwd_workspace_bytes(B, dc, F, d, ctx.prec), x.device)
        call("coper_cpg_fc_bwd", ptr(context), ptr(x), ptr(P), None, ptr(context_bias), ptr(P_bias), ptr(dy), B, dc, F,
             d, dcb, ptr(dP), ptr(dPb), ptr(dx), ptr(dcontext), ptr(dcontext_b), ptr(ws), ws.numel(), ctx.prec, 0)
        return dcontext, dx, dP, dcontext_b, dPb, None


def generate_and_apply(context: torch.Tensor, x: torch.Tensor, P: torch.Tensor, P_bias: torch.Tensor = None,
                       context_bias: torch.Tensor = None, prec: str = "auto") -> torch.Tensor:
    """``einsum('ij,ijk->ik', x, reshape(context @ P, [B, F, d])) + context_bias @ P_bias`` with autograd support.

    context [B, dc]; x [B, F]; P [dc, F*d] (the generator's last projection, ``(f, j)`` row-major per context unit);
    P_bias [dcb, d] and context_bias [B, dcb] (default: ``context``) for the generated bias, omitted -> no bias term.
    All fp32 CUDA tensors."""
    for t in (context, x, P):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise _lib.CoperError("generate_and_apply needs fp32 CUDA tensors (no CPU path exists)")
    context, x, P = context.contiguous(), x.contiguous(), P.contiguous()
    B, dc = context.shape
    F = x.shape[1]
    if x.shape[0] != B or P.shape[0] != dc or P.shape[1] % F:
        raise ValueError("shapes: context [B, dc], x [B, F], P [dc, F*d]")
    d = P.shape[1] // F
    if P_bias is None:
        context_bias = context if context_bias is None else context_bias
        P_bias = torch.zeros(context_bias.shape[1], d, dtype=torch.float32, device=x.device)
    elif context_bias is None:
        context_bias = context
    context_bias, P_bias = context_bias.contiguous(), P_bias.contiguous()
    if context_bias.shape[0] != B or P_bias.shape != (context_bias.shape[1], d):
        raise ValueError("shapes: context_bias [B, dcb], P_bias [dcb, d]")
    tensor_ok = F % 32 == 0 and d <= 256
    if prec == "auto":
        prec = "tf32x3" if tensor_ok else "fp32"
    if prec != "fp32" and not tensor_ok:
        raise ValueError("the tcgen05 engines need F %% 32 == 0 and d <= 256 (got F=%d, d=%d); use prec='fp32'" % (F, d))
    return _GenerateAndApply.apply(context, x, P, context_bias, P_bias, PREC[prec])
