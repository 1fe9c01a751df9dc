"""Entity-sharded execution of the hot path across the GPUs of one box.

Layout: the entity table ``ent_emb`` [N, d], ``pred_bias`` [N], their gradients and AMSGrad slots are
row-sharded by entity id (contiguous ``EntityShard`` ranges); every other variable is replicated.
Round-1 schedule ("replicated front end"): every rank receives the same global batch, runs the cheap
front end (lookups -> conv -> fused CPG-FC -> q) redundantly and bit-identically, and scores all B
queries against ITS entity rows only.  The exchange steps are the functions below — each is one NCCL
collective on a tiny tensor; dE / dbias / optimizer state of the table never leave their GPU:

  exchange_rows          all-reduce(sum) of the zero-masked [B, d] gather -> every rank has E[e1] exactly
  reduce_scorer_partials all-reduce(sum) of the partial loss (fp64) and the partial dq = G_shard . E_shard
  reduce_sharded_sumsq   all-reduce(sum) of the squared-norm partials of the sharded gradients (global-norm clip)
  reduce_gold_and_counts all-reduce(sum) of the owner-provided gold logits, then of the integer rank counts

All functions work on tensors of any device (NCCL on GPU; gloo on CPU in the unit tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class EntityShard:
    """Contiguous row range [lo, hi) of the entity table owned by ``rank`` (rows per rank rounded up to ``align``)."""

    def __init__(self, num_ent: int, rank: int = 0, world: int = 1, align: int = 128):
        per = -(-num_ent // world)
        per = -(-per // align) * align
        self.num_ent, self.rank, self.world, self.per = num_ent, rank, world, per
        self.lo = min(num_ent, rank * per)
        self.hi = min(num_ent, self.lo + per)

    @property
    def rows(self) -> int:
        return self.hi - self.lo

    def owner(self, ent: int) -> int:
        return int(ent) // self.per

    def __repr__(self):
        return "EntityShard(rank=%d/%d rows=[%d,%d) of %d)" % (self.rank, self.world, self.lo, self.hi, self.num_ent)


def _active(world, group):
    return world > 1 and dist.is_available() and dist.is_initialized()


def exchange_rows(x_masked: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """x_masked [B, d]: rows owned by this rank filled, all others exactly zero.  Adding zeros is exact in
    fp32, so one all-reduce(sum) delivers every row bit-exactly to every rank."""
    if _active(world, group):
        dist.all_reduce(x_masked, op=dist.ReduceOp.SUM, group=group)
    return x_masked


def reduce_scorer_partials(loss_sum: torch.Tensor, dq: torch.Tensor, world: int, group=None):
    """Each rank scored all queries against its rows: the loss and dq = sum over shards of G_s . E_s add up."""
    if _active(world, group):
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(dq, op=dist.ReduceOp.SUM, group=group)
    return loss_sum, dq


def reduce_sharded_sumsq(sumsq_sharded: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """Squared-norm partials of the row-sharded gradients add across ranks; replicated gradients are identical on
    every rank and are counted once (they are NOT passed here)."""
    if _active(world, group):
        dist.all_reduce(sumsq_sharded, op=dist.ReduceOp.SUM, group=group)
    return sumsq_sharded


def reduce_gold(gold: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """gold [B]: the logit of e2[b] on the rank that owns entity e2[b], 0 elsewhere."""
    if _active(world, group):
        dist.all_reduce(gold, op=dist.ReduceOp.SUM, group=group)
    return gold


def reduce_counts(n_greater: torch.Tensor, n_equal: torch.Tensor, world: int, group=None):
    """Integer partial counts add exactly -> ranks are bit-identical for any number of shards."""
    if _active(world, group):
        dist.all_reduce(n_greater, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(n_equal, op=dist.ReduceOp.SUM, group=group)
    return n_greater, n_equal
