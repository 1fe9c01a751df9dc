"""World-size-2 (and 3) gloo tests, on CPU, of the entity-sharding exchange steps (coper_b200/sharding.py).

Each rank computes the shard-local quantities with the CPU oracle, runs the product's collective helpers over a
real process group, and checks that the sharded path reproduces the unsharded oracle: embedding rows and integer
rank counts bit-exactly, loss / dq / global norm to fp64 round-off."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import conve_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B):
    from coper_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = O.OracleConfig(num_ent=203, num_rel=6, ent_emb_size=40, rel_emb_size=5, context_rel_out=[])
        p = O.cast_params(O.init_params(cfg, seed=1, bias_noise=0.05), np.float64)
        e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, B, seed=2, mean_pos=5.0)
        dense = O.csr_to_dense(rowptr, col, cfg.num_ent, np.float64)
        full = O.forward(p, cfg, e1, rel, False, None, dense, np.float64)
        gfull = O.backward(full, cfg)
        sh = sharding.EntityShard(cfg.num_ent, rank, world, align=8)
        assert sum(sharding.EntityShard(cfg.num_ent, r, world, align=8).rows for r in range(world)) == cfg.num_ent
        E, bias = p["ent_emb"][sh.lo:sh.hi], p["pred_bias"][sh.lo:sh.hi]
        # 1. embedding rows: masked gather + all-reduce is bit-exact
        x = np.zeros((B, cfg.ent_emb_size), np.float32)
        own = (e1 >= sh.lo) & (e1 < sh.hi)
        x[own] = p["ent_emb"].astype(np.float32)[e1[own]]
        x = sharding.exchange_rows(torch.from_numpy(x), world).numpy()
        assert np.array_equal(x, p["ent_emb"].astype(np.float32)[e1])
        # 2. scorer partials
        q = full["q"]
        S = q @ E.T + bias
        zs = 0.9 * dense[:, sh.lo:sh.hi] + 1.0 / cfg.num_ent
        loss_sum = torch.tensor([(np.maximum(S, 0) - S * zs + np.log1p(np.exp(-np.abs(S)))).sum()])
        G = (O.sigmoid(S) - zs) / (B * cfg.num_ent)
        dq = torch.from_numpy(G @ E)
        sharding.reduce_scorer_partials(loss_sum, dq, world)
        assert abs(loss_sum.item() / (B * cfg.num_ent) - full["loss"]) < 1e-12
        assert np.abs(dq.numpy() - gfull["_dq"]).max() < 1e-15
        # local dE / dbias are exactly the owner's rows of the dense scorer gradient
        dE_local = G.T @ q
        dE_ref = gfull["ent_emb"].copy()
        np.subtract.at(dE_ref, e1, gfull["_dx0"])
        assert np.abs(dE_local - dE_ref[sh.lo:sh.hi]).max() < 1e-15
        # 3. global norm: sharded squared norms add, replicated ones are counted once
        ss = torch.tensor([(gfull["ent_emb"][sh.lo:sh.hi] ** 2).sum(), (gfull["pred_bias"][sh.lo:sh.hi] ** 2).sum()])
        sharding.reduce_sharded_sumsq(ss, world)
        assert abs(ss[0].item() - (gfull["ent_emb"] ** 2).sum()) < 1e-18
        # 4. filtered rank: owner provides the gold logit, integer counts add exactly
        S32 = S.astype(np.float32)
        l = e2 - sh.lo
        mine = (l >= 0) & (l < sh.rows)
        gold = np.zeros(B, np.float32)
        gold[mine] = S32[np.arange(B)[mine], l[mine]]
        gold = sharding.reduce_gold(torch.from_numpy(gold), world).numpy()
        Sfull32 = full["scores"].astype(np.float32)
        assert np.array_equal(gold, Sfull32[np.arange(B), e2])
        filt = dense[:, sh.lo:sh.hi] == 1
        valid = ~filt
        valid[np.arange(B)[mine], l[mine]] = False
        ng = torch.from_numpy(((S32 > gold[:, None]) & valid).sum(1).astype(np.int32))
        ne = torch.from_numpy(((S32 == gold[:, None]) & valid).sum(1).astype(np.int32))
        sharding.reduce_counts(ng, ne, world)
        rc, eq = O.rank_count(Sfull32, e2, dense)
        assert np.array_equal(ng.numpy() + 1, rc) and np.array_equal(ne.numpy(), eq)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 16), (3, 7)])
def test_entity_sharded_exchange_steps_gloo(world, B):
    mp.spawn(_worker, args=(world, _free_port(), B), nprocs=world, join=True)
