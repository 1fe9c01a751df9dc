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
    """A port below the ephemeral range (so no outgoing connection grabs it between the probe and the rendezvous)."""
    import random
    rng = random.Random(os.getpid() ^ int.from_bytes(os.urandom(4), "little"))
    for _ in range(64):
        p = rng.randrange(15000, 30000)
        s = socket.socket()
        try:
            s.bind(("127.0.0.1", p))
            return p
        except OSError:
            continue
        finally:
            s.close()
    raise RuntimeError("no free rendezvous port")


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


def _dp_worker(rank, world, port, Bl):
    """Data-parallel front end exchanges: every rank owns rows [r*Bl, (r+1)*Bl) of the global batch."""
    from coper_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Bg = Bl * world
        cfg = O.OracleConfig(num_ent=203, num_rel=6, ent_emb_size=40, rel_emb_size=5, context_rel_out=[],
                             batch_norm_train_stats=True)
        p = O.cast_params(O.init_params(cfg, seed=1, bias_noise=0.05), np.float64)
        e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, Bg, seed=2, mean_pos=5.0)
        e1[Bg // 2:] = e1[:Bg - Bg // 2]                  # head entities shared across the ranks' slices
        dense = O.csr_to_dense(rowptr, col, cfg.num_ent, np.float64)
        full = O.forward(p, cfg, e1, rel, True, None, dense, np.float64)
        gfull = O.backward(full, cfg)
        sh = sharding.EntityShard(cfg.num_ent, rank, world, align=8)
        mine = slice(rank * Bl, (rank + 1) * Bl)
        d = cfg.ent_emb_size
        # 1. masked gather of ALL Bg head rows + reduce-scatter: this rank's rows, bit-exact
        E32 = p["ent_emb"].astype(np.float32)
        x = np.zeros((Bg, d), np.float32)
        own = (e1 >= sh.lo) & (e1 < sh.hi)
        x[own] = E32[e1[own]]
        xl = sharding.scatter_rows(torch.from_numpy(x), torch.zeros(Bl, d), world).numpy()
        assert np.array_equal(xl, E32[e1[mine]])
        # 2. synchronised batch norm: chunk partials {sum, sum of squares} of the local rows, gathered, finalised over
        #    P*nchunk chunks and P*R rows == statistics of the global batch
        y = full["y"] if "y" in full else full["q"]
        yl = y[mine]
        nch = 2
        parts = np.zeros((nch, y.shape[1], 2))
        for k, rows in enumerate(np.array_split(np.arange(Bl), nch)):
            parts[k, :, 0], parts[k, :, 1] = yl[rows].sum(0), (yl[rows] ** 2).sum(0)
        allp = sharding.gather_stat_partials(torch.zeros(world * nch, y.shape[1], 2, dtype=torch.float64),
                                             torch.from_numpy(parts), world).numpy()
        mean = allp[:, :, 0].sum(0) / Bg
        var = allp[:, :, 1].sum(0) / Bg - mean ** 2
        assert np.abs(mean - y.mean(0)).max() < 1e-12 and np.abs(var - y.var(0)).max() < 1e-12
        # 3. q rows gathered into batch order
        qg = sharding.gather_batch(torch.zeros(Bg, d, dtype=torch.float64), torch.from_numpy(full["q"][mine].copy()),
                                   world).numpy()
        assert np.array_equal(qg, full["q"])
        # 4. scorer over all Bg queries against the local entity rows: partial dq reduce-scattered to the row owners
        E, bias = p["ent_emb"][sh.lo:sh.hi], p["pred_bias"][sh.lo:sh.hi]
        S = qg @ E.T + bias
        zs = 0.9 * dense[:, sh.lo:sh.hi] + 1.0 / cfg.num_ent
        loss_sum = torch.tensor([(np.maximum(S, 0) - S * zs + np.log1p(np.exp(-np.abs(S)))).sum()])
        G = (O.sigmoid(S) - zs) / (Bg * cfg.num_ent)
        dql = sharding.scatter_dq(torch.from_numpy(G @ E), torch.zeros(Bl, d, dtype=torch.float64), loss_sum,
                                  world).numpy()
        assert abs(loss_sum.item() / (Bg * cfg.num_ent) - full["loss"]) < 1e-12
        assert np.abs(dql - gfull["_dq"][mine]).max() < 1e-15
        # 5. replicated-parameter gradients are sums over queries: the rank's rows, then one bucketed all-reduce
        #    (rel_emb: IndexedSlices made dense + the per-row sums of squared slices, in the same bucket)
        drl = gfull["_dr"][mine]
        flat = torch.zeros(2, cfg.num_rel, cfg.rel_emb_size, dtype=torch.float64)
        np.add.at(flat[0].numpy(), rel[mine], drl)
        np.add.at(flat[1].numpy(), rel[mine], drl ** 2)
        sharding.reduce_replicated_grads(flat.view(-1), world)
        assert np.abs(flat[0].numpy() - gfull["rel_emb"]).max() < 1e-15
        sq = np.zeros_like(gfull["rel_emb"])
        np.add.at(sq, rel, gfull["_dr"] ** 2)
        assert np.abs(flat[1].numpy() - sq).max() < 1e-18
        # 6. dx0 of all queries gathered; each shard scatters the rows whose head entity it owns
        dx0 = sharding.gather_batch(torch.zeros(Bg, d, dtype=torch.float64), torch.from_numpy(gfull["_dx0"][mine].copy()),
                                    world).numpy()
        dE = G.T @ qg
        for b in np.nonzero(own)[0]:
            dE[e1[b] - sh.lo] += dx0[b]
        assert np.abs(dE - gfull["ent_emb"][sh.lo:sh.hi]).max() < 1e-15
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Bl", [(2, 8), (3, 5)])
def test_data_parallel_exchange_steps_gloo(world, Bl):
    mp.spawn(_dp_worker, args=(world, _free_port(), Bl), nprocs=world, join=True)
