"""Entity-sharded CUDA path across 2 ranks: the sharded model must reproduce the unsharded oracle — loss to fp32
round-off, local rows of dE / updated table, and filtered ranks bit-exactly equal to the single-GPU ranks.
With >= 2 visible devices the ranks sit on different GPUs and talk NCCL; on a ONE-GPU box the two processes share
cuda:0 and talk gloo (collectives staged through host memory, coper_b200/sharding.py) - the same kernels, exchange
steps, synchronised batch norm and sharded rank reduction run either way, so the single-GPU test tier covers them."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import conve_oracle as O

pytestmark = pytest.mark.gpu


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


def _init_dist(rank, world, port):
    """NCCL with one device per rank when the box has them, else two processes on cuda:0 over gloo."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    ndev = torch.cuda.device_count()
    dev = rank if ndev >= world else 0
    torch.cuda.set_device(dev)
    if ndev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return dev


def _descr(cfg):
    return {"use_negative_sampling": False, "label_smoothing_epsilon": 0.1, "num_ent": cfg.num_ent,
            "num_rel": cfg.num_rel, "ent_emb_size": cfg.ent_emb_size, "rel_emb_size": cfg.rel_emb_size,
            "concat_rel": False, "context_rel_conv": None, "context_rel_out": [], "context_rel_dropout": 0.0,
            "context_rel_use_batch_norm": False, "input_dropout": 0.0, "hidden_dropout": cfg.hidden_dropout,
            "output_dropout": cfg.output_dropout, "learning_rate": 1e-2, "batch_size": 0, "add_loss_summaries": False,
            "add_variable_summaries": False, "add_tensor_summaries": False,
            "batch_norm_momentum": cfg.batch_norm_momentum, "batch_norm_train_stats": cfg.batch_norm_train_stats,
            "do_parameter_lookup": False}


def _worker(rank, world, port, out_dir, prec):
    import torch.distributed as dist
    from coper_b200.models import ConvE
    from coper_b200.sharding import EntityShard
    dev = _init_dist(rank, world, port)
    try:
        cfg = O.OracleConfig(num_ent=1003, num_rel=22, ent_emb_size=200, rel_emb_size=8, context_rel_out=[],
                             batch_norm_train_stats=True, batch_norm_momentum=0.1, hidden_dropout=0.3,
                             output_dropout=0.2)
        params = O.init_params(cfg, seed=3, bias_noise=0.05)
        B = 130
        e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, B, seed=5, mean_pos=5.0)
        batch = {"e1": e1, "rel": rel, "e2": e2, "e2_multi_rowptr": rowptr, "e2_multi_col": col}
        sh = EntityShard(cfg.num_ent, rank, world)
        m = ConvE(_descr(cfg), device="cuda:%d" % dev, seed=0, shard=sh, prec=prec)
        m.load_variables(params)
        ranks, n_equal = m.filtered_ranks(batch)
        S_local = m.predict_all(batch).cpu().numpy().copy()
        loss = float(m.train_step(batch).item())
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ranks=ranks.cpu().numpy(), n_equal=n_equal.cpu().numpy(),
                 S=S_local, loss=loss, dE=m.grads["ent_emb"].cpu().numpy(), ent=m.ent_emb.cpu().numpy(),
                 rel_emb=m.rel_emb.cpu().numpy(), P=m.fc_weights.projections[0].cpu().numpy(), lo=sh.lo, hi=sh.hi,
                 norm=float(m.clip_out[1].item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("prec", ["fp32", "tf32x3", "fp16x3", "bf16"])
@pytest.mark.parametrize("world", [2])
def test_entity_sharded_matches_single_gpu(world, prec, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), prec), nprocs=world, join=True)
    _worker_single(str(tmp_path), prec)
    ref = np.load(os.path.join(str(tmp_path), "single.npz"))
    outs = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    S = np.concatenate([o["S"] for o in outs], axis=1)
    assert np.array_equal(S, ref["S"])                               # same kernels, same rows -> bit-identical logits
    for o in outs:
        assert np.array_equal(o["ranks"], ref["ranks"]) and np.array_equal(o["n_equal"], ref["n_equal"])
        assert abs(o["loss"] - float(ref["loss"])) < 2e-6 * abs(float(ref["loss"]))
        assert abs(o["norm"] - float(ref["norm"])) < 1e-5 * float(ref["norm"])
        lo, hi = int(o["lo"]), int(o["hi"])
        assert np.abs(o["dE"] - ref["dE"][lo:hi]).max() < (1e-5 if prec != "bf16" else 1e-3) * np.abs(ref["dE"]).max()
        if prec == "bf16":      # bf16 dL/dS: the update direction is compared by the single-GPU bf16 tests
            continue
        # updated variables: AMSGrad as written steps by ~ g / sqrt(g^2), which amplifies fp32 summation-order noise
        # on near-zero gradient entries -> compare at 5e-4 (the gradients themselves are compared at 1e-5 above)
        assert np.abs(o["ent"] - ref["ent"][lo:hi]).max() < 5e-4 * np.abs(ref["ent"]).max()
        assert np.abs(o["rel_emb"] - ref["rel_emb"]).max() < 5e-4 * np.abs(ref["rel_emb"]).max()
        assert np.abs(o["P"] - ref["P"]).max() < 5e-4 * np.abs(ref["P"]).max()
    # replicated variables stay bit-identical across ranks without any gradient all-reduce
    assert np.array_equal(outs[0]["P"], outs[1]["P"]) and np.array_equal(outs[0]["rel_emb"], outs[1]["rel_emb"])


def _worker_single(out_dir, prec):
    from coper_b200.models import ConvE
    cfg = O.OracleConfig(num_ent=1003, num_rel=22, ent_emb_size=200, rel_emb_size=8, context_rel_out=[],
                         batch_norm_train_stats=True, batch_norm_momentum=0.1, hidden_dropout=0.3, output_dropout=0.2)
    params = O.init_params(cfg, seed=3, bias_noise=0.05)
    B = 130
    e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, B, seed=5, mean_pos=5.0)
    batch = {"e1": e1, "rel": rel, "e2": e2, "e2_multi_rowptr": rowptr, "e2_multi_col": col}
    m = ConvE(_descr(cfg), device="cuda:0", seed=0, prec=prec)
    m.load_variables(params)
    ranks, n_equal = m.filtered_ranks(batch)
    S = m.predict_all(batch).cpu().numpy().copy()
    loss = float(m.train_step(batch).item())
    np.savez(os.path.join(out_dir, "single.npz"), ranks=ranks.cpu().numpy(), n_equal=n_equal.cpu().numpy(), S=S,
             loss=loss, dE=m.grads["ent_emb"].cpu().numpy(), ent=m.ent_emb.cpu().numpy(),
             rel_emb=m.rel_emb.cpu().numpy(), P=m.fc_weights.projections[0].cpu().numpy(),
             norm=float(m.clip_out[1].item()))


# ---------------------------------------------------------------------------------------------------------------
# Data-parallel front end (weak scaling): every rank runs conv / CPG / FC on its own slice of the GLOBAL batch with
# synchronised batch-norm statistics, the scorer stays entity-sharded.  One step on P ranks over Bg queries must be
# the oracle's step over the same Bg queries (masks: each rank's own dropout draw, assembled in batch order).
DP_CFG = dict(num_ent=1003, num_rel=22, ent_emb_size=200, rel_emb_size=8, context_rel_out=[12],
              context_rel_use_batch_norm=True, context_rel_dropout=0.1, batch_norm_train_stats=True,
              batch_norm_momentum=0.1, hidden_dropout=0.3, output_dropout=0.2)
DP_B = 132


def _dp_worker(rank, world, port, out_dir, prec):
    import torch.distributed as dist
    from coper_b200.models import ConvE
    from coper_b200.sharding import EntityShard
    from test_gpu_model import descriptors, export_masks
    dev = _init_dist(rank, world, port)
    try:
        cfg = O.OracleConfig(**DP_CFG)
        params = O.init_params(cfg, seed=3, bias_noise=0.05)
        e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, DP_B, seed=5, mean_pos=5.0)
        e1[DP_B // 2:] = e1[:DP_B - DP_B // 2]           # heads shared ACROSS the two ranks' slices
        batch = {"e1": e1, "rel": rel, "e2": e2, "e2_multi_rowptr": rowptr, "e2_multi_col": col}
        sh = EntityShard(cfg.num_ent, rank, world)
        m = ConvE(descriptors(cfg, 1e-2), device="cuda:%d" % dev, seed=0, shard=sh, prec=prec, data_parallel=True)
        m.load_variables(params)
        ranks, n_equal = m.filtered_ranks(batch)
        loss = float(m.train_step(batch, apply_update=False).item())
        Bl = DP_B // world
        masks = export_masks(m, cfg, Bl)
        g = {k.replace("/", "__"): v.cpu().numpy() for k, v in m.grads.items()}
        bl = m._bufs[("dp", DP_B)]
        q1, dx01, mm1 = bl.q.cpu().numpy(), bl.dx0.cpu().numpy(), m.conv1_bn.moving_mean.cpu().numpy()
        extra = {"mask_" + k: (np.stack(v) if isinstance(v, list) else v) for k, v in masks.items()}
        m._clip_and_apply()
        loss2 = float(m.train_step(batch).item())          # second step through the full (update) path
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, "dp%d.npz" % rank), ranks=ranks.cpu().numpy(), n_equal=n_equal.cpu().numpy(),
                 loss=loss, loss2=loss2, q=q1, dx0=dx01, lo=sh.lo, hi=sh.hi,
                 norm=float(m.clip_out[1].item()), rel_emb_new=m.rel_emb.cpu().numpy(),
                 P_new=m.fc_weights.projections[-1].cpu().numpy(), conv_mm=mm1,
                 **g, **extra)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("prec", ["fp32", "tf32x3", "fp16x3"])
def test_data_parallel_front_end_matches_oracle(prec, tmp_path):
    world = 2
    import torch.multiprocessing as mp
    mp.spawn(_dp_worker, args=(world, _free_port(), str(tmp_path), prec), nprocs=world, join=True)
    outs = [np.load(os.path.join(str(tmp_path), "dp%d.npz" % r)) for r in range(world)]
    cfg = O.OracleConfig(**DP_CFG)
    params = O.init_params(cfg, seed=3, bias_noise=0.05)
    e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, DP_B, seed=5, mean_pos=5.0)
    e1[DP_B // 2:] = e1[:DP_B - DP_B // 2]
    dense = O.csr_to_dense(rowptr, col, cfg.num_ent)
    masks = {"feature_map": np.concatenate([o["mask_feature_map"] for o in outs], 0),
             "output": np.concatenate([o["mask_output"] for o in outs], 0)}
    for key in ("ctx_w", "ctx_b"):
        masks[key] = list(np.concatenate([o["mask_" + key] for o in outs], 1))      # [layers, B, n]
    out = O.forward(params, cfg, e1, rel, True, masks, dense, np.float64)
    g = O.backward(out, cfg)

    def rel_err(a, b, floor=0.0):
        return np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-30)
    q = np.concatenate([o["q"] for o in outs], 0)
    assert rel_err(q, out["q"]) < 1e-5
    assert rel_err(np.concatenate([o["dx0"] for o in outs], 0), g["_dx0"]) < 2e-4
    scale = max(np.abs(g[k]).max() for k in ("ent_emb", "rel_emb", "conv1_weights"))
    nw = len(g["fc_weights_proj"])
    for o in outs:
        lo, hi = int(o["lo"]), int(o["hi"])
        assert abs(float(o["loss"]) - out["loss"]) < 1e-6 * abs(out["loss"])
        assert rel_err(o["ent_emb"], g["ent_emb"][lo:hi]) < 2e-4
        assert rel_err(o["pred_bias"], g["pred_bias"][lo:hi]) < 2e-4
        assert rel_err(o["rel_emb"].reshape(g["rel_emb"].shape), g["rel_emb"]) < 2e-4
        assert rel_err(o["conv1_weights"].reshape(g["conv1_weights"].shape), g["conv1_weights"]) < 2e-4
        for nm in ("FCBN", "Conv1BN"):
            for k in ("gamma", "beta"):
                assert rel_err(o["%s__%s" % (nm, k)], g[nm][k], 1e-4 * scale) < 2e-4, (nm, k)
        for i in range(nw):
            assert rel_err(o["fc_weights__CPG__Projection%d" % i].reshape(g["fc_weights_proj"][i].shape),
                           g["fc_weights_proj"][i], 1e-4 * scale) < 2e-4
            assert rel_err(o["fc_bias__CPG__Projection%d" % i].reshape(g["fc_bias_proj"][i].shape),
                           g["fc_bias_proj"][i], 1e-4 * scale) < 2e-4
        for i in range(nw - 1):
            assert rel_err(o["fc_weights__CPG__Projection%d__BatchNorm__gamma" % i], g["fc_weights_bn"][i]["gamma"],
                           1e-4 * scale) < 2e-4
        mm, _ = out["moving"]["Conv1BN"]
        assert rel_err(o["conv_mm"], mm) < 1e-5
        assert np.isfinite(float(o["loss2"])) and float(o["loss2"]) < float(o["loss"]) * 1.5
    # evaluation: every rank holds the global ranks; the oracle ranks its own fp64 logits -> compare where no
    # near-ties exist (|score gap| checks live in the single-GPU tests); here: ranks agree between the two ranks and
    # with the oracle on >= 95 % of queries (rounding can flip near-tied neighbours)
    assert np.array_equal(outs[0]["ranks"], outs[1]["ranks"]) and np.array_equal(outs[0]["n_equal"], outs[1]["n_equal"])
    ev = O.forward(params, cfg, e1, rel, False, None, None, np.float64)
    cnt, _ = O.rank_count(ev["scores"], e2, dense)
    assert (outs[0]["ranks"] == cnt).mean() >= 0.95
    # replicated variables identical on both ranks after the bucketed all-reduce + update
    assert np.array_equal(outs[0]["rel_emb_new"], outs[1]["rel_emb_new"])
    assert np.array_equal(outs[0]["P_new"], outs[1]["P_new"])


def test_run_cpg_entry_point_under_torchrun(tmp_path):
    """The train / evaluate entry point launched one process per GPU: entity-sharded table, batch split over the ranks,
    rank 0 writes the config and the (all-gathered) embedding pickle, every rank its checkpoint shard; the saved shards
    restore and evaluate."""
    world = 2
    import glob
    import pickle
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    wd = str(tmp_path)
    base = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
            "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), "-m", "coper_b200.run_cpg",
            "--synthetic", "toy", "--working-dir", wd, "--eval-batches", "2", "--prec", "tf32x3"]
    out = subprocess.run(base + ["--max-steps", "12"], cwd=root, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "Step      0 | Loss" in out.stderr and "MRR" in out.stderr
    shards = sorted(glob.glob(os.path.join(wd, "checkpoints", "*", "model_weights.ckpt", "model_weights.ckpt.rank*")))
    assert len(shards) == world
    emb = glob.glob(os.path.join(wd, "evaluation", "*", "best_embeddings.ckpt"))
    rel_emb, ent_emb = pickle.load(open(emb[0], "rb"))
    assert ent_emb.shape == (997, 40) and rel_emb.shape == (6, 5) and np.abs(ent_emb[-1]).max() > 0
    out = subprocess.run(base + ["--model-load-path", shards[0][:-len(".rank0")]], cwd=root, capture_output=True,
                         text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "test" in out.stderr and "MRR" in out.stderr
