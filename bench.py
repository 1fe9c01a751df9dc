#!/usr/bin/env python
"""bench.py — CoPER-ConvE hot-path throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--shape wn18rr] [--prec fp32] [--impl ours|reference]

A "step" = one full training step of the hot path over one synthetic batch of B (e1, rel) queries with
1-N labels: lookups -> conv -> fused CPG-FC -> 1-N scorer + label-smoothed BCE -> backward -> global-norm
clip -> AMSGrad.  `value` = train rows/s with inputs resident in HBM; `eval` = filtered-rank queries/s
(forward + 1-N scores + filtered rank); `e2e` = the same through the public API from pinned HOST batches
with the loss / ranks read back every step.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            d["_source"] = "measured (MEASURED_PEAKS.json)"
            return d
        except Exception:
            pass
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback (B200_PROFILING.md)"
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path).read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_cfg_of(md, H):
    from oracle import conve_oracle as O
    return O.OracleConfig(num_ent=md["num_ent"], num_rel=md["num_rel"], ent_emb_size=md["ent_emb_size"],
                          rel_emb_size=md["rel_emb_size"], context_rel_out=list(md["context_rel_out"]),
                          conv_in_height=H, hidden_dropout=0.0, output_dropout=0.0,
                          batch_norm_momentum=md["batch_norm_momentum"],
                          batch_norm_train_stats=md["batch_norm_train_stats"])


def run_cpu_baseline(shape, budget_s, note=""):
    """Reference algorithm restated on torch-CPU (oracle/torch_port.py; TensorFlow 1.14 is not installable
    here), all host cores, bounded sample of the same workload."""
    from coper_b200 import synthetic
    from oracle import conve_oracle as O
    from oracle.torch_port import time_cpu_baseline
    s = synthetic.SHAPES[shape]
    md = synthetic.descriptors(shape, dropout=False)
    cfg = oracle_cfg_of(md, s["H"])
    params = O.init_params(cfg, seed=0)
    hb = synthetic.make_batches(s["num_ent"], s["num_rel"], s["batch"], 1, seed=1)[0]
    dense = O.csr_to_dense(hb["e2_multi_rowptr"], hb["e2_multi_col"], s["num_ent"])
    r = time_cpu_baseline(params, cfg, (hb["e1"], hb["rel"], hb["e2"], dense), budget_s=budget_s)
    return {"value": r["train_rows_per_s"], "unit": "train rows/s", "eval_value": r["eval_queries_per_s"],
            "eval_unit": "eval queries/s", "cores": r["cores"], "kind": "port",
            "sample": "%d train steps + %d eval batches of B=%d at the %s shape, torch-CPU fp32 materialising "
                      "restatement of models.py/metrics.py (TensorFlow unavailable offline)%s"
                      % (r["train_steps"], r["eval_batches"], s["batch"], shape, note),
            "train_ms": r["train_ms"], "eval_ms": r["eval_ms"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--shape", default=os.environ.get("COPER_BENCH_SHAPE", "wn18rr"))
    ap.add_argument("--prec", default=os.environ.get("COPER_BENCH_PREC", "tf32x3"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra bf16-engine measurement")
    ap.add_argument("--num-labels", type=int, default=1000,
                    help="also time the sampled-label training step (SURVEY §8f-2) with this many labels; 0 = skip")
    ap.add_argument("--no-graph-multi", action="store_true",
                    help="N > 1: do not capture the step (with its NCCL collectives) in a CUDA graph")
    ap.add_argument("--front-end", default=os.environ.get("COPER_BENCH_FRONT_END", "data-parallel"),
                    choices=["data-parallel", "replicated"],
                    help="N > 1: data-parallel = every rank owns B rows of a global batch of N*B (weak scaling, "
                         "synchronised batch norm, bucketed gradient all-reduce); replicated = all ranks run the "
                         "front end of the same B rows, only the scorer is sharded (strong scaling)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="data-parallel: all-reduce the whole gradient bucket at the end of the backward pass instead "
                         "of overlapping the generator-weight gradient with the conv backward")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from coper_b200 import synthetic
    s = synthetic.SHAPES[args.shape]
    dp = max(world, args.gpus) > 1 and args.front_end == "data-parallel"
    B = s["batch"] * (max(world, args.gpus) if dp else 1)    # data-parallel: the GLOBAL batch, B rows per rank
    workload = "CoPER-ConvE %s shape (N=%d entities, R'=%d, d=%d, dr=%d, B=%d, full 1-N labels)" % (
        args.shape, s["num_ent"], s["num_rel"], s["ent_emb_size"], s["rel_emb_size"], B)

    if args.impl == "reference":
        if rank != 0:
            return
        floor = 10.0 if args.shape != "toy" else 0.5          # (the toy shape only serves the contract test)
        cb = run_cpu_baseline(args.shape, max(args.cpu_budget, floor) * 1.5)
        line = {"impl": "reference", "metric": "train_rows_per_s", "value": cb["value"], "unit": "train rows/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": cb["train_ms"] * B / s["batch"], "higher_is_better": True,
                "scaling": "weak" if (dp or args.gpus == 1) else "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "note": "CPU restatement of the reference path (oracle port)"},
                "eval": {"value": cb["eval_value"], "unit": "eval queries/s", "ms_per_batch": cb["eval_ms"]},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "train rows/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0, "eval_value": cb["eval_value"]}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from coper_b200 import _lib
    from coper_b200.models import ConvE, EntityShard
    torch.cuda.set_device(local_rank)
    # stdout carries exactly ONE JSON line: everything libraries print while we run (e.g. NCCL's version banner)
    # is diverted to stderr at the file-descriptor level
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    md = synthetic.descriptors(args.shape, dropout=True)
    shard = EntityShard(s["num_ent"], rank, world)
    model = ConvE(md, seed=0, prec=args.prec, shard=shard, conv_in_height=s["H"],
                  init_fast=s["num_ent"] > 1_000_000, graphs_multi_gpu=not args.no_graph_multi, data_parallel=dp,
                  overlap_grad_allreduce=not args.no_overlap)
    n_batches = 8
    host = synthetic.make_batches(s["num_ent"], s["num_rel"], B, n_batches, seed=1)
    devb = [{k: torch.as_tensor(v).cuda() for k, v in hb.items()} for hb in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        k0 = lib.coper_launch_count() + model.graph_kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, (lib.coper_launch_count() + model.graph_kernel_launches - k0)

    K, W = args.steps, args.warmup
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident (kernel-side) numbers
    train_ms, train_launches = timed(lambda i: model.train_step(devb[i % n_batches]), K, W)
    eval_ms, eval_launches = timed(lambda i: model.filtered_ranks(devb[i % n_batches]), K, W)
    # ---- end-to-end through the public API: pinned host batches in, loss / ranks read back every step
    def e2e_train(i):
        return float(model.train_step(host[i % n_batches]).item())

    def e2e_eval(i):
        r, _ = model.filtered_ranks(host[i % n_batches])
        return r.cpu()
    e2e_train_ms, _ = timed(e2e_train, K, W)
    e2e_eval_ms, _ = timed(e2e_eval, K, W)
    clocks = sampler.stop() if rank == 0 else None
    nnz = float(np.mean([hb["e2_multi_col"].shape[0] for hb in host]))
    h2d = int(B * 8 * 2 + (B + 1) * 4 + nnz * 4)
    d2h = 8

    # ---- per-kernel breakdown + roofline of the dominant kernel (rank 0, N=1 timing of isolated calls)
    roofline, breakdown = None, None
    peaks = load_peaks()
    if not args.no_breakdown and rank == 0 and world == 1:
        breakdown, roofline = kernel_breakdown(model, devb[0], peaks, args.prec, args.shape)

    # the same workload on the bf16 tensor-pipe engine (north_star's throughput path), reported beside the headline
    alt = None
    if world == 1 and args.prec != "bf16" and not args.no_alt:
        del model
        torch.cuda.empty_cache()
        m16 = ConvE(md, seed=0, prec="bf16", shard=shard, conv_in_height=s["H"], init_fast=s["num_ent"] > 1_000_000)
        model = m16                                   # timed() reads model.graph_kernel_launches
        t_ms, _ = timed(lambda i: m16.train_step(devb[i % n_batches]), K, W)
        e_ms, _ = timed(lambda i: m16.filtered_ranks(devb[i % n_batches]), K, W)
        alt = {"precision": "bf16 (tcgen05 kind::f16, fp32 accumulate)", "value": B / t_ms * 1e3, "unit": "train rows/s",
               "ms_per_step": t_ms, "eval_value": B / e_ms * 1e3, "eval_unit": "eval queries/s", "eval_ms_per_batch": e_ms}
    # the sampled-label training step the shipped big-dataset configs use (num_labels = 1000)
    sampled = None
    if world == 1 and args.num_labels > 0:
        import copy
        md_s = copy.deepcopy(md)
        md_s["use_negative_sampling"] = True
        del model
        torch.cuda.empty_cache()
        ms = ConvE(md_s, seed=0, prec=args.prec, shard=shard, conv_in_height=s["H"])
        model = ms
        sb_host = [synthetic.to_sampled(hb, s["num_ent"], args.num_labels, seed=7 + i) for i, hb in enumerate(host)]
        sb_dev = [{k: torch.as_tensor(v).cuda() for k, v in hb.items()} for hb in sb_host]
        t_ms, n_l = timed(lambda i: ms.train_step(sb_dev[i % n_batches]), K, W)
        e2e_ms, _ = timed(lambda i: float(ms.train_step(sb_host[i % n_batches]).item()), K, W)
        # labels drawn on the device (coper_sample_labels) from host CSR batches: only ids + id lists cross PCIe
        dev_host = [dict(hb, sample_on_device=(args.num_labels, 10.0)) for hb in host]
        ds_ms, _ = timed(lambda i: float(ms.train_step(dev_host[i % n_batches]).item()), K, W)
        sampled = {"num_labels": args.num_labels, "value": B / t_ms * 1e3, "unit": "train rows/s", "ms_per_step": t_ms,
                   "e2e_value": B / e2e_ms * 1e3, "e2e_ms_per_step": e2e_ms,
                   "device_sampling_e2e_value": B / ds_ms * 1e3, "device_sampling_e2e_ms_per_step": ds_ms,
                   "device_sampling_h2d_bytes_per_step": h2d,
                   "h2d_bytes_per_step": int(B * 16 + B * args.num_labels * 8), "gpu_launches_per_step": n_l / K,
                   "gather_bytes_per_step": int(B * args.num_labels * s["ent_emb_size"] * 4)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = run_cpu_baseline(args.shape, args.cpu_budget)
        except Exception as exc:  # the CPU leg must never take the GPU numbers down with it
            cpu = {"value": None, "unit": "train rows/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": "failed: %r" % (exc,)}
    if world > 1:
        dist.barrier()
    if rank != 0:
        _finish(world, torch, dist)
        return
    ws_mb = working_set_mb(s, B)
    line = {
        "metric": "train_rows_per_s", "value": B / train_ms * 1e3, "unit": "train rows/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": train_ms, "higher_is_better": True,
        "scaling": "weak" if (dp or world == 1) else "strong",
        "vs_baseline": None, "dtype": {"fp32": "f32", "bf16": "bf16", "tf32x3": "tf32x3"}[args.prec],
        "data": "synthetic",
        "config": {"workload": workload, "precision": args.prec,
                   "parallelism": "single GPU" if world == 1 else
                   "entity-sharded 1-N scorer x%d (rows/GPU=%d), %s" % (
                       world, shard.rows,
                       "data-parallel front end: global batch %d = %d rows/GPU, sync batch norm, one bucketed "
                       "gradient all-reduce" % (B, B // world) if dp else "replicated front end"),
                   "l2": "no flush: per-step working set ~%d MB > 126 MB L2; %d distinct input batches cycled"
                         % (ws_mb, n_batches),
                   "dropout": "on (feature-map 0.3, output 0.2)", "batch_norm": "batch statistics (train)",
                   "cuda_graph": bool(model.use_graphs and (world == 1 or model.graphs_multi_gpu))},
        "eval": {"value": B / eval_ms * 1e3, "unit": "eval queries/s", "ms_per_batch": eval_ms,
                 "gpu_launches_per_batch": eval_launches / K},
        "e2e": {"value": B / e2e_train_ms * 1e3, "unit": "train rows/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_train_ms,
                "eval_value": B / e2e_eval_ms * 1e3, "eval_unit": "eval queries/s", "eval_ms_per_batch": e2e_eval_ms,
                "eval_d2h_bytes_per_batch": B * 4},
        "gpu_launches": int(train_launches), "gpu_launches_per_step": train_launches / K,
        "clocks": clocks, "roofline": roofline, "kernel_ms": breakdown, "bf16": alt, "sampled_labels": sampled, "cpu_baseline": cpu,
        "peaks": {k: peaks.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained", "_source")},
    }
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    _finish(world, torch, dist)


def _finish(world, torch, dist):
    """N > 1: leave without tearing NCCL down — destroy_process_group() blocks forever once collectives have been
    captured into CUDA graphs (observed on the 2-GPU box); everything is flushed and synchronised first."""
    if world > 1:
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        os._exit(0)


def working_set_mb(s, B):
    N, d, dr = s["num_ent"], s["ent_emb_size"], s["rel_emb_size"]
    H = s["H"]
    F = (H - 2) * (d // H - 2) * 32
    return int((3 * N * d * 4 + 3 * dr * F * d * 4 + B * N * 4 + 4 * B * F * 4) / 1e6)


def kernel_breakdown(model, batch, peaks, prec, shape="wn18rr", reps=10):
    """CUDA-event time of each C-ABI stage in isolation (same stream, warm) and the roofline of the dominant one."""
    import torch
    from coper_b200._lib import call, ptr
    b = model.stage_batch(batch)
    model._train_device(b)                 # populate every buffer
    torch.cuda.synchronize()
    B, d, F, C = b.B, model.ent_emb_size, model.F, model.C
    Ns = model.shard.rows
    g = model.grads
    Pw, Pb = model.fc_weights.projections[-1], model.fc_bias.projections[-1]
    dc = Pw.shape[0]
    pos, neg = 0.9 + 1.0 / model.num_ent, 1.0 / model.num_ent
    bits_q = b.bits                      # query-major filter rows for the standalone (two-pass) rank kernel
    if bits_q is None:
        bits_q = torch.zeros(B, b.words, dtype=torch.int32, device=model.dev)
        call("coper_csr_to_bits", ptr(b.rowptr), ptr(b.col), B, model.shard.lo, model.shard.hi, ptr(bits_q))
    stages = {
        "cpg_fc_fwd": (lambda: call("coper_cpg_fc_fwd", ptr(b.cw), ptr(b.f), ptr(Pw), ptr(model.P_prep), ptr(b.cb), ptr(Pb), B, dc, F, d,
                                    Pb.shape[0], 1.0, None, 0, ptr(b.y), ptr(b.ws_cpg), b.ws_cpg_bytes, model.prec),
                       2.0 * B * dc * F * d, "tensor"),
        "cpg_fc_bwd": (lambda: call("coper_cpg_fc_bwd", ptr(b.cw), ptr(b.f), ptr(Pw), ptr(model.P_prep), ptr(b.cb), ptr(Pb), ptr(b.dy),
                                    B, dc, F, d, Pb.shape[0], ptr(g["fc_weights/CPG/Projection0"]),
                                    ptr(g["fc_bias/CPG/Projection0"]), ptr(b.df), ptr(b.dcw), ptr(b.dcb), ptr(b.ws_cpg),
                                    b.ws_cpg_bytes, model.prec, 0), 4.0 * B * dc * F * d, "tensor"),
        "score1n_bce_fwd_bwd": (lambda: call("coper_score1n_bce_fwd_bwd", ptr(b.q), ptr(model.ent_emb), ptr(model.E_prep),
                                             ptr(model.pred_bias), ptr(b.bits if model.prec == 0 else b.bitsT), B, Ns, d, pos, neg,
                                             1.0 / (B * model.num_ent), ptr(b.loss_sum), ptr(model._grad_buf(b)), b.ld, ptr(b.dq),
                                             ptr(g["ent_emb"]), ptr(g["pred_bias"]), ptr(b.ws), b.ws_bytes,
                                             model.prec), 6.0 * B * d * Ns, "tensor"),
        "score1n_fwd": (lambda: model._score(b), 2.0 * B * d * Ns, "tensor"),
        "score1n_rank_fused": (lambda: model._rank_device(b), 2.0 * B * d * Ns, "tensor"),
        "filtered_rank": (lambda: call("coper_filtered_rank", ptr(model._scores_buf(b)), b.ld, B, Ns, ptr(b.e2), model.shard.lo,
                                       ptr(b.gold), ptr(bits_q), ptr(b.n_greater), ptr(b.n_equal)),
                          B * (4.0 * Ns + Ns / 8.0), "hbm"),
        "clip_and_amsgrad": (lambda: model._clip_and_apply(),
                             sum(p.numel() for _, p, _ in model.trainables) * 4.0 * 5, "hbm"),
    }
    out, best = {}, None
    for name, (fn, work, bound) in stages.items():
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = ms
        in_train = name in ("cpg_fc_fwd", "cpg_fc_bwd", "score1n_bce_fwd_bwd", "clip_and_amsgrad")
        if in_train and (best is None or ms > best[1]):
            best = (name, ms, work, bound)
    name, ms, work, bound = best
    if bound == "tensor":
        peak = peaks["bf16_tflops"]
        achieved = work / (ms * 1e-3) / 1e12
        unit = "TFLOP/s"
    else:
        peak = peaks["hbm_gbs"]
        achieved = work / (ms * 1e-3) / 1e9
        unit = "GB/s"
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        traffic = tj.get(shape, {}).get(prec, {}).get(name)
    except Exception:
        pass
    roof = {"kernel": name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
            "traffic": traffic, "ms": ms, "algorithmic_work": work,
            "peak_source": peaks["_source"] + (", bf16 dense burst" if bound == "tensor" else ", copy bandwidth"),
            "note": "isolated launches of the C-ABI stage (its tcgen05 kernels + reductions), CUDA events on the launching "
                    "stream; prec=%s%s; traffic = DRAM bytes of the stage from the committed ncu capture" % (
                        prec, " (3 tf32 MMAs per product: tensor-pipe time = 6x the bf16-equivalent of the algorithmic "
                              "FLOPs)" if prec == "tf32x3" else "")}
    return out, roof


if __name__ == "__main__":
    main()
