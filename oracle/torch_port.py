"""CPU ORACLE / CPU BASELINE (test infrastructure, NOT product code).

Op-for-op torch-CPU port of the reference's *materialising* formulation of the hot path,
used (a) to cross-check ``conve_oracle``'s analytic backward via autograd and (b) as the
timed CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``; kind = "port",
because TensorFlow 1.14 cannot be installed here — see DESIGN.md).

Follows /root/reference/CoPER_ConvE/qa_cpg/:
  models.py:176-180 (lookups), :56-76 (CPG generate: ONE GEMM that writes [B,F,d] weights),
  :354-426 (conv -> BN -> relu -> dropout -> flatten -> batched mat-vec -> dropout -> BN -> relu),
  :433-437 (transpose + dense 1-N GEMM + bias), :448-457 (smoothed sigmoid-BCE, mean),
  :196-200 (autodiff, clip_by_global_norm 5.0, AMSGrad apply), metrics.py:44-57 (argsort loop).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this file.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from .conve_oracle import BN_EPS, OracleConfig, rank_literal, summarize_ranks


def _t(x, dtype, requires_grad=False):
    t = torch.tensor(np.asarray(x), dtype=dtype)
    t.requires_grad_(requires_grad)
    return t


class TorchPort:
    """Holds the variables as torch CPU tensors; ``loss_and_grads`` / ``train_step`` / ``eval_batch``."""

    def __init__(self, params, cfg: OracleConfig, dtype=torch.float32, lr=1e-3, reference_bug_compat=True):
        self.cfg, self.dtype = cfg, dtype
        self.v = {}
        for k in ("ent_emb", "rel_emb", "conv1_weights", "conv1_bias", "pred_bias"):
            self.v[k] = _t(params[k], dtype, True)
        for name in ("fc_weights_proj", "fc_bias_proj"):
            for i, a in enumerate(params[name]):
                self.v[f"{name}.{i}"] = _t(a, dtype, True)
        self.bn = {}
        for name in ("Conv1BN", "FCBN"):
            self._add_bn(name, params[name])
        for name in ("fc_weights_bn", "fc_bias_bn"):
            for i, b in enumerate(params[name]):
                self._add_bn(f"{name}.{i}", b)
        self.n_w = len(params["fc_weights_proj"])
        self.n_b = len(params["fc_bias_proj"])
        self.lr, self.b1, self.b2, self.eps = lr, 0.9, 0.999, 1e-8
        self.b1p, self.b2p = self.b1, self.b2
        self.compat = reference_bug_compat
        self.slots = {}

    def _add_bn(self, name, b):
        self.v[name + ".gamma"] = _t(b["gamma"], self.dtype, True)
        self.v[name + ".beta"] = _t(b["beta"], self.dtype, True)
        self.bn[name] = {"mm": _t(b["moving_mean"], self.dtype), "mv": _t(b["moving_var"], self.dtype)}

    # -- tf.layers.batch_normalization ------------------------------------------------------------
    def _bn(self, x, name, use_batch, bessel, update):
        red = tuple(range(x.dim() - 1))
        st = self.bn[name]
        if use_batch:
            mean = x.mean(dim=red)
            var = ((x - mean) ** 2).mean(dim=red)
            if update:
                n = x.numel() // x.shape[-1]
                mom = self.cfg.batch_norm_momentum
                with torch.no_grad():
                    vm = var * (n / max(n - 1, 1)) if bessel else var
                    st["mm"] = st["mm"] * mom + mean * (1 - mom)
                    st["mv"] = st["mv"] * mom + vm * (1 - mom)
        else:
            mean, var = st["mm"], st["mv"]
        return (x - mean) * torch.rsqrt(var + BN_EPS) * self.v[name + ".gamma"] + self.v[name + ".beta"]

    def _generate(self, r, which, n_layers, is_train, masks, update):
        cfg = self.cfg
        h = r
        keep = 1.0 - (cfg.context_rel_dropout if is_train else 0.0)
        for i in range(n_layers - 1):                                      # models.py:59-68
            h = h @ self.v[f"{which}_proj.{i}"]
            if cfg.context_rel_use_batch_norm:
                h = self._bn(h, f"{which}_bn.{i}", bool(cfg.batch_norm_train_stats and is_train), False, update)
            h = torch.relu(h)
            if keep < 1.0 and masks is not None:
                h = h * _t(masks[i], self.dtype) / keep
        return h @ self.v[f"{which}_proj.{n_layers - 1}"]                  # models.py:70 — materialises [B, F*d]

    def predict(self, e1, rel, is_train=False, masks=None, update_moving=False):
        cfg, v = self.cfg, self.v
        masks = masks or {}
        e1 = torch.as_tensor(np.asarray(e1), dtype=torch.int64)
        rel = torch.as_tensor(np.asarray(rel), dtype=torch.int64)
        B = e1.shape[0]
        H, W, C = cfg.conv_in_height, cfg.conv_in_width, cfg.conv_num_channels
        F, d = cfg.fc_input_size, cfg.ent_emb_size
        x0 = v["ent_emb"][e1]
        r = v["rel_emb"][rel]
        img = x0.reshape(B, 1, H, W)
        w = v["conv1_weights"].permute(3, 2, 0, 1)                          # HWIO -> OIHW
        z = torch.nn.functional.conv2d(img, w) + v["conv1_bias"].view(1, C, 1, 1)
        z = z.permute(0, 2, 3, 1)                                           # NHWC like the reference
        use_batch = bool(cfg.batch_norm_train_stats and is_train)
        z = self._bn(z, "Conv1BN", use_batch, True, update_moving)
        a = torch.relu(z)
        keep1 = 1.0 - (cfg.hidden_dropout if is_train else 0.0)
        if keep1 < 1.0 and "feature_map" in masks:
            a = a * _t(masks["feature_map"], self.dtype) / keep1
        f = a.reshape(B, F)                                                 # (h,w,c) order, models.py:404
        Wgen = self._generate(r, "fc_weights", self.n_w, is_train, masks.get("ctx_w"), update_moving).reshape(B, F, d)
        bgen = self._generate(r, "fc_bias", self.n_b, is_train, masks.get("ctx_b"), update_moving)
        y = torch.bmm(f[:, None, :], Wgen)[:, 0, :] + bgen                  # models.py:412
        keep2 = 1.0 - (cfg.output_dropout if is_train else 0.0)
        if keep2 < 1.0 and "output" in masks:
            y = y * _t(masks["output"], self.dtype) / keep2
        y = self._bn(y, "FCBN", use_batch, False, update_moving)
        q = torch.relu(y)
        S = q @ v["ent_emb"].t() + v["pred_bias"]                           # models.py:434-437
        return S, q

    def loss(self, S, labels):
        cfg = self.cfg
        z = torch.as_tensor(np.asarray(labels), dtype=self.dtype)
        zs = (1 - cfg.label_smoothing_epsilon) * z + (1.0 / cfg.num_ent)    # models.py:450
        return torch.nn.functional.binary_cross_entropy_with_logits(S, zs, reduction="mean")

    def loss_and_grads(self, e1, rel, labels, is_train=True, masks=None, update_moving=False):
        for t in self.v.values():
            t.grad = None
        S, _ = self.predict(e1, rel, is_train, masks, update_moving)
        L = self.loss(S, labels)
        L.backward()
        return float(L.detach()), {k: (None if t.grad is None else t.grad.detach().numpy().copy()) for k, t in self.v.items()}

    def train_step(self, e1, rel, labels, masks=None):
        """One full reference step: fwd, autograd bwd, clip 5.0, AMSGrad (amsgrad.py dense rule)."""
        for t in self.v.values():
            t.grad = None
        S, _ = self.predict(e1, rel, True, masks, True)
        L = self.loss(S, labels)
        L.backward()
        with torch.no_grad():
            gs = [t.grad for t in self.v.values() if t.grad is not None]
            norm = torch.sqrt(sum((g.double() ** 2).sum() for g in gs)).item()
            scale = 5.0 / max(norm, 5.0)
            lr_t = self.lr * math.sqrt(1 - self.b2p) / (1 - self.b1p)
            for k, t in self.v.items():
                if t.grad is None:
                    continue
                g = t.grad * scale
                st = self.slots.setdefault(k, {"m": torch.zeros_like(t), "v": torch.zeros_like(t),
                                               "vhat": torch.zeros_like(t)})
                if self.compat:
                    m_t = st["m"] * self.b1 + g * (1 - self.b1)
                    v_t = st["v"] * self.b2 + g * g * (1 - self.b2)
                    st["m"] *= self.b1
                    st["v"] *= self.b2
                else:
                    st["m"] = m_t = st["m"] * self.b1 + g * (1 - self.b1)
                    st["v"] = v_t = st["v"] * self.b2 + g * g * (1 - self.b2)
                st["vhat"] = torch.maximum(st["vhat"], v_t)
                t -= lr_t * m_t / (st["vhat"].sqrt() + self.eps)
            self.b1p *= self.b1
            self.b2p *= self.b2
        return float(L)

    def eval_batch(self, e1, rel, e2, e2_multi):
        """Forward + the literal metrics.py:44-57 host loop; returns ranks."""
        with torch.no_grad():
            S, _ = self.predict(e1, rel, False)
        return rank_literal(S.numpy(), e2, e2_multi)


def time_cpu_baseline(params, cfg, batch, budget_s=20.0, threads=None, min_steps=2):
    """Timed train rows/s and eval queries/s of the port on this host's cores (bounded sample)."""
    import os
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    e1, rel, e2, labels = batch
    port = TorchPort(params, cfg, torch.float32)
    port.train_step(e1, rel, labels)                       # warm-up
    n, t0 = 0, time.perf_counter()
    while n < min_steps or (time.perf_counter() - t0) < budget_s / 2:
        port.train_step(e1, rel, labels)
        n += 1
    train_s = (time.perf_counter() - t0) / n
    port.eval_batch(e1, rel, e2, labels)
    m, t0 = 0, time.perf_counter()
    while m < min_steps or (time.perf_counter() - t0) < budget_s / 2:
        ranks = port.eval_batch(e1, rel, e2, labels)
        m += 1
    eval_s = (time.perf_counter() - t0) / m
    summarize_ranks(ranks)
    B = len(e1)
    return {"train_rows_per_s": B / train_s, "eval_queries_per_s": B / eval_s, "cores": threads,
            "train_steps": n, "eval_batches": m, "train_ms": train_s * 1e3, "eval_ms": eval_s * 1e3}
