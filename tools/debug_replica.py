import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oracle import conve_oracle as O
from coper_b200.models import ConvE
from coper_b200.sharding import EntityShard
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_multi import _descr
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cfg = O.OracleConfig(num_ent=1003, num_rel=22, ent_emb_size=200, rel_emb_size=8, context_rel_out=[],
                     batch_norm_train_stats=True, batch_norm_momentum=0.1, hidden_dropout=0.3, output_dropout=0.2)
params = O.init_params(cfg, seed=3, bias_noise=0.05)
B = 130
e1, rel, e2, rowptr, col = O.synthetic_batch(cfg, B, seed=5, mean_pos=5.0)
batch = {"e1": e1, "rel": rel, "e2": e2, "e2_multi_rowptr": rowptr, "e2_multi_col": col}
m = ConvE(_descr(cfg), device="cuda:%d" % rank, seed=0, shard=EntityShard(cfg.num_ent, rank, world))
m.load_variables(params)
m.train_step(batch)
torch.cuda.synchronize()
b = m._bufs[B]
def h(t): return hashlib.md5(t.detach().cpu().numpy().tobytes()).hexdigest()[:8]
items = {"x0": b.x0, "r": b.r, "z": b.z, "f": b.f, "y": b.y, "q": b.q, "dq": b.dq, "dy": b.dy, "df": b.df, "dz": b.dz,
         "dx0": b.dx0, "dr": b.dr, "loss": b.loss_sum, "sumsq": m.sumsq, "clip": m.clip_out, "state": m.step_state,
         "gP": m.grads["fc_weights/CPG/Projection0"], "grel": m.grads["rel_emb"], "gconv": m.grads["conv1_weights"],
         "P": m.fc_weights.projections[0], "vhatP": m.vhat["fc_weights/CPG/Projection0"], "a1": m.conv1_bn.a, "a2": m.fc_bn.a}
print("rank", rank, " ".join("%s=%s" % (k, h(v)) for k, v in items.items()), flush=True)
dist.destroy_process_group()
