"""Time the collectives of the data-parallel step in isolation (torchrun --nproc-per-node P tools/dp_collectives.py).
Each collective is captured in a CUDA graph of 20 back-to-back calls and replayed; device time, max over ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    B, d, F, dc = 512, 200, 4608, 8
    f32 = dict(dtype=torch.float32, device="cuda")
    big = torch.zeros(dc * F * d + 4096, **f32)
    xg, xl = torch.zeros(B * world, d, **f32), torch.zeros(B, d, **f32)
    st, st_all = torch.zeros(36 * 32 * 2, **f32), torch.zeros(world * 36 * 32 * 2, **f32)
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    cases = {
        "all_reduce flat bucket (%.1f MB)" % (big.numel() * 4 / 1e6): lambda: dist.all_reduce(big),
        "reduce_scatter [Bg,d] -> [Bl,d]": lambda: dist.reduce_scatter_tensor(xl, xg),
        "all_gather [Bl,d] -> [Bg,d]": lambda: dist.all_gather_into_tensor(xg, xl),
        "all_gather BN partials (9 KB)": lambda: dist.all_gather_into_tensor(st_all, st),
        "all_reduce loss (8 B)": lambda: dist.all_reduce(loss),
    }
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for _ in range(20):
                fn()
        g.replay()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 100], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("%-44s %8.1f us" % (name, t.item() * 1e3), flush=True)
    sys.stdout.flush()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
