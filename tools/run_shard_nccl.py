#!/usr/bin/env python
"""Source-view shard over real GPUs (NCCL): torchrun --nproc-per-node N tools/run_shard_nccl.py [--config 4]
Rank 0 also runs the single-GPU job and checks that the sharded result is bit-identical; prints one JSON line."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from gipuma_b200 import scene as S, api, multigpu as M

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--rows", type=int, default=None)
ap.add_argument("--cols", type=int, default=None)
ap.add_argument("--views", type=int, default=None)
ap.add_argument("--iters", type=int, default=None)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = S.make_config(args.config, rows=args.rows, cols=args.cols, n_views=args.views, iterations=args.iters)
run = M.ViewShardRunner(sc, rank, world, device=local)
run.run()                                              # warm-up (NCCL communicator, kernels)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
n4, c = run.run()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
out = None
if rank == 0:
    single, ms, _ = api.runcuda(sc, device=local)
    same = np.array_equal(n4.view(np.uint32), single.norm4.view(np.uint32)) and np.array_equal(c.view(np.uint32), single.c.view(np.uint32))
    out = {"mode": "view_shard", "config": sc.name, "rows": sc.rows, "cols": sc.cols, "views": sc.n_views, "world": world,
           "wall_s_incl_init": float(t), "mpixel_iters_per_s": sc.rows * sc.cols * sc.params.iterations / 1e6 / float(t),
           "single_gpu_sweep_ms": ms, "bit_identical_to_single_gpu": bool(same), "collectives": run.collectives}
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
if out:
    print(json.dumps(out))
