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
ap.add_argument("--hybrid", type=int, default=0, help="view-shard width; world/width groups take different reference views")
ap.add_argument("--refs", type=int, default=2)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if args.hybrid:
    mk = lambda ref: S.make_config(args.config, rows=args.rows, cols=args.cols, n_views=args.views, iterations=args.iters, seed=1234 + 17 * ref)
    outs = {}
    times = M.run_hybrid(mk, args.refs, rank, world, args.hybrid, device=local, on_result=lambda ref, n4, c: outs.__setitem__(ref, (n4, c)))
    ok = True
    for ref, (n4, c) in outs.items():
        single, _, _ = api.runcuda(mk(ref), device=local)
        ok = ok and np.array_equal(n4.view(np.uint32), single.norm4.view(np.uint32)) and np.array_equal(c.view(np.uint32), single.c.view(np.uint32))
    t = torch.tensor([sum(times), float(ok)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX); tn = t.clone(); dist.all_reduce(tn, op=dist.ReduceOp.MIN)
        t = torch.stack([tm[0], tn[1]])
    if rank == 0:
        print(json.dumps({"mode": "hybrid", "world": world, "shard": args.hybrid, "reference_views": args.refs, "wall_s": float(t[0]),
                          "all_groups_bit_identical_to_single_gpu": bool(t[1] > 0.5)}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0)
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
