#!/usr/bin/env python
"""Run gipuma_b200 on one BASELINE.json configuration and print its timing (development tool)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gipuma_b200 import scene as S, api

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--iters", type=int, default=None)
ap.add_argument("--rows", type=int, default=None)
ap.add_argument("--cols", type=int, default=None)
ap.add_argument("--views", type=int, default=None)
ap.add_argument("--color", action="store_true", help="float4 images (-color_processing)")
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--opt", nargs="*", default=[])
args = ap.parse_args()
sc = S.make_config(args.config, rows=args.rows, cols=args.cols, n_views=args.views, iterations=args.iters)
if args.color:
    sc = S.colorize(sc)
opts = {k: int(v) for k, v in (o.split("=") for o in args.opt)}
out = {"config": sc.name, "rows": sc.rows, "cols": sc.cols, "views": sc.n_views, "iters": sc.params.iterations, "opts": opts, "runs": []}
with api.Context(sc.cols, sc.rows, sc.n_views) as ctx:
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.load_scene(sc)
    for r in range(args.repeat):
        ms = ctx.run()
        st = ctx.stats()
        out["runs"].append({"sweep_ms": ms, "mpixel_iters_per_s": sc.rows * sc.cols * sc.params.iterations / 1e3 / ms, "stats": st})
    n4, c = ctx.get_state()
d = n4[..., 3]
ok = d > 0
out["frac_within_1pct_of_gt"] = float((np.abs(d - sc.gt_depth)[ok] / sc.gt_depth[ok] < 0.01).mean())
out["mean_cost"] = float(c.mean())
print(json.dumps(out))
