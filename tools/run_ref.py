#!/usr/bin/env python
"""Run the pinned reference build (oracle/_ref) on one BASELINE.json configuration and print its timing.

Bench/test infrastructure (uses oracle/).  Usage on the GPU box:
    python tools/run_ref.py --config 2 [--iters 8] [--rows R --cols C --views V] [--save out.npz] [--steps]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gipuma_b200 import scene as S          # noqa: E402
from oracle import pyref                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--iters", type=int, default=None)
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--cols", type=int, default=None)
    ap.add_argument("--views", type=int, default=None)
    ap.add_argument("--color", action="store_true", help="float4 images (-color_processing)")
    ap.add_argument("--fused", action="store_true", help="the fused 20-neighbour kernels (reference built without SMALLKERNEL)")
    ap.add_argument("--save", type=str, default=None)
    ap.add_argument("--steps", action="store_true", help="also time each kernel of one iteration (hx_steps)")
    ap.add_argument("--repeat", type=int, default=1)
    args = ap.parse_args()

    t0 = time.time()
    sc = S.make_config(args.config, rows=args.rows, cols=args.cols, n_views=args.views, iterations=args.iters)
    if args.color:
        sc = S.colorize(sc)
    t_scene = time.time() - t0
    h = pyref.Harness("ref64" if sc.n_views > 32 else "ref")
    out = {"config": sc.name, "rows": sc.rows, "cols": sc.cols, "views": sc.n_views, "iters": sc.params.iterations,
           "box": sc.params.box_hsize, "scene_s": round(t_scene, 2), "backend": h.backend, "runs": []}
    n4 = c = None
    for r in range(args.repeat):
        if args.fused:
            n4, c, ms = h.run_fused(sc)
            printed_s, wall_ms = ms / 1000.0, ms
        else:
            n4, c, printed_s, wall_ms = h.run(sc)
        mpix = sc.rows * sc.cols * sc.params.iterations / 1e6 / printed_s if printed_s > 0 else float("nan")
        out["runs"].append({"printed_s": printed_s, "wall_ms": wall_ms, "mpixel_iters_per_s": mpix})
    if sc.gt_depth is not None:
        d = n4[..., 3]
        ok = d > 0
        rel = np.abs(d - sc.gt_depth) / sc.gt_depth
        out["frac_within_1pct_of_gt"] = float((rel[ok] < 0.01).mean()) if ok.any() else 0.0
        out["mean_cost"] = float(c.mean())
    if args.steps:
        n4s, cs, ms = h.steps(sc, [0, 1, 2, 3, 4, 5, 6, 7])
        out["step_ms"] = [float(x) for x in ms]
    print(json.dumps(out))
    if args.save:
        np.savez_compressed(args.save, norm4=n4, cost=c)


if __name__ == "__main__":
    main()
