#!/usr/bin/env python
"""Root-cause probe of the experimental "packed" sampling mode (gradients from one RG32F fetch): run a configuration with
packed = 3 (sample both ways, use the reference's four fetches, record every lane that passed the exactness conditions and
still differs), then time packed = 0 / 2.  Usage: python tools/packed_probe.py [--config 2]"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gipuma_b200 import scene as S, api

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--rows", type=int, default=None)
ap.add_argument("--cols", type=int, default=None)
args = ap.parse_args()
sc = S.make_config(args.config, rows=args.rows, cols=args.cols, workers=16)
out = {"config": sc.name, "rows": sc.rows, "cols": sc.cols, "views": sc.n_views}
res = {}
for mode in (0, 3, 2):
    with api.Context(sc.cols, sc.rows, sc.n_views) as ctx:
        ctx.set_option("packed", mode)
        ctx.load_scene(sc)
        ctx.packed_mismatches(reset=True)
        ms = ctx.run()
        ms = ctx.run()
        n4, c = ctx.get_state()
        n, rec = ctx.packed_mismatches(reset=True)
        res[mode] = (n4, c)
        out["packed=%d" % mode] = {"sweep_ms": ms, "mismatching_fetches": n, "pairs": ctx.stats()["pairs"]}
        if mode == 3:
            out["records"] = [[float(x) for x in r] for r in rec[:24]]
for mode in (3, 2):
    out["packed=%d" % mode]["pixels_differing_from_packed0"] = int((res[mode][0].view(np.uint32) != res[0][0].view(np.uint32)).any(axis=-1).sum())
print(json.dumps(out))
