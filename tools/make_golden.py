#!/usr/bin/env python
"""Generate tests/golden/*.npz with the PINNED REFERENCE BUILD (oracle/_ref) on a B200.

Each fixture holds a complete small scene (8-bit images, Camera_cu field values, parameters, view subset) and
the reference's outputs on it: state after gipuma_init_cu2, after the three black kernels of iteration 1, after
the whole iteration 1, and the final runcuda() output.  The GPU tests compare gipuma_b200 with these bit for
bit, so parity stays pinned on boxes where /root/reference and oracle/_ref do not exist.
Run on the GPU box:  python tools/make_golden.py gpurun_out/golden   (then copy the files to tests/golden/)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gipuma_b200 import scene as S          # noqa: E402
from gipuma_b200.golden import scene_to_arrays   # noqa: E402
from oracle import pyref                    # noqa: E402

CASES = {
    # name: (config, rows, cols, views, iterations, box, n_best)
    "gA_box15_v2": (1, 64, 96, 2, 2, 15, 2),
    "gB_box11_v5": (2, 64, 96, 5, 2, 11, 3),
    "gC_box11_v34": (4, 64, 96, 34, 1, 11, 3),
    "gD_box25_v3": (3, 64, 96, 3, 1, 25, 2),
    "gE_color_box11_v4": (2, 64, 96, 4, 2, 11, 3),      # -color_processing: float4 images (S.colorize)
}


FUSED_CASES = {
    # the fused 20-neighbour sweep (reference built without SMALLKERNEL): name: (config, rows, cols, views, iterations, box, n_best, colour)
    "fused_box11_v4": (2, 64, 96, 4, 2, 11, 3, False),
    "fused_color_box9_v3": (2, 64, 96, 3, 2, 9, 2, True),
}


def fused(out_dir, only):
    for name, (cfg, rows, cols, views, iters, box, nbest, colour) in FUSED_CASES.items():
        if only and name not in only:
            continue
        sc = S.make_config(cfg, rows=rows, cols=cols, n_views=views, iterations=iters)
        sc.params.box_hsize = sc.params.box_vsize = box
        sc.params.n_best = nbest
        if colour:
            sc = S.colorize(sc)
        h = pyref.Harness("ref")
        seed = 0xC0FFEE
        i_n4, i_c, _ = h.steps(sc, [pyref.STEP_INIT], seed=seed)
        b_n4, b_c, _ = h.steps(sc, [pyref.STEP_BLACK_FUSED], norm4=i_n4, cost=i_c, seed=seed)
        t_n4, t_c, _ = h.steps(sc, [pyref.STEP_RED_FUSED], norm4=b_n4, cost=b_c, seed=seed)
        f_n4, f_c, _ = h.run_fused(sc, seed=seed)
        arrs = scene_to_arrays(sc)
        arrs.update(seed=np.uint64(seed), init_norm4=i_n4, init_cost=i_c, black_norm4=b_n4, black_cost=b_c,
                    iter1_norm4=t_n4, iter1_cost=t_c, final_norm4=f_n4, final_cost=f_c)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrs)
        print(name, os.path.getsize(path), "bytes")


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[2:]
    fused(out_dir, only)
    for name, (cfg, rows, cols, views, iters, box, nbest) in CASES.items():
        if only and name not in only:
            continue
        sc = S.make_config(cfg, rows=rows, cols=cols, n_views=views, iterations=iters)
        sc.params.box_hsize = sc.params.box_vsize = box
        sc.params.n_best = nbest
        if "color" in name:
            sc = S.colorize(sc)
        h = pyref.Harness("ref64" if sc.n_views > 32 else "ref")
        seed = 0xC0FFEE
        i_n4, i_c, _ = h.steps(sc, [0], seed=seed)
        b_n4, b_c, _ = h.steps(sc, [1, 2, 3], norm4=i_n4, cost=i_c, seed=seed)
        t_n4, t_c, _ = h.steps(sc, [4, 5, 6], norm4=b_n4, cost=b_c, seed=seed)
        sweep_cost = h.cost_eval(sc, i_n4)
        f_n4, f_c, _, _ = h.run(sc, seed=seed)
        arrs = scene_to_arrays(sc)
        arrs.update(seed=np.uint64(seed), init_norm4=i_n4, init_cost=i_c, init_planes_sweep_cost=sweep_cost,
                    black_norm4=b_n4, black_cost=b_c, iter1_norm4=t_n4, iter1_cost=t_c,
                    final_norm4=f_n4, final_cost=f_c)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrs)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
