"""Stage-by-stage comparison of the float4 (-color_processing) path against the live pinned reference build.
Run on a B200:  python tools/color_probe.py [rows cols views box]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gipuma_b200 import api, scene as S          # noqa: E402
from oracle import pyref                         # noqa: E402


def nbits(a, b):
    return int((np.ascontiguousarray(a).view(np.uint32) != np.ascontiguousarray(b).view(np.uint32)).sum())


def main():
    rows, cols, views, box = (int(v) for v in (sys.argv[1:5] + ["96", "128", "5", "11"][len(sys.argv) - 1:]))
    sc = S.colorize(S.make_config(2, rows=rows, cols=cols, n_views=views, iterations=2, seed=4242))
    sc.params.box_hsize = sc.params.box_vsize = box
    ref = pyref.Harness("ref")
    seed = 31337
    n4, c, _ = ref.steps(sc, [pyref.STEP_INIT], seed=seed)
    rc = ref.cost_eval(sc, n4)
    print("ref init cost vs ref cost_eval kernel: %d differing" % nbits(c, rc))
    with api.Context(sc.cols, sc.rows, sc.n_views) as ctx:
        ctx.load_scene(sc, seed=seed)
        ctx.init()
        m4, mc = ctx.get_state()
        print("init planes diff %d  cost diff %d  (max abs %.3g)" % (nbits(m4, n4), nbits(mc, c), np.abs(mc - c).max()))
        for variant in (0, 1, 2, 3):
            ctx.set_option("cost_variant", variant)
            e = ctx.cost_eval(n4)
            print("cost_eval variant %d: vs ref init cost %d, vs ref cost_eval %d" % (variant, nbits(e, c), nbits(e, rc)))
        ctx.set_option("cost_variant", -1)
        ctx.set_state(n4, c)
        for step, (colour, mask) in zip(range(1, 7), [(0, 1), (0, 2), (0, 4), (1, 1), (1, 2), (1, 4)]):
            n4, c, _ = ref.steps(sc, [step], norm4=n4, cost=c, seed=seed)
            ctx.phase(colour, mask)
            m4, mc = ctx.get_state()
            print("step %d: planes diff %d  cost diff %d  (max abs cost %.3g)" % (step, nbits(m4, n4), nbits(mc, c),
                                                                                np.abs(mc - c).max()))
            ctx.set_state(n4, c)                      # keep the two in lock step
    r4, rcst, _, _ = ref.run(sc, seed=seed)
    for opts in ({}, {"memo": 0, "prune": 0, "dedupe": 0}):
        ls, ms, _ = api.runcuda(sc, seed=seed, options=opts)
        print("full run %s: planes diff %d  cost diff %d  (%.2f ms)" % (opts, nbits(ls.norm4, r4), nbits(ls.c, rcst), ms))
    drop = pyref.Harness("dropin")
    d4, dc, _, _ = drop.run(sc, seed=seed)
    print("drop-in harness: planes diff %d  cost diff %d" % (nbits(d4, r4), nbits(dc, rcst)))


if __name__ == "__main__":
    main()
