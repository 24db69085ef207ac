"""Which rounding variant does each of the 21 inlined call sites (20 candidates + refinement) of the reference's fused
kernels gipuma_black_cu / gipuma_red_cu use?  Run on a B200 with oracle/_ref present:  python tools/fused_probe.py
Prints the site tables (the GPM_FUSED_* masks of gpm_device.cuh) and then checks whole fused iterations with them.

Method.  Propagation sites: start every pixel at cost 1000 with refinement switched off (max_disparity so small that the
refinement loop of gipuma.cu:958 has no step); the fused kernel then leaves, per pixel, the cheapest in-range neighbour
plane and ITS cost as evaluated at the winning call site.  gpm_cost_eval gives the cost of every neighbour plane under
each rounding variant; the variant of site k is the one that reproduces all pixels won by site k.  Refinement site: make
all neighbour planes fall outside the depth range, so only the refinement acts, and try each variant."""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gipuma_b200 import api, scene as S          # noqa: E402
from oracle import pyref                         # noqa: E402

DX = [0, 0, 0, 0, 0, 0, -1, -3, -5, 1, 3, 5, 2, 2, -2, -2, -1, 1, -1, 1]
DY = [-1, -3, -5, 1, 3, 5, 0, 0, 0, 0, 0, 0, -1, 1, -1, 1, -2, -2, 2, 2]
KERNELS = (("black", pyref.STEP_BLACK_FUSED, 0), ("red", pyref.STEP_RED_FUSED, 1))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def shifted(n4, dx, dy):
    """planes of the neighbour at (x+dx, y+dy) for every pixel (wrapped at the edges; masked by the caller)."""
    return np.roll(np.roll(n4, -dy, axis=0), -dx, axis=1)


def probe(color):
    tag = "float4" if color else "float"
    sc = S.make_config(2, rows=256, cols=256, n_views=3, iterations=1, seed=77)
    sc.params.box_hsize = sc.params.box_vsize = 9
    if color:
        sc = S.colorize(sc)
    ref = pyref.Harness("ref")
    seed = 4711
    n4, c, _ = ref.steps(sc, [pyref.STEP_INIT], seed=seed)
    H, W = sc.rows, sc.cols
    yy, xx = np.mgrid[0:H, 0:W]
    variants = tuple(range(16)) if color else (0, 1, 4, 5, 8, 9, 12, 13)
    no_refine = dataclasses.replace(sc, params=dataclasses.replace(sc.params, max_disparity=0.015, min_disparity=0.0))
    masks = {}
    with api.Context(W, H, sc.n_views) as ctx:
        ctx.load_scene(sc, seed=seed)
        costs = {}
        for k in range(20):
            pk = np.ascontiguousarray(shifted(n4, DX[k], DY[k]))
            for v in variants:
                ctx.set_option("cost_variant", v)
                costs[(k, v)] = ctx.cost_eval(pk)
        ctx.set_option("cost_variant", -1)
        for name, step, colour in KERNELS:
            start_c = np.full((H, W), 1000.0, np.float32)
            r4, rc, _ = ref.steps(no_refine, [step], norm4=n4, cost=start_c, seed=seed)
            changed = (bits(r4) != bits(n4)).any(axis=-1)
            print("%s %s: %d pixels changed, (x+y)&1 of changed pixels: %s" % (tag, name, changed.sum(),
                                                                               np.unique((xx + yy)[changed] & 1)))
            sites = []
            for k in range(20):
                pk = shifted(n4, DX[k], DY[k])
                inside = (xx + DX[k] >= 0) & (xx + DX[k] < W) & (yy + DY[k] >= 0) & (yy + DY[k] < H)
                won = changed & inside & (bits(r4) == bits(pk)).all(axis=-1)
                n = int(won.sum())
                agree = {v: int((bits(costs[(k, v)])[won] == bits(rc)[won]).sum()) for v in variants}
                ok = [v for v in variants if agree[v] == n]
                print("  site %2d (%+d,%+d): won %5d  disagree %s  -> %s" % (
                    k, DX[k], DY[k], n, {v: n - a for v, a in agree.items() if v < 4 or a == n}, ok))
                sites.append(ok[0] if ok else 0)
            masks[name] = sites
        # refinement site
        far = np.array([0.0, 0.0, -1.0, 10.0 * sc.params.depthMax], np.float32)
        ctx.set_option("neighbours", 20)
        for name, step, colour in KERNELS:
            m4 = n4.copy()
            m4[((xx + yy) & 1) != colour_parity(name, ref, sc, n4, c, seed, xx, yy)] = far
            r4, rc, _ = ref.steps(sc, [step], norm4=m4, cost=c, seed=seed)
            for k in range(20):
                ctx.set_option("site%d" % k, masks[name][k])
            for v in variants:
                ctx.set_option("site20", v)
                ctx.set_state(m4, c)
                ctx.phase(colour, 7)
                o4, oc = ctx.get_state()
                d4, dc = int((bits(o4) != bits(r4)).sum()), int((bits(oc) != bits(rc)).sum())
                print("  %s %s refinement site, variant %d: planes diff %d cost diff %d" % (tag, name, v, d4, dc))
                if d4 == 0 and dc == 0 and len(masks[name]) == 20:
                    masks[name].append(v)
            if len(masks[name]) == 20:
                masks[name].append(0)
    for name in masks:
        print("%s %s: sites = {%s}" % (tag, name, ", ".join(str(v) for v in masks[name])))
    return masks


def colour_parity(name, ref, sc, n4, c, seed, xx, yy):
    """(x+y)&1 of the pixels the kernel updates."""
    step = dict((k[0], k[1]) for k in KERNELS)[name]
    r4, _, _ = ref.steps(sc, [step], norm4=n4, cost=c, seed=seed)
    ch = (bits(r4) != bits(n4)).any(axis=-1)
    return int(np.unique((xx + yy)[ch] & 1)[0])


def verify(color, masks):
    tag = "float4" if color else "float"
    sc = S.make_config(2, rows=128, cols=160, n_views=5, iterations=2, seed=99)
    sc.params.box_hsize = sc.params.box_vsize = 11
    if color:
        sc = S.colorize(sc)
    ref = pyref.Harness("ref")
    seed = 99
    n4, c, _ = ref.steps(sc, [pyref.STEP_INIT], seed=seed)
    with api.Context(sc.cols, sc.rows, sc.n_views) as ctx:
        ctx.set_option("neighbours", 20)
        ctx.load_scene(sc, seed=seed)
        ctx.init()
        for it in range(2):
            for name, step, colour in KERNELS:
                for k in range(21):
                    ctx.set_option("site%d" % k, masks[name][k])
                n4, c, _ = ref.steps(sc, [step], norm4=n4, cost=c, seed=seed)
                ctx.phase(colour, 7)
                m4, mc = ctx.get_state()
                print("  verify %s it %d %s: planes diff %d cost diff %d" % (
                    tag, it, name, int((bits(m4) != bits(n4)).sum()), int((bits(mc) != bits(c)).sum())))


if __name__ == "__main__":
    for color in (False, True):
        m = probe(color)
        verify(color, m)
