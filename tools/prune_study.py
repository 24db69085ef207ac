"""How much earlier could a rejected hypothesis be cut off?  CPU study with the C restatement (no GPU).

For every hypothesis of a PatchMatch run on a small DTU-like scene it records all per-sample cost terms of all views and
asks: given only a SUBSET of the window samples, is the (exact, monotone) lower bound of the combined cost already
>= the pixel's current cost?  Subsets compared (each half of the window):
  prefix   the first half of the window in accumulation order (columns 0..3 of 8) — what k_sweep uses today
  centre   the centre columns (2..5)
  heavy    the half of the samples with the largest adaptive support weights
Usage:  python tools/prune_study.py [rows cols views iterations]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gipuma_b200 import scene as S                      # noqa: E402
from oracle import pyoracle                             # noqa: E402
from oracle.pyoracle import GpmCamera, GpmParams        # noqa: E402

MAXCOST = 1000.0


def combine(per_view, n_best):
    c = np.minimum(per_view, MAXCOST)
    valid = int((c < MAXCOST).sum())
    c = np.sort(c)
    nb = min(valid, n_best)
    return float(c[:nb].mean()) if nb > 0 else MAXCOST


def main():
    a = [int(v) for v in sys.argv[1:5]] + [64, 96, 10, 5][len(sys.argv) - 1:]
    rows, cols, views, iters = a
    sc = S.make_config(2, rows=rows, cols=cols, n_views=views, iterations=iters, seed=31)
    o = pyoracle.Oracle(sc)
    lib = o.lib
    fp = C.POINTER(C.c_float)
    lib.gpo_multiview_terms.restype = C.c_float
    lib.gpo_multiview_terms.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(GpmParams), C.POINTER(GpmCamera),
                                        C.POINTER(GpmCamera), fp, C.POINTER(fp), C.c_int, C.c_int, fp, fp, C.c_int]
    p = sc.params
    side = (p.box_hsize - 1) // 2 // 2 * 2 // 2 + 1 if False else len(range(-((p.box_hsize - 1) // 2), (p.box_hsize - 1) // 2 + 1, 2))
    ns = side * side
    V = sc.n_views
    terms = np.zeros((V, ns), np.float32)
    common = o._common()

    def evaluate(px, py, plane):
        pl = np.ascontiguousarray(plane, np.float32)
        c = lib.gpo_multiview_terms(*common, px, py, pl.ctypes.data_as(fp), terms.ctypes.data_as(fp), ns)
        return float(c), terms.copy()

    # sample subsets (window order: x offset outer, y offset inner)
    col = np.repeat(np.arange(side), side)
    prefix = col < side // 2
    centre = (col >= side // 4) & (col < side // 4 + side // 2)
    img = sc.images[0]
    rad = (p.box_hsize - 1) // 2

    def heavy_mask(px, py):
        xs = np.clip(px + np.arange(-rad, rad + 1, 2), 0, cols - 1)
        ys = np.clip(py + np.arange(-rad, rad + 1, 2), 0, rows - 1)
        w = np.exp(-np.abs(img[np.ix_(ys, xs)].T - img[py, px]) / p.gamma).reshape(-1)      # [x outer, y inner]
        order = np.argsort(-w, kind="stable")
        m = np.zeros(ns, bool)
        m[order[:ns // 2]] = True
        return m, w

    rng = np.random.default_rng(5)
    planes = np.zeros((rows, cols, 4), np.float32)
    st = np.zeros(6, np.uint32)
    for y in range(rows):
        for x in range(cols):
            st[:5] = rng.integers(1, 2 ** 32 - 1, 5, dtype=np.uint64).astype(np.uint32)
            planes[y, x] = o.random_plane(x, y, st)
    cost = o.cost_eval(planes)
    ref = pyoracle.pack_camera(sc.cameras[0]) if hasattr(pyoracle, "pack_camera") else None
    fb = sc.cameras[0].f * sc.cameras[0].baseline
    stats = {}

    def record(it, kind, cnow, c, t, hm, w):
        key = (it, kind)
        s = stats.setdefault(key, dict(n=0, rejected=0, prefix=0, centre=0, heavy=0, wfrac_prefix=0.0, wfrac_centre=0.0, wfrac_heavy=0.0))
        s["n"] += 1
        # per-view elimination after the first half (prefix): a view whose partial cost already exceeds the n_best-th smallest
        # UPPER bound (partial + remaining weight x max dissimilarity) of all views, or n_best x the current cost, cannot enter
        # the result; how many (view, second-half) evaluations would that save among hypotheses the whole-hypothesis bound keeps?
        dmax = (1.0 - p.alpha) * p.tau_color + p.alpha * p.tau_gradient
        lbv = t[:, prefix].sum(axis=1)
        ubv = lbv + dmax * float(w[~prefix].sum())
        kth = np.sort(ubv)[min(p.n_best, len(ubv)) - 1]
        hyp_pruned = combine(lbv, p.n_best) >= cnow
        dead = (lbv > kth) | (lbv >= p.n_best * cnow)
        s.setdefault("views_total", 0);  s.setdefault("views_dead", 0);  s.setdefault("hyp_kept", 0)
        if not hyp_pruned:
            s["hyp_kept"] += 1;  s["views_total"] += len(lbv);  s["views_dead"] += int(dead.sum())
        if c < cnow:
            return
        s["rejected"] += 1
        for name, m in (("prefix", prefix), ("centre", centre), ("heavy", hm)):
            lb = combine(t[:, m].sum(axis=1), p.n_best)
            s[name] += lb >= cnow
            s["wfrac_" + name] += float(w[m].sum() / w.sum())

    off = [(0, -1), (0, 1), (-1, 0), (1, 0), (0, -5), (0, 5), (-5, 0), (5, 0)]
    depth_of = o.lib.gpo_plane_depth
    depth_of.restype = C.c_float
    for it in range(iters):
        for colour in (0, 1):
            for py in range(rows):
                for px in range(cols):
                    if ((px + py) & 1) != colour:
                        continue
                    hm, w = heavy_mask(px, py)
                    cnow, pnow = float(cost[py, px]), planes[py, px].copy()
                    for dx, dy in off:
                        qx, qy = px + dx, py + dy
                        if not (0 <= qx < cols and 0 <= qy < rows):
                            continue
                        cand = planes[qy, qx]
                        if np.array_equal(cand.view(np.uint32), pnow.view(np.uint32)):
                            continue
                        c, t = evaluate(px, py, cand)
                        record(it, "prop", cnow, c, t, hm, w)
                        d = o.plane_depth(cand, px, py) if hasattr(o, "plane_depth") else None
                        if c < cnow and (d is None or p.depthMin <= d <= p.depthMax):
                            cnow, pnow = c, cand.copy()
                    # refinement (perturbation scales of gipuma.cu:958-992; plain RNG — the statistics are what matters)
                    deltaZ, deltaN = p.max_disparity / 2.0, 1.0
                    while deltaZ >= 0.01:
                        n = pnow[:3] + rng.uniform(-deltaN, deltaN, 3).astype(np.float32)
                        n /= np.linalg.norm(n)
                        if n[2] > 0:
                            n = -n
                        depth = o.plane_depth(pnow, px, py)
                        disp = float(np.clip(fb / depth + rng.uniform(-deltaZ, deltaZ), p.min_disparity, p.max_disparity))
                        cand = np.array([n[0], n[1], n[2], 0.0], np.float32)
                        cand[3] = o.plane_d(cand, px, py, fb / disp)
                        c, t = evaluate(px, py, cand)
                        record(it, "refine", cnow, c, t, hm, w)
                        if c < cnow:
                            cnow, pnow = c, cand
                        deltaZ /= 10.0
                        deltaN /= 4.0
                    cost[py, px], planes[py, px] = cnow, pnow
        for kind in ("prop", "refine"):
            s = stats.get((it, kind))
            if not s:
                continue
            r = max(1, s["rejected"])
            print("iter %d %-6s: %6d evaluated, %5.1f%% rejected; of the rejected cut after half the samples: "
                  "prefix %5.1f%% (weight %.2f)  centre %5.1f%% (%.2f)  heavy %5.1f%% (%.2f)" % (
                      it + 1, kind, s["n"], 100.0 * s["rejected"] / s["n"], 100.0 * s["prefix"] / r, s["wfrac_prefix"] / r,
                      100.0 * s["centre"] / r, s["wfrac_centre"] / r, 100.0 * s["heavy"] / r, s["wfrac_heavy"] / r), flush=True)
            print("        per-view elimination after the first half: %d hypotheses survive the whole-hypothesis bound; %.1f%% of their views are dead"
                  % (s.get("hyp_kept", 0), 100.0 * s.get("views_dead", 0) / max(1, s.get("views_total", 0))), flush=True)


if __name__ == "__main__":
    main()
