"""GPM_RNG_STATEFUL — the refinement RNG the reference evidently intended (a per-pixel XORWOW state seeded at initialisation
and advanced by every draw; the reference allocates gs.cs but never writes it, gipuma.cu:1840, 1608, 1702).  There is no
reference output to pin it against bit for bit, so it is pinned statistically against the mode that IS pinned
(GPM_RNG_REFERENCE, bit-exact with the reference build): ground-truth hit rate per iteration on a smooth and on a hard
scene (occluders, texture-less band, sensor noise), determinism, and seed sensitivity."""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _hit_rates(sc, rng_mode, iters, seed=0xC0FFEE):
    """Fraction of pixels within 1 % of the rendered depth after 1 .. iters iterations, and the final mean cost."""
    from gipuma_b200 import api
    rates = []
    out = None
    for k in range(1, iters + 1):
        sc.params.iterations = k
        out, _, _ = api.runcuda(sc, seed=seed, rng_mode=rng_mode)
        d = out.norm4[..., 3]
        rates.append(float((np.abs(d - sc.gt_depth) / sc.gt_depth < 0.01).mean()))
    return rates, float(out.c.mean()), out


@pytest.mark.parametrize("hard", [False, True])
def test_stateful_rng_converges_like_the_reference_stream(hard):
    from gipuma_b200 import api, scene as S
    sc = S.make_config(2, rows=240, cols=320, n_views=6, iterations=5, hard=hard)
    ref_rates, ref_cost, ref_out = _hit_rates(sc, api.GPM_RNG_REFERENCE, 5)
    st_rates, st_cost, st_out = _hit_rates(sc, api.GPM_RNG_STATEFUL, 5)
    # the stateful stream recovers the surface at least as fast as the reference's degenerate zero-state stream (measured on
    # B200, smooth scene: 0.276 0.820 0.969 0.986 0.990 vs 0.320 0.890 0.986 0.993 0.995) and never falls behind it
    for a, b in zip(ref_rates, st_rates):
        assert b > a - 0.02 and abs(a - b) < 0.15, (ref_rates, st_rates)
    assert st_rates[-1] > ref_rates[-1] - 0.01 and st_rates[-1] > (0.60 if hard else 0.95), (ref_rates, st_rates)
    assert all(b2 >= b1 - 0.005 for b1, b2 in zip(st_rates, st_rates[1:]))          # the hit rate does not fall back
    assert st_cost < ref_cost * 1.02                                                  # proper random perturbations find costs at least as low
    assert bits_equal(ref_out.norm4, st_out.norm4) > 0                                # a genuinely different random sequence


def test_stateful_rng_is_deterministic_and_seed_dependent():
    from gipuma_b200 import api, scene as S
    sc = S.make_config(2, rows=128, cols=160, n_views=4, iterations=3)
    a, _, _ = api.runcuda(sc, seed=11, rng_mode=api.GPM_RNG_STATEFUL)
    b, _, _ = api.runcuda(sc, seed=11, rng_mode=api.GPM_RNG_STATEFUL)
    c, _, _ = api.runcuda(sc, seed=12, rng_mode=api.GPM_RNG_STATEFUL)
    assert bits_equal(a.norm4, b.norm4) == 0 and bits_equal(a.c, b.c) == 0
    assert bits_equal(a.norm4, c.norm4) > 0
    # the exact shortcuts stay exact in this mode too (the refinement memo is simply not used: gpm_kernels.cuh, rng_mode == 0)
    d, _, _ = api.runcuda(sc, seed=11, rng_mode=api.GPM_RNG_STATEFUL, options={"memo": 0, "prune": 0, "dedupe": 0})
    assert bits_equal(a.norm4, d.norm4) == 0 and bits_equal(a.c, d.c) == 0


def test_hard_scene_bit_exact_vs_live_reference():
    """The hard variant (what `bench.py --scene hard` times) against the live pinned reference build."""
    import os
    from gipuma_b200 import api, scene as S
    from oracle import pyref
    if not os.path.exists(os.path.join(pyref.REF_DIR, "libhx_ref.so")):
        pytest.skip("pinned reference build not present")
    sc = S.make_config(2, rows=160, cols=224, n_views=7, iterations=3, hard=True, seed=99)
    r_n4, r_c, _, _ = pyref.Harness("ref").run(sc)
    for opts in ({}, {"memo": 0}):
        ls, _, _ = api.runcuda(sc, options=opts)
        assert bits_equal(ls.norm4, r_n4) == 0 and bits_equal(ls.c, r_c) == 0
