"""GPU parity proper: gipuma_b200 (through the C-ABI) vs golden outputs of the pinned reference build.

The fixtures under tests/golden/ were produced by tools/make_golden.py with oracle/_ref (the reference's own
gipuma.cu compiled for sm_100a with pins P1-P3) on a B200.  Bar: BIT-EXACT (float bit patterns) at every level
— initial planes and costs, the three black kernels of iteration 1 (run here as ONE fused launch), the whole
first iteration, and the final runcuda() output.  The north-star tolerance (1e-4 depth, 1e-3 normals) is
therefore met with zero slack.
"""
import numpy as np
import pytest

from conftest import bits_equal, golden_names

pytestmark = pytest.mark.gpu


def _ctx(sc, **opts):
    from gipuma_b200 import api
    ctx = api.Context(sc.cols, sc.rows, sc.n_views)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.load_scene(sc, seed=0xC0FFEE)
    return ctx


@pytest.mark.parametrize("name", golden_names())
def test_init_matches_reference_bit_for_bit(golden, name):
    sc, z = golden[name]
    with _ctx(sc) as ctx:
        ctx.init()
        n4, c = ctx.get_state()
    assert bits_equal(n4, z["init_norm4"]) == 0          # gipuma_init_cu2: curand_init + random plane (gipuma.cu:996-1036)
    assert bits_equal(c, z["init_cost"]) == 0            # texture-path cost (gipuma.cu:1040-1049)


@pytest.mark.parametrize("name", golden_names())
def test_cost_eval_matches_reference_sweep_path(golden, name):
    sc, z = golden[name]
    with _ctx(sc) as ctx:
        if sc.params.color_processing:
            # the harness's own float4 cost kernel compiled to the x-term-first rounding (its float one as well, which is
            # the propagation kernels' form and gpm_cost_eval's default for float images)
            ctx.set_option("cost_variant", 1)
        c = ctx.cost_eval(z["init_norm4"])
    assert bits_equal(c, z["init_planes_sweep_cost"]) == 0


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("opts", [{}, {"prune": 0}, {"dedupe": 0}, {"prune": 0, "dedupe": 0, "nwarps": 4}])
def test_fused_black_sweep_matches_three_reference_kernels(golden, name, opts):
    sc, z = golden[name]
    with _ctx(sc, **opts) as ctx:
        ctx.set_option("trust_state", 1)
        ctx.set_state(z["init_norm4"], z["init_cost"])
        ctx.phase(0, 7)                                   # close + far + refine in one launch
        n4, c = ctx.get_state()
    assert bits_equal(n4, z["black_norm4"]) == 0
    assert bits_equal(c, z["black_cost"]) == 0


@pytest.mark.parametrize("name", golden_names())
def test_separate_phase_launches_match_too(golden, name):
    sc, z = golden[name]
    with _ctx(sc) as ctx:
        ctx.set_state(z["init_norm4"], z["init_cost"])
        for mask in (1, 2, 4):
            ctx.phase(0, mask)
        n4, c = ctx.get_state()
        assert bits_equal(n4, z["black_norm4"]) == 0 and bits_equal(c, z["black_cost"]) == 0
        for mask in (1, 2, 4):
            ctx.phase(1, mask)
        n4, c = ctx.get_state()
    assert bits_equal(n4, z["iter1_norm4"]) == 0 and bits_equal(c, z["iter1_cost"]) == 0


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("opts", [{}, {"prune": 0, "dedupe": 0, "memo": 0}, {"memo": 0}, {"prune": 0}])
def test_full_run_matches_reference_runcuda(golden, name, opts):
    from gipuma_b200 import api
    sc, z = golden[name]
    ls, ms, st = api.runcuda(sc, seed=0xC0FFEE, options=opts)
    assert bits_equal(ls.norm4, z["final_norm4"]) == 0    # world normal + depth (gipuma_compute_disp)
    assert bits_equal(ls.c, z["final_cost"]) == 0
    assert ms > 0 and st["launches"] == 2 + 2 * sc.params.iterations + 1
    # and, redundantly, the tolerance BASELINE.json states
    d, dr = ls.norm4[..., 3], z["final_norm4"][..., 3]
    assert np.all(np.abs(d - dr) <= 1e-4 * np.maximum(1.0, np.abs(dr)))
    assert np.all(np.abs(ls.norm4[..., :3] - z["final_norm4"][..., :3]) <= 1e-3)


def test_exact_pruning_and_dedupe_actually_skip_work(golden):
    from gipuma_b200 import api
    if not golden_names():
        pytest.skip("no golden fixtures")
    name = golden_names()[0]
    sc, z = golden[name]
    _, _, st_on = api.runcuda(sc, seed=0xC0FFEE)
    _, _, st_off = api.runcuda(sc, seed=0xC0FFEE, options={"prune": 0, "dedupe": 0, "memo": 0})
    assert st_off["pruned"] == 0 and st_off["pairs"] == st_off["pairs_full"]
    assert st_on["pairs"] < st_off["pairs"] and st_on["skipped"] > st_off["skipped"]


def test_drop_in_boundary_through_the_reference_headers(golden):
    """The same OpenCV-free main.cpp stand-in that drives the reference, linked against our runcuda()."""
    import os
    from oracle import pyref
    if not os.path.exists(os.path.join(pyref.REF_DIR, "libhx_dropin.so")):
        pytest.skip("oracle/_ref/libhx_dropin.so not built (needs the reference headers)")
    h = pyref.Harness("dropin")
    assert h.backend == "dropin"
    for name in golden_names():
        sc, z = golden[name]
        n4, c, printed_s, wall_ms = h.run(sc, seed=0xC0FFEE)
        assert bits_equal(n4, z["final_norm4"]) == 0 and bits_equal(c, z["final_cost"]) == 0
        assert printed_s > 0            # "Total time needed for computation" line is printed like the reference's
