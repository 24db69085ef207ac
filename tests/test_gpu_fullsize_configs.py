"""Every BASELINE.json configuration at FULL size against the live pinned reference build (oracle/_ref), bit for bit:
configs[2] (dtu_accurate, 1600x1200, 30 views, blocksize 25), configs[3] (templeRing shape, 640x480, 47 views, pin P3 build),
configs[4] (3200x2400, 64 views, pin P3 build) and north_star's 60-view job.  configs[1] is covered by
test_gpu_live_reference.py::test_baseline_config2_full_size_bit_exact_vs_live_reference.

configs[2] needs care: with 1200 rows and blocksize 25 the reference's tile loader leaves the bottom of the last block row's
shared-memory window unloaded (threads outside the image return first, gipuma.cu:1488-1491 vs 1510-1525), so
its pixels in rows >= 1189 read stale shared memory.  The damage travels upwards by at most 6 rows per colour pass
(close + far propagation) = 96 rows in 8 iterations, so rows < 1088 must still be identical; and the same job at 1600 x 1216
(whole tile rows, no undefined behaviour in the reference) must be identical everywhere."""
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu
WORKERS = max(1, min(32, len(os.sched_getaffinity(0))))


def _ref(n_views):
    from oracle import pyref
    which = "ref64" if n_views > 32 else "ref"
    path = os.path.join(pyref.REF_DIR, {"ref": "libhx_ref.so", "ref64": "libhx_ref64.so"}[which])
    if not os.path.exists(path):
        pytest.skip("pinned reference build not present")
    return pyref.Harness(which)


def _both(sc):
    from gipuma_b200 import api
    ref = _ref(sc.n_views)
    r_n4, r_c, printed_s, _ = ref.run(sc)
    ls, ms, _ = api.runcuda(sc)
    return ls, r_n4, r_c, ms / 1e3, printed_s


def test_config3_dtu_accurate_full_size_tile_aligned_bit_exact():
    from gipuma_b200 import scene as S
    sc = S.make_config(3, rows=1216, cols=1600, workers=WORKERS)          # 38 whole tile rows: the reference loads every window
    ls, r_n4, r_c, ours_s, ref_s = _both(sc)
    assert bits_equal(ls.norm4, r_n4) == 0 and bits_equal(ls.c, r_c) == 0
    assert ours_s < ref_s


def test_config3_dtu_accurate_full_size_as_specified():
    from gipuma_b200 import scene as S
    sc = S.make_config(3, workers=WORKERS)                                 # 1600 x 1200: reference UB in the last block row
    ls, r_n4, r_c, ours_s, ref_s = _both(sc)
    safe = 1189 - 6 * 2 * sc.params.iterations - 5
    assert bits_equal(ls.norm4[:safe], r_n4[:safe]) == 0 and bits_equal(ls.c[:safe], r_c[:safe]) == 0
    # below that line only pixels reached by the reference's stale shared memory may differ: a small minority
    diff = (ls.c[safe:].view(np.uint32) != r_c[safe:].view(np.uint32)).mean()
    assert diff < 0.5
    assert ours_s < ref_s


def test_config4_temple_ring_full_size_bit_exact():
    from gipuma_b200 import scene as S
    sc = S.make_config(4, workers=WORKERS)                                 # 640 x 480, 47 views, blocksize 11
    ls, r_n4, r_c, ours_s, ref_s = _both(sc)
    assert bits_equal(ls.norm4, r_n4) == 0 and bits_equal(ls.c, r_c) == 0
    assert ours_s < ref_s


def test_northstar_60_view_job_full_size_bit_exact():
    from gipuma_b200 import scene as S
    sc = S.make_config(6, workers=WORKERS)                                 # 1600 x 1200, 60 views, dtu_fast parameters
    ls, r_n4, r_c, ours_s, ref_s = _both(sc)
    assert bits_equal(ls.norm4, r_n4) == 0 and bits_equal(ls.c, r_c) == 0
    assert ours_s < ref_s


def test_config5_3200x2400_64_views_full_size_bit_exact():
    from gipuma_b200 import scene as S
    sc = S.make_config(5, workers=WORKERS)
    ls, r_n4, r_c, ours_s, ref_s = _both(sc)
    assert bits_equal(ls.norm4, r_n4) == 0 and bits_equal(ls.c, r_c) == 0
    assert ours_s < ref_s
