"""f2: the reference-view batch driver (gpm_batch_run, host C++) — one process, a worker per device, a shared page-locked image
cache, reference views from a common queue; per reference view the cameras are re-based (cameraGeometryUtils.h:174-353), the
source views selected (main.cpp:430-499) and the job run.  Every job's output must equal, bit for bit, what a hand-made
single-context run of the same reference view produces; the .dmb files must round-trip (fileIoUtils.h:247-368)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _dataset(rows=96, cols=128, n_cams=24):
    """Images of the first `n_cams` DTU positions looking at one height field (world = frame of camera 0)."""
    from gipuma_b200 import scene as S
    Ps = [S.load_dtu_projections()[i] for i in range(n_cams)]
    scale = 1600.0 / cols
    cams = S.prepare_cameras(Ps, scale)
    hf = S.HeightField(z0=550.0, unit=550.0 / cams[0].fx, seed=77)
    images = np.stack([S.render_view(c, rows, cols, hf)[0] for c in cams])
    return Ps, scale, images


def _single_job(images, Ps, scale, ref, params, max_views, min_angle, max_angle, seed):
    """The same job through one context, cameras and view subset from the C-ABI host functions."""
    from gipuma_b200 import api
    n, H, W = images.shape
    order = [ref] + [i for i in range(n) if i != ref]
    cams = api.prepare_cameras([Ps[i] for i in order], scale)
    subset, _ = api.select_views(cams, W, H, min_angle, max_angle, max_views)
    p = api.pack_params(params)
    p.min_disparity = np.float32(cams[0].f) * np.float32(cams[0].baseline) / np.float32(p.depthMax)
    p.max_disparity = np.float32(cams[0].f) * np.float32(cams[0].baseline) / np.float32(p.depthMin)
    lib = api.load_library()
    with api.Context(W, H, max(1, len(subset))) as ctx:
        ctx._check(lib.gpm_set_params(ctx.h, C.byref(p)))
        ctx._check(lib.gpm_set_reference(ctx.h, C.c_void_p(images[ref].ctypes.data), 0, 0, C.byref(cams[0])))
        for v, idx in enumerate(subset):
            ctx._check(lib.gpm_set_view(ctx.h, v, C.c_void_p(images[order[idx]].ctypes.data), 0, 0, C.byref(cams[idx])))
        ctx.set_num_views(len(subset))
        ctx.set_rng(seed)
        ctx.run()
        return ctx.get_state() + (len(subset),)


def test_batch_driver_equals_single_jobs_and_writes_dmb(tmp_path):
    import torch
    from gipuma_b200 import api, scene as S
    Ps, scale, images = _dataset()
    params = S.AlgorithmParameters(box_hsize=11, box_vsize=11, iterations=2, n_best=3, cost_comb=S.COMB_BEST_N, gamma=10.0)
    params.depthMin, params.depthMax = 300.0, 800.0
    refs = [0, 7, 13, 20]
    devices = list(range(min(2, torch.cuda.device_count())))
    n4, cost, st = api.batch_run(images, Ps, params, refs, devices=devices, cam_scale=scale, min_angle=5.0, max_angle=45.0,
                                 max_views=6, seed=4242, out_dir=str(tmp_path))
    assert st["jobs_done"] == len(refs) and all(d in devices for d in st["device"])
    for j, ref in enumerate(refs):
        s4, sc, nviews = _single_job(images, Ps, scale, ref, params, 6, 5.0, 45.0, 4242)
        assert st["views"][j] == nviews and nviews >= 1
        assert bits_equal(n4[j], s4) == 0 and bits_equal(cost[j], sc) == 0
        d = api.read_dmb(os.path.join(str(tmp_path), "%08d" % ref, "disp.dmb"))
        nm = api.read_dmb(os.path.join(str(tmp_path), "%08d" % ref, "normals.dmb"))
        assert bits_equal(d, n4[j][..., 3]) == 0 and bits_equal(nm, n4[j][..., :3]) == 0


def test_batch_driver_reports_errors():
    from gipuma_b200 import api, scene as S
    Ps, scale, images = _dataset(rows=64, cols=96, n_cams=4)
    params = S.AlgorithmParameters(box_hsize=9, box_vsize=9, iterations=1, n_best=2, cost_comb=S.COMB_BEST_N)
    params.depthMin, params.depthMax = 300.0, 800.0
    with pytest.raises(api.GipumaError):                     # no camera within 0.1 .. 0.2 degrees of the reference
        api.batch_run(images, Ps, params, [0], cam_scale=scale, min_angle=0.1, max_angle=0.2, max_views=4)
