"""SURVEY.md §8f rows built in host C++ (gipuma_b200/csrc/gpm_host.cpp): camera preparation, view selection, .dmb files.
All of it runs without a GPU.  The NumPy restatement in gipuma_b200/scene.py (checked against OpenCV in test_scene.py) is
the checker for f1/f2; f3 is checked against the byte layout of fileIoUtils.h:320-368."""
import struct

import numpy as np
import pytest

from gipuma_b200 import api, scene as S


def test_prepare_cameras_matches_numpy_restatement():
    Ps = S._dtu_Ps()[:12]
    want = S.prepare_cameras(Ps, cam_scale=2.0)
    got = api.prepare_cameras(Ps, cam_scale=2.0)
    for w, g in zip(want, got):
        for name, tol in (("K", 1e-3), ("K_inv", 1e-7), ("R", 1e-6), ("R_orig_inv", 1e-6), ("M_inv", 1e-7)):
            assert np.allclose(np.array(getattr(g, name)).reshape(3, 3), getattr(w, name), atol=tol, rtol=1e-5), name
        assert np.allclose(np.array(g.t), w.t, atol=2e-2, rtol=1e-5)            # translations are O(100) mm in float32
        assert np.allclose(np.array(g.C), w.C, atol=2e-2, rtol=1e-5)
        assert np.allclose(np.array(g.P_col34), w.P[:, 3], rtol=1e-5, atol=1e-1)
        assert g.fx == pytest.approx(w.fx, rel=1e-6) and g.alpha == pytest.approx(w.alpha, rel=1e-6)
        assert g.baseline == pytest.approx(0.54)


def test_reference_camera_canonical():
    got = api.prepare_cameras(S._dtu_Ps()[:3])
    R0 = np.array(got[0].R).reshape(3, 3)
    assert np.allclose(R0, np.eye(3), atol=1e-6) and np.allclose(np.array(got[0].t), 0, atol=1e-3)


def test_select_views_matches_numpy_restatement():
    Ps = S._dtu_Ps()
    cams_np = S.prepare_cameras(Ps)
    prm = S.AlgorithmParameters(min_angle=10, max_angle=30, max_views=9)
    want = S.select_views(cams_np, 1600, 1200, prm)
    cams_c = api.prepare_cameras(Ps)
    got, (dmin, dmax) = api.select_views(cams_c, 1600, 1200, 10.0, 30.0, 9)
    assert got == want
    assert 0 < dmin < dmax            # main.cpp:470-474 range estimate from baselines and angles
    all_views, _ = api.select_views(cams_c, 1600, 1200, 10.0, 30.0, 64)
    assert len(all_views) == 29


def test_dmb_layout_and_roundtrip(tmp_path):
    depth = np.arange(12, dtype=np.float32).reshape(3, 4) + 0.5
    p = str(tmp_path / "disp.dmb")
    api.write_dmb(p, depth)
    raw = open(p, "rb").read()
    assert struct.unpack("<4i", raw[:16]) == (1, 3, 4, 1)                      # type=1 (float), h, w, nb  (fileIoUtils.h:346-353)
    assert np.array_equal(np.frombuffer(raw[16:], dtype="<f4").reshape(3, 4), depth)
    assert np.array_equal(api.read_dmb(p), depth)
    n4 = np.random.default_rng(0).normal(size=(5, 7, 4)).astype(np.float32)
    pd, pn = str(tmp_path / "d.dmb"), str(tmp_path / "n.dmb")
    api.write_result_dmb(pd, pn, n4)
    assert np.array_equal(api.read_dmb(pd), n4[..., 3])
    assert np.array_equal(api.read_dmb(pn), n4[..., :3])
    assert struct.unpack("<4i", open(pn, "rb").read(16)) == (1, 5, 7, 3)       # writeDmbNormal :320-343
    with pytest.raises(api.GipumaError):
        api.read_dmb(str(tmp_path / "missing.dmb"))
