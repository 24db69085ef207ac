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


# ---- property tests (hypothesis) ----------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st      # noqa: E402


def _random_rotation(rng):
    q, r = np.linalg.qr(rng.normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(min_value=0, max_value=2 ** 31 - 1))
def test_prepare_cameras_recovers_random_pinhole_rigs(seed):
    """P = K [R | t] with random intrinsics (incl. skew) and poses: the C++ RQ restatement (cameraGeometryUtils.h:174-353) must
    return an upper-triangular K with positive diagonal, a proper rotation, and K [R | t] must reproduce the re-based
    projection; the reference camera comes out canonical."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 6))
    Ps = []
    for _ in range(n):
        fx, fy = rng.uniform(500, 4000, 2)
        K = np.array([[fx, rng.uniform(-2, 2), rng.uniform(200, 1500)], [0, fy, rng.uniform(200, 1200)], [0, 0, 1.0]])
        R = _random_rotation(rng)
        t = rng.uniform(-300, 300, 3)
        Ps.append(rng.uniform(0.5, 2.0) * K @ np.hstack([R, t[:, None]]))            # arbitrary positive scale
    cams = api.prepare_cameras(Ps)
    for i, c in enumerate(cams):
        K = np.array(c.K).reshape(3, 3)
        R = np.array(c.R).reshape(3, 3)
        assert abs(K[1, 0]) < 1e-3 and abs(K[2, 0]) < 1e-5 and abs(K[2, 1]) < 1e-5 and K[0, 0] > 0 and K[1, 1] > 0
        assert K[2, 2] == pytest.approx(1.0, abs=1e-5)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-4) and np.linalg.det(R) == pytest.approx(1.0, abs=1e-3)
        assert np.allclose(K @ np.array(c.K_inv).reshape(3, 3), np.eye(3), atol=1e-3)
    R0 = np.array(cams[0].R).reshape(3, 3)
    assert np.allclose(R0, np.eye(3), atol=1e-5) and np.allclose(np.array(cams[0].t), 0, atol=1e-2)
    # relative pose preserved by the re-basing: R_i R_0^T of the inputs == R_i of the outputs
    Rin = []
    for P in Ps:
        K, R, _ = S.decompose_projection(np.asarray(P, dtype=np.float64))[:3]
        Rin.append(R)
    for i in range(1, n):
        assert np.allclose(np.array(cams[i].R).reshape(3, 3), Rin[i] @ Rin[0].T, atol=2e-3)


@settings(max_examples=20, deadline=None)
@given(rows=st.integers(1, 40), cols=st.integers(1, 40), ch=st.sampled_from([1, 3]), seed=st.integers(0, 2 ** 31 - 1))
def test_dmb_roundtrip_any_shape(rows, cols, ch, seed):
    import os
    import tempfile
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(rows, cols) if ch == 1 else (rows, cols, ch)).astype(np.float32)
    p = os.path.join(tempfile.mkdtemp(prefix="gpm_dmb_"), "x.dmb")
    api.write_dmb(p, a)
    raw = open(p, "rb").read()
    assert len(raw) == 16 + 4 * rows * cols * ch and struct.unpack("<4i", raw[:16]) == (1, rows, cols, ch)
    assert np.array_equal(api.read_dmb(p), a)
    open(p, "wb").write(raw[:-4])                                                    # truncated payload
    with pytest.raises(api.GipumaError):
        api.read_dmb(p)


def test_select_views_is_monotone_in_max_views_and_respects_the_angle_window():
    Ps = S._dtu_Ps()
    cams = api.prepare_cameras(Ps)
    prev = []
    for mv in (1, 3, 9, 20, 64):
        sub, _ = api.select_views(cams, 1600, 1200, 10.0, 30.0, mv)
        assert len(sub) <= mv and len(set(sub)) == len(sub) and 0 not in sub
        assert sub[:len(prev)] == prev                                                # deterministic prefix order
        prev = sub
    none, _ = api.select_views(cams, 1600, 1200, 89.0, 90.0, 9)                       # nobody that oblique on the DTU rig
    assert none == []
    ang = S.view_angles(S.prepare_cameras(Ps), 1600, 1200)
    for i in prev:
        assert 10.0 <= np.degrees(ang[i]) <= 30.0


def test_prepare_cameras_matches_opencv_getCameraParameters_fixture():
    """f1 pinned: tests/golden/camera_prep_dtu.npz holds what the reference's getCameraParameters
    (cameraGeometryUtils.h:174-353) computes for the 64 DTU .P files — produced by tools/make_camera_fixture.py, which runs
    the same OpenCV routines (decomposeProjectionMatrix, Mat::inv LU / SVD, determinant) on float32 matrices through cv2.
    gpm_prepare_cameras (host C++, double precision, no OpenCV) must reproduce every Camera_cu field to float32 rounding:
    1e-6 of the field's scale over the rig (the reference's own float32 pipeline is only that precise), for both scale factors."""
    import os
    from conftest import GOLDEN_DIR
    from gipuma_b200 import api
    z = np.load(os.path.join(GOLDEN_DIR, "camera_prep_dtu.npz"))
    Ps = [z["P"][i] for i in range(z["P"].shape[0])]
    for scale, tag in ((1.0, "s1"), (0.5, "s05")):
        cams = api.prepare_cameras(Ps, scale)
        assert len(cams) == 64
        for key in ("K", "K_inv", "R", "t", "C", "M_inv", "R_orig_inv", "P_col34"):
            ref = z["%s_%s" % (tag, key)]
            ours = np.stack([np.array(getattr(c, key)[:], np.float32).reshape(ref[0].shape) for c in cams])
            assert np.abs(ours - ref).max() <= 1e-6 * np.abs(ref).max(), key
        for key in ("fx", "fy", "f", "alpha", "baseline"):
            ref = z["%s_%s" % (tag, key)]
            ours = np.array([getattr(c, key) for c in cams], np.float32)
            assert np.abs(ours - ref).max() <= 1e-6 * np.abs(ref).max(), key
