"""Host-side input preparation: camera restatement (cameraGeometryUtils.h:174-353), view selection
(main.cpp:430-499) and the synthetic renderer."""
import numpy as np
import pytest

from gipuma_b200 import scene as S


def test_decomposition_reconstructs_projection():
    for k in (0, 24, 63):
        P = S.load_dtu_projections()[k]
        K, R, C = S.decompose_projection(P)
        assert np.allclose(np.tril(K, -1), 0)
        assert np.all(np.diag(K) > 0)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-10) and np.linalg.det(R) > 0
        P2 = K @ np.hstack([R, (-R @ C)[:, None]])
        assert np.allclose(P2 / P2[2, 3], P / P[2, 3], rtol=1e-9, atol=1e-7)


def test_decomposition_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    P = S.load_dtu_projections()[24]
    K, R, T = cv2.decomposeProjectionMatrix(P)[:3]
    K2, R2, C2 = S.decompose_projection(P)
    assert np.allclose(K / K[2, 2], K2 / K2[2, 2], atol=1e-8)
    assert np.allclose(R, R2, atol=1e-12)
    assert np.allclose((T[:3] / T[3]).ravel(), C2, atol=1e-8)


def test_reference_camera_is_canonical_after_rebase():
    cams = S.prepare_cameras(S._dtu_Ps())
    c0 = cams[0]
    assert np.allclose(c0.R, np.eye(3), atol=1e-5) and np.allclose(c0.t, 0, atol=1e-2)
    assert np.allclose(c0.P[:, :3], c0.K, rtol=1e-5, atol=1e-3)
    assert np.allclose(c0.M_inv @ c0.P[:, :3], np.eye(3), atol=1e-4)
    assert c0.alpha == pytest.approx(c0.fx / c0.fy)
    assert c0.baseline == pytest.approx(0.54)            # cameraGeometryUtils.h:305


def test_scale_k():
    K = np.array([[2892.0, 0, 823.0], [0, 2883.0, 619.0], [0, 0, 1]])
    Ks = S.scale_K(K, 5.0)
    assert Ks[0, 0] == pytest.approx(578.4) and Ks[1, 2] == pytest.approx(123.8) and Ks[2, 2] == 1


def test_view_selection_angle_filter():
    cams = S.prepare_cameras(S._dtu_Ps())
    prm = S.AlgorithmParameters(min_angle=10, max_angle=30, max_views=9)
    ang = np.degrees(S.view_angles(cams, 1600, 1200))
    sel = S.select_views(cams, 1600, 1200, prm)
    assert len(sel) == 9 and all(10 < ang[i] < 30 for i in sel) and sel == sorted(sel)
    sel30 = S.select_views(cams, 1600, 1200, prm, n_views=30)
    assert len(sel30) == 30 and len(set(sel30)) == 30 and 0 not in sel30
    assert sum(10 < ang[i] < 30 for i in sel30) == 29            # 29 DTU positions pass from position 25


def test_rendered_views_are_photo_consistent():
    sc = S.make_config(1, rows=120, cols=160)
    assert sc.images.shape == (3, 120, 160) and sc.images.dtype == np.float32
    assert np.array_equal(sc.images, np.rint(sc.images)) and sc.images.min() >= 0 and sc.images.max() <= 255
    c0, c1 = sc.cameras[0], sc.cameras[1]
    ys, xs = np.mgrid[10:110:5, 10:150:5]
    d = sc.gt_depth[ys, xs].astype(np.float64)
    pt = np.stack([d * xs - c0.P[0, 3], d * ys - c0.P[1, 3], d - c0.P[2, 3]], -1)
    X = pt @ c0.M_inv.astype(np.float64).T
    q = X @ c1.P[:, :3].astype(np.float64).T + c1.P[:, 3]
    u, v = q[..., 0] / q[..., 2], q[..., 1] / q[..., 2]
    ok = (u > 1) & (u < 158) & (v > 1) & (v < 118)
    assert ok.sum() > 100
    a = sc.images[1][np.rint(v[ok]).astype(int), np.rint(u[ok]).astype(int)]
    b = sc.images[0][ys[ok], xs[ok]]
    assert np.abs(a - b).mean() < 0.4 * np.abs(a - np.roll(b, 7)).mean()


@pytest.mark.parametrize("k,shape,V,box", [(1, (240, 320), 2, 15), (2, (1200, 1600), 10, 15), (3, (1200, 1600), 30, 25),
                                           (4, (480, 640), 47, 11), (5, (2400, 3200), 64, 15)])
def test_config_table_matches_baseline(k, shape, V, box):
    # geometry only (small render) — the BASELINE.json table: size, views, window
    sc = S.make_config(k, rows=shape[0] // 20, cols=shape[1] // 20)
    assert sc.n_views == V and sc.params.box_hsize == box and len(sc.cameras) == V + 1
    assert sc.params.max_disparity > sc.params.min_disparity > 0
    f, b = np.float32(sc.cameras[0].f), np.float32(0.54)
    assert sc.params.max_disparity == pytest.approx(float(f * b / np.float32(sc.params.depthMin)))   # main.cpp:906
