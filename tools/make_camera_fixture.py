#!/usr/bin/env python
"""tests/golden/camera_prep_dtu.npz — what the reference's getCameraParameters (cameraGeometryUtils.h:174-353) computes for
the 64 DTU projection matrices shipped with it (data/dtu/calib), reference camera = position 25 first.

The reference's own code cannot be built here (OpenCV C++ headers are absent), so this script executes the SAME OpenCV
routines through cv2 on float32 matrices, statement by statement:
    decomposeProjectionMatrix (:252)            cv2.decomposeProjectionMatrix on the Mat_<float> P
    C = T[0:3] / T[3],  t = -R C (:259-260)
    transform = [R0|t0]^-1 (:109-115, Mat::inv = LU)          cv2.invert(..., DECOMP_LU)
    K scaled (:136-147), K^-1 (:292, LU)
    R_orig_inv = R.inv(DECOMP_SVD) (:297)
    transformCamera (:117-134): [R|t] * transform, P = K_ref * (.)[0:3], C from the 3x3 minors of P (:22-49, cv::determinant)
    M_inv = P[:, 0:3].inv() (:301, LU),  P_col34 = P[:, 3] (:337-343),  fx, fy, f, alpha (:313-318), baseline 0.54 (:305)
The committed fixture pins gpm_prepare_cameras / scene.prepare_cameras (tests/test_host_rows.py).
Run:  python tools/make_camera_fixture.py        (needs cv2; the test only needs the .npz)
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gipuma_b200 import scene as S  # noqa: E402

f32 = np.float32


def mat(a):
    return np.ascontiguousarray(a, dtype=f32)


def get_camera_center(P):                      # cameraGeometryUtils.h:22-49
    def sub(idx):
        return mat(P[:, idx])
    C = np.zeros((4, 1), f32)
    C[0, 0] = f32(cv2.determinant(sub([1, 2, 3])))
    C[1, 0] = -f32(cv2.determinant(sub([0, 2, 3])))
    C[2, 0] = f32(cv2.determinant(sub([0, 1, 3])))
    C[3, 0] = -f32(cv2.determinant(sub([0, 1, 2])))
    return C


def transformation_matrix(R, t):               # :93-100
    M = np.eye(4, dtype=f32)
    M[0:3, 0:3] = R
    M[0:3, 3:4] = t
    return M


def scale_k(K, s):                             # :136-147
    K = K.copy()
    K[0, 0] = K[0, 0] / f32(s);  K[1, 1] = K[1, 1] / f32(s);  K[0, 2] = K[0, 2] / f32(s);  K[1, 2] = K[1, 2] / f32(s)
    return K


def get_camera_parameters(Ps, scale=1.0):
    n = len(Ps)
    K, R, t = [], [], []
    for P in Ps:
        k, r, T = cv2.decomposeProjectionMatrix(mat(P))[:3]            # :252 (Mat_<float> in, Mat_<float> out)
        k, r, T = mat(k), mat(r), mat(T)
        C = mat(T[0:3, 0:1] / T[3, 0])                                  # :259
        K.append(k);  R.append(r);  t.append(mat(-(r @ C)))            # :260 (float32 matrix product)
    transform = mat(cv2.invert(transformation_matrix(R[0], t[0]), flags=cv2.DECOMP_LU)[1])      # :282-283
    Kref = scale_k(K[0], scale)                                                                  # :290
    out = []
    for i in range(n):
        Ki = scale_k(K[i], scale)
        Ki_inv = mat(cv2.invert(Ki, flags=cv2.DECOMP_LU)[1])
        R_orig_inv = mat(cv2.invert(R[i], flags=cv2.DECOMP_SVD)[1])                              # :297
        Mt = mat(transformation_matrix(R[i], t[i]) @ transform)                                  # :121-124
        P = mat(Kref @ Mt[0:3, 0:4])                                                             # :127
        Rn, tn = mat(Mt[0:3, 0:3]), mat(Mt[0:3, 3:4])
        C = get_camera_center(P)
        C = mat(C / C[3, 0])                                                                     # :133
        M_inv = mat(cv2.invert(mat(P[:, 0:3]), flags=cv2.DECOMP_LU)[1])                          # :301
        out.append(dict(K=Ki, K_inv=Ki_inv, R=Rn, t=tn[:, 0], C=C[0:3, 0], M_inv=M_inv, R_orig_inv=R_orig_inv, P_col34=P[:, 3],
                        fx=Kref[0, 0], fy=Kref[1, 1], f=Kref[0, 0], alpha=f32(Kref[0, 0] / Kref[1, 1]), baseline=f32(0.54)))
    return out


def main():
    Ps = S._dtu_Ps()                            # position 25 first, then the other 63 in file order
    arrays = {"P": np.stack(Ps).astype(np.float64), "opencv_version": np.array(cv2.__version__)}
    for scale, tag in ((1.0, "s1"), (0.5, "s05")):          # 0.5: what config 5 (3200 x 2400) uses
        cams = get_camera_parameters(Ps, scale)
        for key in ("K", "K_inv", "R", "t", "C", "M_inv", "R_orig_inv", "P_col34"):
            arrays["%s_%s" % (tag, key)] = np.stack([np.asarray(c[key], f32) for c in cams])
        for key in ("fx", "fy", "f", "alpha", "baseline"):
            arrays["%s_%s" % (tag, key)] = np.array([c[key] for c in cams], f32)
    out = os.path.join(ROOT, "tests", "golden", "camera_prep_dtu.npz")
    np.savez_compressed(out, **arrays)
    print(out, os.path.getsize(out), "bytes; OpenCV", cv2.__version__)


if __name__ == "__main__":
    main()
