"""ctypes loader for the harness libraries under oracle/_ref/ (hx_api.h).

TEST / BENCH INFRASTRUCTURE — the parity checker and the reference timing arm.  Only tests/,
__graft_entry__.smoke() and bench.py's reference/cpu_baseline legs may import this module; the
product package (gipuma_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

(STEP_INIT, STEP_BLACK_CLOSE, STEP_BLACK_FAR, STEP_BLACK_REFINE,
 STEP_RED_CLOSE, STEP_RED_FAR, STEP_RED_REFINE, STEP_COMPUTE_DISP, STEP_BLACK_FUSED, STEP_RED_FUSED) = range(10)


class HxParams(C.Structure):
    _fields_ = [("box_hsize", C.c_int), ("box_vsize", C.c_int),
                ("tau_color", C.c_float), ("tau_gradient", C.c_float),
                ("alpha", C.c_float), ("gamma", C.c_float),
                ("min_disparity", C.c_float), ("max_disparity", C.c_float),
                ("iterations", C.c_int), ("n_best", C.c_int), ("cost_comb", C.c_int),
                ("good_factor", C.c_float), ("color_processing", C.c_int),
                ("depthMin", C.c_float), ("depthMax", C.c_float)]


class HxCamera(C.Structure):
    _fields_ = [("K", C.c_float * 9), ("K_inv", C.c_float * 9), ("R", C.c_float * 9),
                ("R_orig_inv", C.c_float * 9), ("M_inv", C.c_float * 9), ("P", C.c_float * 12),
                ("t", C.c_float * 3), ("C", C.c_float * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("f", C.c_float), ("alpha", C.c_float),
                ("baseline", C.c_float)]


def pack_params(p) -> HxParams:
    return HxParams(p.box_hsize, p.box_vsize, p.tau_color, p.tau_gradient, p.alpha, p.gamma,
                    p.min_disparity, p.max_disparity, p.iterations, p.n_best, p.cost_comb,
                    p.good_factor, int(p.color_processing), p.depthMin, p.depthMax)


def pack_cameras(cams) -> C.Array:
    arr = (HxCamera * len(cams))()
    for a, c in zip(arr, cams):
        a.K[:] = c.K.ravel().tolist()
        a.K_inv[:] = c.K_inv.ravel().tolist()
        a.R[:] = c.R.ravel().tolist()
        a.R_orig_inv[:] = c.R_orig_inv.ravel().tolist()
        a.M_inv[:] = c.M_inv.ravel().tolist()
        a.P[:] = c.P.ravel().tolist()
        a.t[:] = c.t.ravel().tolist()
        a.C[:] = c.C.ravel().tolist()
        a.fx, a.fy, a.f, a.alpha, a.baseline = c.fx, c.fy, c.f, c.alpha, c.baseline
    return arr


class Harness:
    """One of oracle/_ref/libhx_ref.so (stock reference), libhx_ref64.so (pin P3) or
    libhx_dropin.so (gipuma_b200 behind the unchanged runcuda boundary)."""

    def __init__(self, which: str = "ref"):
        name = {"ref": "libhx_ref.so", "ref64": "libhx_ref64.so", "dropin": "libhx_dropin.so"}[which]
        path = os.path.join(REF_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with oracle/build_ref.sh / __graft_entry__.build())")
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        L = self.lib
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.hx_backend.restype = C.c_char_p
        L.hx_max_views.restype = C.c_int
        L.hx_run.restype = C.c_int
        L.hx_run.argtypes = [C.POINTER(HxParams), C.c_int, C.c_int, C.c_int, fp, C.POINTER(HxCamera), C.c_int, ip,
                             C.c_ulonglong, fp, fp, C.POINTER(C.c_double)]
        L.hx_steps.restype = C.c_int
        L.hx_steps.argtypes = [C.POINTER(HxParams), C.c_int, C.c_int, C.c_int, fp, C.POINTER(HxCamera), C.c_int, ip,
                               C.c_ulonglong, ip, C.c_int, fp, fp, fp, fp, fp]
        L.hx_cost_eval.restype = C.c_int
        L.hx_cost_eval.argtypes = [C.POINTER(HxParams), C.c_int, C.c_int, C.c_int, fp, C.POINTER(HxCamera), C.c_int,
                                   ip, fp, fp]
        self.backend = L.hx_backend().decode()
        self.max_views = L.hx_max_views()

    @staticmethod
    def _fp(a):
        return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None

    def _common(self, scene):
        prm = pack_params(scene.params)
        cams = pack_cameras(scene.cameras)
        imgs = np.ascontiguousarray(scene.images, dtype=np.float32)
        sub = np.asarray(scene.subset, dtype=np.int32)
        return prm, cams, imgs, sub

    def run(self, scene, seed: int = 0xC0FFEE):
        """Full runcuda(): returns (norm4 [rows,cols,4], cost [rows,cols], printed_seconds, wall_ms)."""
        prm, cams, imgs, sub = self._common(scene)
        n4 = np.zeros((scene.rows, scene.cols, 4), dtype=np.float32)
        c = np.zeros((scene.rows, scene.cols), dtype=np.float32)
        times = (C.c_double * 2)()
        rc = self.lib.hx_run(C.byref(prm), scene.rows, scene.cols, len(scene.cameras), self._fp(imgs), cams,
                             len(sub), sub.ctypes.data_as(C.POINTER(C.c_int)), seed, self._fp(n4), self._fp(c), times)
        if rc != 0:
            raise RuntimeError("hx_run failed: %d" % rc)
        return n4, c, times[0], times[1]

    def steps(self, scene, steps: Sequence[int], norm4: Optional[np.ndarray] = None,
              cost: Optional[np.ndarray] = None, seed: int = 0xC0FFEE):
        """Launch the listed reference kernels from a given raw state; returns (norm4, cost, ms[])."""
        prm, cams, imgs, sub = self._common(scene)
        st = np.asarray(list(steps), dtype=np.int32)
        n4 = np.zeros((scene.rows, scene.cols, 4), dtype=np.float32)
        c = np.zeros((scene.rows, scene.cols), dtype=np.float32)
        ms = np.zeros(len(st), dtype=np.float32)
        i4 = None if norm4 is None else np.ascontiguousarray(norm4, dtype=np.float32)
        ic = None if cost is None else np.ascontiguousarray(cost, dtype=np.float32)
        rc = self.lib.hx_steps(C.byref(prm), scene.rows, scene.cols, len(scene.cameras), self._fp(imgs), cams,
                               len(sub), sub.ctypes.data_as(C.POINTER(C.c_int)), seed,
                               st.ctypes.data_as(C.POINTER(C.c_int)), len(st), self._fp(i4), self._fp(ic),
                               self._fp(n4), self._fp(c), self._fp(ms))
        if rc != 0:
            raise RuntimeError("hx_steps failed: %d" % rc)
        return n4, c, ms

    def run_fused(self, scene, seed: int = 0xC0FFEE):
        """runcuda() as the reference behaves when built without SMALLKERNEL (gipuma.cu:1906-1945): init, then per
        iteration the fused 20-neighbour black and red kernels, then gipuma_compute_disp.  Returns (norm4, cost,
        ms of the sweep span = everything after init)."""
        seq = [STEP_INIT] + [STEP_BLACK_FUSED, STEP_RED_FUSED] * scene.params.iterations + [STEP_COMPUTE_DISP]
        n4, c, ms = self.steps(scene, seq, seed=seed)
        return n4, c, float(ms[1:].sum())

    def cost_eval(self, scene, planes: np.ndarray):
        prm, cams, imgs, sub = self._common(scene)
        pl = np.ascontiguousarray(planes, dtype=np.float32)
        c = np.zeros((scene.rows, scene.cols), dtype=np.float32)
        rc = self.lib.hx_cost_eval(C.byref(prm), scene.rows, scene.cols, len(scene.cameras), self._fp(imgs), cams,
                                   len(sub), sub.ctypes.data_as(C.POINTER(C.c_int)), self._fp(pl), self._fp(c))
        if rc != 0:
            raise RuntimeError("hx_cost_eval failed: %d" % rc)
        return c
