"""ctypes binding of oracle/libgipuma_oracle.so — the single-thread C restatement (gipuma_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gipuma_b200.api import GpmParams, GpmCamera, pack_params, pack_camera

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libgipuma_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gipuma_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        base = ["gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-o", LIB, src, "-lm"]
        # OpenMP row loops where the toolchain has libgomp (bench.py's multi-core CPU baseline); scalar otherwise
        if subprocess.call(base[:2] + ["-fopenmp"] + base[2:], stderr=subprocess.DEVNULL) != 0:
            subprocess.check_call(base)
    return LIB


class Oracle:
    def __init__(self, scene):
        build()
        self.lib = C.CDLL(LIB)
        self.sc = scene
        self.prm = pack_params(scene.params)
        self.ref = pack_camera(scene.cameras[0])
        self.views = (GpmCamera * len(scene.subset))(*[pack_camera(scene.cameras[i]) for i in scene.subset])
        self.ref_img = np.ascontiguousarray(scene.images[0], dtype=np.float32)
        self.view_imgs = [np.ascontiguousarray(scene.images[i], dtype=np.float32) for i in scene.subset]
        self.vptrs = (C.POINTER(C.c_float) * len(self.view_imgs))(
            *[a.ctypes.data_as(C.POINTER(C.c_float)) for a in self.view_imgs])
        self.W, self.H, self.V = scene.cols, scene.rows, len(scene.subset)
        fp = C.POINTER(C.c_float)
        common = [C.c_int, C.c_int, C.c_int, C.POINTER(GpmParams), C.POINTER(GpmCamera), C.POINTER(GpmCamera), fp,
                  C.POINTER(fp)]
        self.lib.gpo_cost_eval.argtypes = common + [fp, fp, C.c_int, C.c_int, C.c_int]
        self.lib.gpo_phase.argtypes = common + [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int]
        self.lib.gpo_sweep.argtypes = common + [fp, fp, C.c_int, C.c_int, C.c_int]
        self.lib.gpo_finalize.argtypes = [C.c_int, C.c_int, C.POINTER(GpmCamera), fp, fp]
        self.lib.gpo_tex2d.argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_float]
        self.lib.gpo_tex2d.restype = C.c_float
        self.lib.gpo_random_plane.argtypes = [C.POINTER(GpmParams), C.POINTER(GpmCamera), C.c_int, C.c_int,
                                              C.POINTER(C.c_uint32), fp]

    def set_threads(self, n: int) -> int:
        """OpenMP threads for the row loops (results are independent of it); returns the number in effect."""
        n = max(1, min(int(n), int(self.lib.gpo_max_threads())))
        self.lib.gpo_set_threads(n)
        return n

    def _common(self):
        self.lib.gpo_set_color(int(self.ref_img.ndim == 3))          # [H, W, 4] images: the float4 path
        return (self.W, self.H, self.V, C.byref(self.prm), C.byref(self.ref), self.views,
                self.ref_img.ctypes.data_as(C.POINTER(C.c_float)), self.vptrs)

    @staticmethod
    def _fp(a):
        return a.ctypes.data_as(C.POINTER(C.c_float))

    def cost_eval(self, planes, y0=0, y1=None, init_radius=False):
        pl = np.ascontiguousarray(planes, dtype=np.float32)
        out = np.zeros((self.H, self.W), dtype=np.float32)
        y1 = self.H if y1 is None else y1
        self.lib.gpo_cost_eval(*self._common(), self._fp(pl), self._fp(out), y0, y1, int(init_radius))
        return out

    def phase(self, planes, cost, colour, phase_mask, y0=0, y1=None):
        pl = np.array(planes, dtype=np.float32, order="C", copy=True)
        c = np.array(cost, dtype=np.float32, order="C", copy=True)
        y1 = self.H if y1 is None else y1
        self.lib.gpo_phase(*self._common(), self._fp(pl), self._fp(c), colour, phase_mask, y0, y1)
        return pl, c

    def sweep(self, planes, cost, iterations, y0=0, y1=None):
        pl = np.array(planes, dtype=np.float32, order="C", copy=True)
        c = np.array(cost, dtype=np.float32, order="C", copy=True)
        y1 = self.H if y1 is None else y1
        self.lib.gpo_sweep(*self._common(), self._fp(pl), self._fp(c), iterations, y0, y1)
        return pl, c

    def finalize(self, planes, cost):
        pl = np.array(planes, dtype=np.float32, order="C", copy=True)
        c = np.ascontiguousarray(cost, dtype=np.float32)
        self.lib.gpo_finalize(self.W, self.H, C.byref(self.ref), self._fp(pl), self._fp(c))
        return pl

    def tex2d(self, img, x, y):
        a = np.ascontiguousarray(img, dtype=np.float32)
        return float(self.lib.gpo_tex2d(self._fp(a), a.shape[1], a.shape[0], x, y))

    def random_plane(self, px, py, state6):
        st = (C.c_uint32 * 6)(*[int(v) for v in state6])
        out = np.zeros(4, dtype=np.float32)
        self.lib.gpo_random_plane(C.byref(self.prm), C.byref(self.ref), px, py, st, self._fp(out))
        return out

    def plane_depth(self, plane, px, py) -> float:
        """getDisparity_cu / getDepthFromPlane3_cu (gipuma.cu:694-715): depth of the plane (n, d) along the ray of pixel (px, py)."""
        p = np.ascontiguousarray(plane, dtype=np.float32)
        self.lib.gpo_plane_depth.restype = C.c_float
        self.lib.gpo_plane_depth.argtypes = [C.POINTER(GpmCamera), C.POINTER(C.c_float), C.c_int, C.c_int]
        return float(self.lib.gpo_plane_depth(C.byref(self.ref), self._fp(p), px, py))

    def plane_d(self, normal, px, py, depth) -> float:
        """getD_cu (gipuma.cu:96-111): d of the plane with this normal through the point at `depth` on the pixel's ray."""
        n = np.ascontiguousarray(normal[:3], dtype=np.float32)
        self.lib.gpo_plane_d.restype = C.c_float
        self.lib.gpo_plane_d.argtypes = [C.POINTER(GpmCamera), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float]
        return float(self.lib.gpo_plane_d(C.byref(self.ref), self._fp(n), px, py, depth))

    def curand_init(self, seed: int, subsequence: int, offset: int) -> np.ndarray:
        """XORWOW state (v[0..4], d) after curand_init(seed, subsequence, offset) — restated in gipuma_oracle.c."""
        st = (C.c_uint32 * 6)()
        self.lib.gpo_curand_init.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
        self.lib.gpo_curand_init.restype = None
        self.lib.gpo_curand_init(seed, subsequence, offset, st)
        return np.array(list(st), dtype=np.uint32)

    def init_planes(self, seed: int) -> np.ndarray:
        """The random planes of gipuma_init_cu2 (gipuma.cu:1019-1034) for the whole image: curand_init(seed, y, x) per pixel."""
        out = np.zeros((self.H, self.W, 4), dtype=np.float32)
        self.lib.gpo_init_planes.argtypes = [C.c_int, C.c_int, C.POINTER(GpmParams), C.POINTER(GpmCamera), C.c_uint64,
                                             C.POINTER(C.c_float)]
        self.lib.gpo_init_planes(self.W, self.H, C.byref(self.prm), C.byref(self.ref), seed, self._fp(out))
        return out

    def run(self, seed: int, fused: bool = False):
        """runcuda() on the CPU (gipuma.cu:1825-1960): random planes, initial costs, `iterations` red/black sweeps
        (fused = the 20-neighbour kernels of a build without SMALLKERNEL), gipuma_compute_disp.  Returns (norm4, cost)."""
        pl = self.init_planes(seed)
        c = self.cost_eval(pl, init_radius=True)
        for _ in range(self.sc.params.iterations):
            for colour in (0, 1):
                for mask in ((8 | 4,) if fused else (1, 2, 4)):
                    pl, c = self.phase(pl, c, colour, mask)
        return self.finalize(pl, c), c
