"""Host-side mirror of the reference's operator boundary for the PatchMatch hot path.

`runcuda(gs)` here takes the same information the reference's `int runcuda(GlobalState &gs)`
(gipuma.h:2) takes — AlgorithmParameters, the per-view Camera_cu fields, the view selection subset and
the images — and produces the same outputs (`lines.norm4`: world normal + depth, `lines.c`: cost).  All
compute goes through the C-ABI of `libgipuma_b200.so` (include/gipuma_b200.h, hand-written sm_100a
CUDA); there is no CPU path: loading fails loudly when the library has not been built, and every call
fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GIPUMA_B200_LIB", os.path.join(_HERE, "libgipuma_b200.so"))

GPM_RNG_REFERENCE, GPM_RNG_STATEFUL = 0, 1


class GpmParams(C.Structure):
    _fields_ = [("box_hsize", C.c_int), ("box_vsize", C.c_int),
                ("tau_color", C.c_float), ("tau_gradient", C.c_float),
                ("alpha", C.c_float), ("gamma", C.c_float),
                ("min_disparity", C.c_float), ("max_disparity", C.c_float),
                ("iterations", C.c_int), ("n_best", C.c_int), ("cost_comb", C.c_int),
                ("good_factor", C.c_float), ("depthMin", C.c_float), ("depthMax", C.c_float)]


class GpmCamera(C.Structure):
    _fields_ = [("K", C.c_float * 9), ("K_inv", C.c_float * 9), ("R", C.c_float * 9),
                ("M_inv", C.c_float * 9), ("R_orig_inv", C.c_float * 9),
                ("t", C.c_float * 3), ("C", C.c_float * 3), ("P_col34", C.c_float * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("f", C.c_float), ("alpha", C.c_float),
                ("baseline", C.c_float)]


class GpmBatchDesc(C.Structure):
    _fields_ = [("n_images", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("images", C.POINTER(C.c_void_p)), ("pitch_bytes", C.c_size_t), ("P", C.POINTER(C.c_double)),
                ("cam_scale", C.c_double), ("params", GpmParams), ("min_angle", C.c_float), ("max_angle", C.c_float),
                ("max_views", C.c_int), ("ref_indices", C.POINTER(C.c_int)), ("n_refs", C.c_int),
                ("devices", C.POINTER(C.c_int)), ("n_devices", C.c_int), ("seed", C.c_ulonglong),
                ("out_dir", C.c_char_p), ("out_norm4", C.POINTER(C.c_float)), ("out_cost", C.POINTER(C.c_float))]


class GpmBatchStats(C.Structure):
    _fields_ = [("jobs_done", C.c_int), ("sweep_ms_total", C.c_double), ("per_job_sweep_ms", C.POINTER(C.c_float)),
                ("per_job_views", C.POINTER(C.c_int)), ("per_job_device", C.POINTER(C.c_int))]


class GipumaError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libgipuma_b200.so and declare the C-ABI.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GipumaError("%s is missing: build it with `python -m gipuma_b200.build` "
                          "(no CPU fallback exists)" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, fp = C.c_void_p, C.POINTER(C.c_float)
    L.gpm_last_error.restype = C.c_char_p
    L.gpm_version.restype = C.c_char_p
    L.gpm_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int]
    L.gpm_destroy.argtypes = [vp]
    L.gpm_destroy.restype = None
    L.gpm_set_params.argtypes = [vp, C.POINTER(GpmParams)]
    L.gpm_set_reference.argtypes = [vp, vp, C.c_size_t, C.c_int, C.POINTER(GpmCamera)]
    L.gpm_set_view.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, C.POINTER(GpmCamera)]
    L.gpm_set_reference_color.argtypes = L.gpm_set_reference.argtypes
    L.gpm_set_view_color.argtypes = L.gpm_set_view.argtypes
    L.gpm_set_num_views.argtypes = [vp, C.c_int]
    L.gpm_set_rng.argtypes = [vp, C.c_ulonglong, C.c_int]
    L.gpm_set_state.argtypes = [vp, vp, vp, C.c_int]
    L.gpm_get_state.argtypes = [vp, vp, vp, C.c_int]
    L.gpm_init.argtypes = [vp]
    L.gpm_sweep.argtypes = [vp, C.c_int]
    L.gpm_phase.argtypes = [vp, C.c_int, C.c_int]
    L.gpm_finalize.argtypes = [vp]
    L.gpm_cost_eval.argtypes = [vp, vp, vp, C.c_int]
    L.gpm_run.argtypes = [vp, fp]
    L.gpm_get_stats.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.gpm_reset_stats.argtypes = [vp]
    L.gpm_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    L.gpm_stream.argtypes = [vp]
    L.gpm_prepare_cameras.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.POINTER(GpmCamera)]
    L.gpm_select_views.argtypes = [C.POINTER(GpmCamera), C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                   C.POINTER(C.c_int), fp]
    L.gpm_write_dmb.argtypes = [C.c_char_p, fp, C.c_int, C.c_int, C.c_int]
    L.gpm_read_dmb.argtypes = [C.c_char_p, fp, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.gpm_write_result_dmb.argtypes = [C.c_char_p, C.c_char_p, fp, C.c_int, C.c_int]
    L.gpm_init_planes.argtypes = [vp]
    L.gpm_shard_num_stages.argtypes = [vp]
    L.gpm_shard_stage_floats.argtypes = [vp, C.c_int]
    L.gpm_shard_stage_floats.restype = C.c_longlong
    L.gpm_shard_stage.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp]
    L.gpm_shard_finish_init.argtypes = [vp, vp, C.c_int]
    L.gpm_shard_unique_id.argtypes = [vp]
    L.gpm_shard_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.gpm_shard_comm_attach.argtypes = [vp, vp, C.c_int, C.c_int]
    L.gpm_shard_run.argtypes = [vp, fp]
    L.gpm_measure_fetch_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.gpm_debug_packed_mismatches.argtypes = [vp, C.POINTER(C.c_uint), fp, C.c_int]
    L.gpm_batch_run.argtypes = [C.POINTER(GpmBatchDesc), C.POINTER(GpmBatchStats)]
    L.gpm_batch_last_error.restype = C.c_char_p
    L.gpm_shard_p2p_export.argtypes = [vp, C.c_int, vp, C.POINTER(vp)]
    L.gpm_shard_p2p_attach.argtypes = [vp, vp, C.POINTER(vp), C.c_int, C.c_int]
    L.gpm_stream.restype = vp
    for name in ("gpm_create", "gpm_set_params", "gpm_set_reference", "gpm_set_view", "gpm_set_reference_color",
                 "gpm_set_view_color", "gpm_set_num_views",
                 "gpm_set_rng", "gpm_set_state", "gpm_get_state", "gpm_init", "gpm_sweep", "gpm_phase",
                 "gpm_finalize", "gpm_cost_eval", "gpm_run", "gpm_get_stats", "gpm_reset_stats", "gpm_set_option",
                 "gpm_init_planes", "gpm_shard_num_stages", "gpm_shard_stage", "gpm_shard_finish_init", "gpm_shard_unique_id",
                 "gpm_shard_comm_init", "gpm_shard_comm_attach", "gpm_shard_run", "gpm_measure_fetch_peak", "gpm_shard_p2p_export", "gpm_shard_p2p_attach", "gpm_batch_run", "gpm_debug_packed_mismatches",
                 "gpm_prepare_cameras", "gpm_select_views", "gpm_write_dmb", "gpm_read_dmb", "gpm_write_result_dmb"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def pack_params(p) -> GpmParams:
    return GpmParams(p.box_hsize, p.box_vsize, p.tau_color, p.tau_gradient, p.alpha, p.gamma,
                     p.min_disparity, p.max_disparity, p.iterations, p.n_best, p.cost_comb,
                     p.good_factor, p.depthMin, p.depthMax)


def pack_camera(c) -> GpmCamera:
    g = GpmCamera()
    g.K[:] = c.K.ravel().tolist()
    g.K_inv[:] = c.K_inv.ravel().tolist()
    g.R[:] = c.R.ravel().tolist()
    g.M_inv[:] = c.M_inv.ravel().tolist()
    g.R_orig_inv[:] = c.R_orig_inv.ravel().tolist()
    g.t[:] = c.t.ravel().tolist()
    g.C[:] = c.C.ravel().tolist()
    g.P_col34[:] = c.P[:, 3].ravel().tolist()
    g.fx, g.fy, g.f, g.alpha, g.baseline = c.fx, c.fy, c.f, c.alpha, c.baseline
    return g


def _ptr(a):
    """void* of a numpy array, a torch tensor (host or CUDA) or a raw integer address.  The C-ABI reads / writes float32
    row-major memory: anything else is refused here instead of being reinterpreted silently."""
    if a is None:
        return None, 0
    if isinstance(a, int):
        return C.c_void_p(a), 1
    if isinstance(a, np.ndarray):
        if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
            raise TypeError("numpy buffers must be C-contiguous float32 (got %s, contiguous=%s)" % (a.dtype, a.flags["C_CONTIGUOUS"]))
        return C.c_void_p(a.ctypes.data), 0
    if hasattr(a, "data_ptr"):
        import torch
        if a.dtype != torch.float32 or not a.is_contiguous():
            raise TypeError("torch buffers must be contiguous float32 (got %s)" % (a.dtype,))
        if a.is_cuda:
            torch.cuda.current_stream(a.device).synchronize()      # the context works on its own stream: the producer must be done
        return C.c_void_p(a.data_ptr()), 1 if a.is_cuda else 0
    raise TypeError("unsupported buffer type %r" % type(a))


def _image(a):
    """Images may come as any real numpy dtype: converted to contiguous float32 here (what main.cpp's loaders produce)."""
    if isinstance(a, np.ndarray) and (a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]):
        return np.ascontiguousarray(a, dtype=np.float32)
    return a


class Context:
    """One gpm_ctx: a reference view with its source views on one GPU."""

    def __init__(self, width: int, height: int, max_views: int, device: int = 0):
        self.lib = load_library()
        self.W, self.H, self.max_views, self.device = width, height, max_views, device
        h = C.c_void_p()
        self._check(self.lib.gpm_create(C.byref(h), device, width, height, max_views))
        self.h = h

    def _check(self, rc: int):
        if rc != 0:
            raise GipumaError("gipuma_b200 error %d: %s" % (rc, self.lib.gpm_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- inputs ---------------------------------------------------------------------------------
    def set_params(self, params):
        p = pack_params(params)
        self._check(self.lib.gpm_set_params(self.h, C.byref(p)))

    @staticmethod
    def _is_color(img) -> bool:
        """[H, W, 4] float images take the reference's -color_processing (float4) path; [H, W] the gray one."""
        shape = getattr(img, "shape", ())
        return len(shape) == 3 and shape[-1] == 4

    def set_reference(self, img, cam, pitch_bytes: int = 0):
        img = _image(img)
        ptr, dev = _ptr(img)
        g = pack_camera(cam)
        fn = self.lib.gpm_set_reference_color if self._is_color(img) else self.lib.gpm_set_reference
        self._check(fn(self.h, ptr, pitch_bytes, dev, C.byref(g)))

    def set_view(self, v: int, img, cam, pitch_bytes: int = 0):
        img = _image(img)
        ptr, dev = _ptr(img)
        g = pack_camera(cam)
        fn = self.lib.gpm_set_view_color if self._is_color(img) else self.lib.gpm_set_view
        self._check(fn(self.h, v, ptr, pitch_bytes, dev, C.byref(g)))

    def set_num_views(self, n: int):
        self._check(self.lib.gpm_set_num_views(self.h, n))

    def set_rng(self, seed: int, mode: int = GPM_RNG_REFERENCE):
        self._check(self.lib.gpm_set_rng(self.h, seed, mode))

    def set_option(self, name: str, value: int):
        self._check(self.lib.gpm_set_option(self.h, name.encode(), int(value)))

    def load_scene(self, scene, seed: int = 0xC0FFEE, rng_mode: int = GPM_RNG_REFERENCE, images=None):
        """Upload a gipuma_b200.scene.Scene (images may be overridden by a list of device tensors)."""
        imgs = scene.images if images is None else images
        self.set_params(scene.params)
        self.set_reference(np.ascontiguousarray(imgs[0]) if isinstance(imgs, np.ndarray) else imgs[0], scene.cameras[0])
        for v, idx in enumerate(scene.subset):
            im = np.ascontiguousarray(imgs[idx]) if isinstance(imgs, np.ndarray) else imgs[idx]
            self.set_view(v, im, scene.cameras[idx])
        self.set_num_views(len(scene.subset))
        self.set_rng(seed, rng_mode)

    # -- state ----------------------------------------------------------------------------------
    def set_state(self, norm4=None, cost=None):
        """Overwrite the raw planes [H,W,4] and/or costs [H,W] (numpy or torch, host or device)."""
        if isinstance(norm4, np.ndarray):
            norm4 = np.ascontiguousarray(norm4, dtype=np.float32)
        if isinstance(cost, np.ndarray):
            cost = np.ascontiguousarray(cost, dtype=np.float32)
        p4, d4 = _ptr(norm4)
        pc, dc = _ptr(cost)
        if norm4 is not None and cost is not None and d4 != dc:
            raise ValueError("norm4 and cost must both be host or both be device buffers")
        self._check(self.lib.gpm_set_state(self.h, p4, pc, max(d4, dc)))

    def get_state(self):
        n4 = np.empty((self.H, self.W, 4), dtype=np.float32)
        c = np.empty((self.H, self.W), dtype=np.float32)
        self._check(self.lib.gpm_get_state(self.h, C.c_void_p(n4.ctypes.data), C.c_void_p(c.ctypes.data), 0))
        return n4, c

    def get_state_into(self, norm4, cost):
        """Copy the state into caller buffers (numpy or torch, host or device)."""
        p4, d4 = _ptr(norm4)
        pc, dc = _ptr(cost)
        self._check(self.lib.gpm_get_state(self.h, p4, pc, max(d4, dc)))

    # -- compute --------------------------------------------------------------------------------
    def init(self):
        self._check(self.lib.gpm_init(self.h))

    def sweep(self, iterations: int):
        self._check(self.lib.gpm_sweep(self.h, iterations))

    def phase(self, colour: int, phase_mask: int):
        self._check(self.lib.gpm_phase(self.h, colour, phase_mask))

    def finalize(self):
        self._check(self.lib.gpm_finalize(self.h))

    def cost_eval(self, planes: np.ndarray) -> np.ndarray:
        pl = np.ascontiguousarray(planes, dtype=np.float32)
        out = np.empty((self.H, self.W), dtype=np.float32)
        self._check(self.lib.gpm_cost_eval(self.h, C.c_void_p(pl.ctypes.data), C.c_void_p(out.ctypes.data), 0))
        return out

    def run(self) -> float:
        """init + params.iterations sweeps + finalize; returns the sweep time in ms (reference's timed span)."""
        ms = C.c_float(0)
        self._check(self.lib.gpm_run(self.h, C.byref(ms)))
        return float(ms.value)

    # -- source-view sharding (multi-GPU) -----------------------------------------------------------
    def init_planes(self):
        self._check(self.lib.gpm_init_planes(self.h))

    def shard_num_stages(self) -> int:
        n = self.lib.gpm_shard_num_stages(self.h)
        if n < 0:
            self._check(n)
        return n

    def shard_stage_floats(self, stage: int) -> int:
        n = self.lib.gpm_shard_stage_floats(self.h, stage)
        if n < 0:
            self._check(int(n))
        return int(n)

    def shard_stage(self, colour: int, stage: int, gathered_prev, world: int, xchg):
        """Accept of the previous stage (from `gathered_prev`, device tensor or None) + evaluation of `stage` into `xchg`."""
        gp = C.c_void_p(gathered_prev.data_ptr()) if gathered_prev is not None else None
        xp = C.c_void_p(xchg.data_ptr()) if xchg is not None else None
        self._check(self.lib.gpm_shard_stage(self.h, colour, stage, gp, world, xp))

    def shard_finish_init(self, gathered, world: int):
        self._check(self.lib.gpm_shard_finish_init(self.h, C.c_void_p(gathered.data_ptr()), world))

    def shard_comm_init(self, unique_id: Optional[bytes], rank: int, world: int):
        """ncclCommInitRank behind the C-ABI; `unique_id` = the 128 bytes of shard_unique_id() of the group's rank 0."""
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._check(self.lib.gpm_shard_comm_init(self.h, buf, rank, world))

    def shard_p2p_export(self, world: int):
        """Allocate this rank's peer-memory exchange region; returns (64-byte CUDA IPC handle, device address)."""
        h = C.create_string_buffer(64)
        ptr = C.c_void_p()
        self._check(self.lib.gpm_shard_p2p_export(self.h, world, h, C.byref(ptr)))
        return h.raw, int(ptr.value or 0)

    def shard_p2p_attach(self, handles, rank: int, world: int, local_ptrs=None):
        """handles: list of `world` 64-byte IPC handles (rank-major); local_ptrs: optional addresses of same-process ranks."""
        blob = C.create_string_buffer(b"".join(handles), 64 * world)
        lp = None
        if local_ptrs is not None:
            lp = (C.c_void_p * world)(*[C.c_void_p(p) if p else None for p in local_ptrs])
        self._check(self.lib.gpm_shard_p2p_attach(self.h, blob, lp, rank, world))

    def shard_run(self) -> float:
        """runcuda() with sharded views (gpm_shard_run): returns the sweep time in ms."""
        ms = C.c_float(0)
        self._check(self.lib.gpm_shard_run(self.h, C.byref(ms)))
        return float(ms.value)

    def stats(self) -> dict:
        s = (C.c_ulonglong * 8)()
        self._check(self.lib.gpm_get_stats(self.h, s))
        return {"launches": s[0], "hypotheses": s[1], "skipped": s[2], "pruned": s[3],
                "pairs": s[4], "pairs_full": s[5], "collectives": s[6]}

    def packed_mismatches(self, reset: bool = True):
        n = C.c_uint(0)
        rec = np.zeros((64, 8), np.float32)
        self._check(self.lib.gpm_debug_packed_mismatches(self.h, C.byref(n), rec.ctypes.data_as(C.POINTER(C.c_float)), int(reset)))
        return int(n.value), rec[: min(64, int(n.value))]

    def measure_fetch_peak(self) -> float:
        """Texture-unit ceiling on this GPU, in 1e9 filtered R32F fetches per second (gpm_measure_fetch_peak)."""
        v = C.c_double(0)
        self._check(self.lib.gpm_measure_fetch_peak(self.h, C.byref(v)))
        return float(v.value)

    def reset_stats(self):
        self._check(self.lib.gpm_reset_stats(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.gpm_stream(self.h) or 0)


def shard_unique_id() -> bytes:
    """128-byte NCCL unique id for gpm_shard_comm_init (call on the shard group's rank 0, broadcast to the others)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.gpm_shard_unique_id(buf)
    if rc != 0:
        raise GipumaError("gpm_shard_unique_id: %s" % lib.gpm_last_error().decode())
    return buf.raw


class LineState:
    """Outputs in the reference's LineState layout (linestate.h:8-24)."""

    def __init__(self, norm4: np.ndarray, c: np.ndarray):
        self.norm4 = norm4          # [rows, cols, 4]: world normal xyz, depth (0 where cost == MAXCOST)
        self.c = c                  # [rows, cols]


def runcuda(scene, seed: int = 0xC0FFEE, rng_mode: int = GPM_RNG_REFERENCE, device: int = 0,
            options: Optional[dict] = None):
    """The reference's `runcuda(GlobalState&)` for a Scene: returns (LineState, sweep_ms, stats)."""
    with Context(scene.cols, scene.rows, max(1, len(scene.subset)), device) as ctx:
        for k, v in (options or {}).items():
            ctx.set_option(k, v)
        ctx.load_scene(scene, seed=seed, rng_mode=rng_mode)
        ms = ctx.run()
        n4, c = ctx.get_state()
        st = ctx.stats()
    return LineState(n4, c), ms, st


# ---- host-side rows of SURVEY.md §8f (no GPU needed) ---------------------------------------------------------------

def prepare_cameras(Ps, cam_scale: float = 1.0):
    """C++ restatement of getCameraParameters (cameraGeometryUtils.h:174-353): list of 3x4 projections -> GpmCamera[]."""
    lib = load_library()
    P = np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.float64) for p in Ps]).reshape(-1))
    out = (GpmCamera * len(Ps))()
    rc = lib.gpm_prepare_cameras(P.ctypes.data_as(C.POINTER(C.c_double)), len(Ps), float(cam_scale), out)
    if rc != 0:
        raise GipumaError("gpm_prepare_cameras failed: %d" % rc)
    return out


def select_views(cams, cols: int, rows: int, min_angle: float, max_angle: float, max_views: int):
    """Deterministic selectViews (main.cpp:430-499).  Returns (subset, (depth_min, depth_max))."""
    lib = load_library()
    sub = (C.c_int * max(1, len(cams)))()
    rng = (C.c_float * 2)()
    n = lib.gpm_select_views(cams, len(cams), cols, rows, min_angle, max_angle, max_views, sub, rng)
    if n < 0:
        raise GipumaError("gpm_select_views failed: %d" % n)
    return [sub[i] for i in range(n)], (rng[0], rng[1])


def batch_run(images, Ps, params, refs, devices=(0,), cam_scale: float = 1.0, min_angle: float = 10.0, max_angle: float = 30.0,
              max_views: int = 9, seed: int = 0xC0FFEE, out_dir: Optional[str] = None, want_outputs: bool = True):
    """Reference-view batch in ONE process (gpm_batch_run): `images` [n, H, W] float32 (shared, page-locked once), `Ps` n 3x4
    projections, `refs` the reference views to process, `devices` the CUDA ordinals to use (one worker thread each).
    Returns (norm4 [len(refs), H, W, 4] or None, cost or None, stats dict)."""
    lib = load_library()
    imgs = np.ascontiguousarray(images, dtype=np.float32)
    n, H, W = imgs.shape
    ptrs = (C.c_void_p * n)(*[C.c_void_p(imgs[i].ctypes.data) for i in range(n)])
    P = np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.float64) for p in Ps]).reshape(-1))
    refs_a = (C.c_int * len(refs))(*[int(r) for r in refs])
    devs = (C.c_int * len(devices))(*[int(d) for d in devices])
    n4 = np.empty((len(refs), H, W, 4), np.float32) if want_outputs else None
    cc = np.empty((len(refs), H, W), np.float32) if want_outputs else None
    d = GpmBatchDesc()
    d.n_images, d.width, d.height = n, W, H
    d.images, d.pitch_bytes = ptrs, 0
    d.P, d.cam_scale = P.ctypes.data_as(C.POINTER(C.c_double)), float(cam_scale)
    d.params = pack_params(params)
    d.min_angle, d.max_angle, d.max_views = float(min_angle), float(max_angle), int(max_views)
    d.ref_indices, d.n_refs, d.devices, d.n_devices, d.seed = refs_a, len(refs), devs, len(devices), seed
    d.out_dir = out_dir.encode() if out_dir else None
    d.out_norm4 = n4.ctypes.data_as(C.POINTER(C.c_float)) if want_outputs else None
    d.out_cost = cc.ctypes.data_as(C.POINTER(C.c_float)) if want_outputs else None
    ms = (C.c_float * len(refs))()
    nv = (C.c_int * len(refs))()
    dv = (C.c_int * len(refs))()
    st = GpmBatchStats(0, 0.0, ms, nv, dv)
    rc = lib.gpm_batch_run(C.byref(d), C.byref(st))
    if rc != 0:
        raise GipumaError("gpm_batch_run failed (%d): %s" % (rc, lib.gpm_batch_last_error().decode()))
    return n4, cc, {"jobs_done": st.jobs_done, "sweep_ms_total": st.sweep_ms_total, "sweep_ms": list(ms), "views": list(nv), "device": list(dv)}


def write_dmb(path: str, data: np.ndarray):
    a = np.ascontiguousarray(data, dtype=np.float32)
    ch = 1 if a.ndim == 2 else a.shape[2]
    rc = load_library().gpm_write_dmb(path.encode(), a.ctypes.data_as(C.POINTER(C.c_float)), a.shape[0], a.shape[1], ch)
    if rc != 0:
        raise GipumaError("gpm_write_dmb failed")


def read_dmb(path: str) -> np.ndarray:
    lib = load_library()
    r, c, ch = C.c_int(), C.c_int(), C.c_int()
    if lib.gpm_read_dmb(path.encode(), None, 0, C.byref(r), C.byref(c), C.byref(ch)) != 0:
        raise GipumaError("gpm_read_dmb: cannot read %s" % path)
    out = np.empty((r.value, c.value, ch.value), dtype=np.float32)
    if lib.gpm_read_dmb(path.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), out.size, C.byref(r), C.byref(c), C.byref(ch)) != 0:
        raise GipumaError("gpm_read_dmb: short file %s" % path)
    return out[..., 0] if ch.value == 1 else out


def write_result_dmb(depth_path: str, normal_path: str, norm4: np.ndarray):
    a = np.ascontiguousarray(norm4, dtype=np.float32)
    rc = load_library().gpm_write_result_dmb(depth_path.encode(), normal_path.encode(), a.ctypes.data_as(C.POINTER(C.c_float)),
                                             a.shape[0], a.shape[1])
    if rc != 0:
        raise GipumaError("gpm_write_result_dmb failed")
