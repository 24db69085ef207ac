"""(De)serialisation of a Scene to plain arrays, used by the golden fixtures under tests/golden/."""
from __future__ import annotations

import numpy as np

from .scene import AlgorithmParameters, Camera, Scene

_CAM_MATS = ("K", "K_inv", "R", "R_orig_inv", "M_inv", "P", "t", "C")
_CAM_SCALARS = ("fx", "fy", "f", "alpha", "baseline")
_PARAM_FIELDS = ("max_disparity", "min_disparity", "box_hsize", "box_vsize", "tau_color", "tau_gradient", "alpha",
                 "gamma", "iterations", "good_factor", "n_best", "cost_comb", "depthMin", "depthMax")


def scene_to_arrays(sc: Scene) -> dict:
    out = {"images_u8": np.asarray(sc.images, dtype=np.float32).astype(np.uint8), "subset": np.asarray(sc.subset, np.int32),
           "rows": np.int32(sc.rows), "cols": np.int32(sc.cols)}
    assert np.array_equal(out["images_u8"].astype(np.float32), sc.images), "scene images must be 8-bit valued"
    for m in _CAM_MATS:
        out["cam_" + m] = np.stack([np.asarray(getattr(c, m), dtype=np.float32) for c in sc.cameras])
    out["cam_scalars"] = np.array([[getattr(c, s) for s in _CAM_SCALARS] for c in sc.cameras], dtype=np.float32)
    out["params"] = np.array([float(getattr(sc.params, f)) for f in _PARAM_FIELDS], dtype=np.float64)
    return out


def scene_from_arrays(name: str, z) -> Scene:
    p = AlgorithmParameters()
    for f, v in zip(_PARAM_FIELDS, z["params"]):
        cur = getattr(p, f)
        setattr(p, f, int(round(v)) if isinstance(cur, int) and not isinstance(cur, bool) else float(np.float32(v)))
    cams = []
    for i in range(z["cam_K"].shape[0]):
        kw = {m: np.ascontiguousarray(z["cam_" + m][i], dtype=np.float32) for m in _CAM_MATS}
        sc = {s: float(z["cam_scalars"][i, k]) for k, s in enumerate(_CAM_SCALARS)}
        cams.append(Camera(**kw, **sc))
    p.color_processing = z["images_u8"].ndim == 4          # [n, rows, cols, 4]: the float4 path
    return Scene(name=name, rows=int(z["rows"]), cols=int(z["cols"]), images=z["images_u8"].astype(np.float32),
                 cameras=cams, subset=[int(v) for v in z["subset"]], params=p)
