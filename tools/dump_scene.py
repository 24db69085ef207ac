#!/usr/bin/env python
"""Write a gipuma_b200.scene.Scene as the flat file examples/shard_host.cpp reads (C structs of include/gipuma_b200.h):
int32 W, H, n_images, n_views; uint64 seed; gpm_params; int32 subset[n_views]; gpm_camera[n_images]; float32 images."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gipuma_b200 import api


def dump(scene, path: str, seed: int = 0xC0FFEE):
    imgs = np.ascontiguousarray(scene.images, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(np.array([scene.cols, scene.rows, imgs.shape[0], len(scene.subset)], np.int32).tobytes())
        f.write(np.array([seed], np.uint64).tobytes())
        f.write(bytes(api.pack_params(scene.params)))
        f.write(np.array(scene.subset, np.int32).tobytes())
        for cam in scene.cameras:
            f.write(bytes(api.pack_camera(cam)))
        f.write(imgs.tobytes())


if __name__ == "__main__":
    from gipuma_b200 import scene as S
    cfg, out = int(sys.argv[1]), sys.argv[2]
    dump(S.make_config(cfg, workers=16), out)
    print(out, os.path.getsize(out), "bytes")
