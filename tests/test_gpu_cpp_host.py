"""north_star: "Host stays C++ calling a thin C-ABI".  examples/shard_host.cpp is a C++ program (g++, no CUDA headers, no
Python) that runs one reference view with its source views sharded over the visible GPUs — one thread per GPU, NCCL id handed
over in memory, gpm_shard_run — and writes rank 0's LineState arrays.  Its output must equal the Python-driven single-GPU run
bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_runs_the_view_shard_through_the_cabi(tmp_path):
    import torch
    from gipuma_b200 import api, build, scene as S
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import dump_scene
    exe = build.build_examples()
    sc = S.make_config(4, rows=96, cols=128, n_views=11, iterations=2, seed=321)
    scene_file, out_file = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    dump_scene.dump(sc, scene_file, seed=0xC0FFEE)
    world = max(1, min(2, torch.cuda.device_count()))
    p = subprocess.run([exe, scene_file, out_file, str(world)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    info = json.loads(p.stdout.strip().splitlines()[-1])
    assert info["world"] == world and info["views"] == 11 and info["sweep_ms_max_over_ranks"] > 0
    raw = np.fromfile(out_file, dtype=np.float32)
    n = sc.rows * sc.cols
    n4, c = raw[: 4 * n].reshape(sc.rows, sc.cols, 4), raw[4 * n:].reshape(sc.rows, sc.cols)
    single, _, _ = api.runcuda(sc, seed=0xC0FFEE)
    assert bits_equal(n4, single.norm4) == 0 and bits_equal(c, single.c) == 0
