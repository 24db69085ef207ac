#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
for b in 11 9 25 17; do
  echo "== box $b tma=1"; timeout 120 python - <<PY 2>&1 | tail -2
import sys; sys.path.insert(0,'.')
from gipuma_b200 import api, scene as S
sc = S.make_config(2, rows=64, cols=96, n_views=4, iterations=2)
sc.params.box_hsize = sc.params.box_vsize = $b
a, ms, st = api.runcuda(sc, options={"tma": 1}); b_, _, _ = api.runcuda(sc, options={"tma": 0})
import numpy as np
print("ok", ms, np.array_equal(a.norm4.view(np.uint32), b_.norm4.view(np.uint32)))
PY
done
timeout 1800 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -16 > $out/pytest_gpu_exp3.log; cat $out/pytest_gpu_exp3.log
timeout 600 python tools/run_shard_nccl.py --config 4 2>&1 | tail -1 | tee $out/exp3_shard_cfg4_1gpu.json | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 2>$out/bench.err | tail -1 > $out/bench_ours_1gpu.json; cut -c1-400 $out/bench_ours_1gpu.json
