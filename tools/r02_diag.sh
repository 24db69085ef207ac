#!/bin/bash
out=gpurun_out; mkdir -p $out
for o in "tma=0" "tma=1"; do
for b in 11 15 9 25; do
  echo "== box $b $o"; timeout 120 python - <<PY 2>&1 | tail -3
import sys; sys.path.insert(0,'.')
from gipuma_b200 import api, scene as S
sc = S.make_config(2, rows=64, cols=96, n_views=4, iterations=2)
sc.params.box_hsize = sc.params.box_vsize = $b
try:
    ls, ms, st = api.runcuda(sc, options={"${o%=*}": ${o#*=}})
    print("ok", ms, float(ls.c.mean()))
except Exception as e:
    print("ERR", e)
PY
done; done
timeout 300 python -m pytest tests/test_gpu_fused_sweep.py -x -q 2>&1 | tail -30
