#!/bin/bash
# round-2 experiment 1: texture-unit weight read-out, parity with the quad-friendly lane permutation, its effect on time
set -u
out=gpurun_out; mkdir -p $out
timeout 300 tools/texprobe > $out/texprobe.txt 2>&1; tail -5 $out/texprobe.txt
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > $out/pytest_gpu_exp1.log; tail -2 $out/pytest_gpu_exp1.log
for cfg in 2 3 4; do
  for qp in 0 1; do
    timeout 600 python tools/run_mine.py --config $cfg --repeat 3 --opt quadperm=$qp 2>&1 | tail -1 > $out/exp1_cfg${cfg}_qp${qp}.json
    python - <<PY
import json
d=json.load(open("$out/exp1_cfg${cfg}_qp${qp}.json"))
print("cfg$cfg quadperm=$qp", [round(r["sweep_ms"],1) for r in d["runs"]], d["mean_cost"], d["frac_within_1pct_of_gt"])
PY
  done
done
