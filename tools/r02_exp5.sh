#!/bin/bash
# 2-GPU run: source-view shard over real ranks — fused peer-memory exchange vs NCCL all-gather — tests, timings, bench
set -u
out=gpurun_out; mkdir -p $out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 python -m pytest tests/test_gpu_view_shard_nccl.py -q -m gpu --tb=short 2>&1 | tail -40 > $out/pytest_gpu_nccl_${N}gpu.log; tail -30 $out/pytest_gpu_nccl_${N}gpu.log
for ex in p2p nccl; do
  for cfg in 4 6; do
    timeout 900 $TR tools/run_shard_nccl.py --config $cfg --exchange $ex --repeat 3 2>$out/shard_cfg${cfg}_${ex}_${N}gpu.err | tail -1 > $out/shard_cfg${cfg}_${ex}_${N}gpu.json
    cut -c1-700 $out/shard_cfg${cfg}_${ex}_${N}gpu.json; tail -3 $out/shard_cfg${cfg}_${ex}_${N}gpu.err | cut -c1-300
  done
done
for o in 0 1; do
  timeout 300 python tools/run_mine.py --config 4 --repeat 3 --opt equal_rounds=$o 2>&1 | tail -1 > $out/exp5_cfg4_equal_rounds$o.json
  python -c "
import json; d=json.load(open('$out/exp5_cfg4_equal_rounds$o.json')); print('cfg4 equal_rounds=$o', [round(r['sweep_ms'],1) for r in d['runs']], d['mean_cost'])"
done
timeout 1200 $TR bench.py --gpus $N --steps 3 --warmup 3 2>$out/bench_${N}gpu.err | tail -1 > $out/bench_ours_${N}gpu.json; cut -c1-2500 $out/bench_ours_${N}gpu.json; tail -3 $out/bench_${N}gpu.err | cut -c1-300
timeout 1200 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>$out/bench_ref_${N}gpu.err | tail -1 > $out/bench_reference_${N}gpu.json; cut -c1-600 $out/bench_reference_${N}gpu.json
