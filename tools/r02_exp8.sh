#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
for v in "p2p fused_warps=16" "p2p fused_warps=8" "nccl fused_warps=16"; do
  set -- $v; ex=$1; opt=$2
  tag=fix_cfg6_${ex}_${opt#*=}_${N}gpu
  timeout 900 $TR tools/run_shard_nccl.py --config 6 --exchange $ex --opt $opt --repeat 3 2>$out/shard_$tag.err | tail -1 > $out/shard_$tag.json
  python -c "
import json
try:
    d=json.load(open('$out/shard_$tag.json')); print('$tag', 'sweep_ms', round(d['sweep_ms_max_over_ranks'],1), 'single', round(d['single_gpu_sweep_ms'],1), 'speedup', round(d['single_over_sharded_sweep'],2), 'identical', d['bit_identical_to_single_gpu'])
except Exception as e: print('$tag FAILED', e)"
  tail -2 $out/shard_$tag.err | cut -c1-300
done
