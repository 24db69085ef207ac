#!/bin/bash
# N-GPU run (N = 4 or 8): view shard over real ranks (fused peer-memory exchange with 1 or 2 blocks per SM, NCCL), bench lines, hybrid
set -u
out=gpurun_out; mkdir -p $out
N=${1:-4}; G=${2:-2}; FULL=${3:-1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
if [ "$FULL" = "1" ]; then
timeout 900 python -m pytest tests/test_gpu_view_shard_nccl.py -q -m gpu --tb=short 2>&1 | tail -30 > $out/pytest_gpu_nccl_${N}gpu.log; tail -5 $out/pytest_gpu_nccl_${N}gpu.log
for v in "p2p fused_warps=8" "p2p fused_warps=16" "nccl fused_warps=8"; do
  set -- $v; ex=$1; opt=$2
  for cfg in 4 6; do
    tag=cfg${cfg}_${ex}_${opt#*=}_${N}gpu
    timeout 900 $TR tools/run_shard_nccl.py --config $cfg --exchange $ex --opt $opt --repeat 3 2>$out/shard_$tag.err | tail -1 > $out/shard_$tag.json
    python -c "
import json
try:
    d=json.load(open('$out/shard_$tag.json')); print('$tag', 'sweep_ms', round(d['sweep_ms_max_over_ranks'],1), 'single', round(d['single_gpu_sweep_ms'],1), 'speedup', round(d['single_over_sharded_sweep'],2), 'identical', d['bit_identical_to_single_gpu'])
except Exception as e: print('$tag FAILED', e)"
    tail -2 $out/shard_$tag.err | cut -c1-300
  done
done
fi
timeout 1500 $TR bench.py --gpus $N --steps 3 --warmup 3 2>$out/bench_${N}gpu.err | tail -1 > $out/bench_ours_${N}gpu.json; python -c "
import json; d=json.load(open('$out/bench_ours_${N}gpu.json')); print('bench N=$N', d['value'], d['bit_identical_to_single_gpu'], d['strong_scaling'], d['e2e']['value'], d['collective']['exchange'])"; tail -3 $out/bench_${N}gpu.err | cut -c1-300
timeout 1500 $TR bench.py --gpus $N --mode hybrid --shard $G --steps 2 --warmup 3 2>$out/bench_hybrid_${N}gpu.err | tail -1 > $out/bench_ours_hybrid_${N}gpu.json; python -c "
import json; d=json.load(open('$out/bench_ours_hybrid_${N}gpu.json')); print('hybrid N=$N', d['value'], d['bit_identical_to_single_gpu'], d['strong_scaling'], d['config']['workload'])"; tail -3 $out/bench_hybrid_${N}gpu.err | cut -c1-300
