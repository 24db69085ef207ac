#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 1200 $TR bench.py --gpus $N --steps 3 --warmup 3 > $out/bench_${N}gpu.stdout 2>$out/bench_${N}gpu.err
grep '^{' $out/bench_${N}gpu.stdout | tail -1 > $out/bench_ours_${N}gpu.json
grep -h -E "NCCL INFO (Connected|comm|ncclCommInitRank|Using network|NVLS)" $out/bench_${N}gpu.stdout $out/bench_${N}gpu.err | head -30 > $out/nccl_init_${N}gpu.log
python -c "
import json; d=json.load(open('$out/bench_ours_${N}gpu.json')); print('bench N=$N', d['value'], d['bit_identical_to_single_gpu'], d['strong_scaling'], d['e2e']['value'], d['collective']['exchange'], d['clocks'])"; tail -3 $out/bench_${N}gpu.err | cut -c1-300; wc -l $out/nccl_init_${N}gpu.log
