#!/bin/bash
# 4-GPU check (one process per GPU, NCCL): weak-scaling bench and the NCCL source-view shard.
out=gpurun_out; mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 2>$out/bench4.err | tail -1 | tee $out/bench_ours_4gpu.json | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 tools/run_shard_nccl.py --config 4 2>$out/shard4.err | tail -1 | tee $out/view_shard_4gpu_nccl.json
