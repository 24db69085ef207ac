#!/bin/bash
# 2-GPU check (one process per GPU, NCCL): weak-scaling bench, both arms, and the NCCL source-view shard.
out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 2>$out/bench2.err | tail -1 | tee $out/bench_ours_2gpu.json | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/run_shard_nccl.py --config 4 2>$out/shard2.err | tail -1 | tee $out/view_shard_2gpu_nccl.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>$out/benchref2.err | tail -1 | cut -c1-200
tail -n 3 $out/bench2.err $out/shard2.err | grep -v "^\*\|OMP_NUM" | tail -5
