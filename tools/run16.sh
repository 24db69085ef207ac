nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 2>gpurun_out/bench4.err | tail -1 | tee gpurun_out/bench_ours_4gpu.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 tools/run_shard_nccl.py --config 4 2>gpurun_out/shard4.err | tail -1 | tee gpurun_out/shard_world4.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>gpurun_out/benchref4.err | tail -1 | cut -c1-200
tail -n 3 gpurun_out/bench4.err gpurun_out/shard4.err | grep -v "^\*\|OMP_NUM" | tail -5
