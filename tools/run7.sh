nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_shard_nccl.py --config 4 2>gpurun_out/shard2.err | tail -1 | tee gpurun_out/shard_world2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/bench2.err | tail -1 | tee gpurun_out/bench_ours_2gpu.json
tail -n 5 gpurun_out/shard2.err gpurun_out/bench2.err
