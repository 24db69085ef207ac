python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
python tools/run_shard_nccl.py --config 4 2>&1 | tail -2 | tee gpurun_out/shard_world1.json
tail -n 3 gpurun_out/bench_ours.err; tail -n 3 gpurun_out/bench_ref.err
