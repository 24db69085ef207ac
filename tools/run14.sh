python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu14.log
python tools/run_mine.py --config 2 | cut -c1-420
python tools/run_mine.py --config 3 --repeat 1 | cut -c1-420
python tools/bisect_opts.py 2 2>&1 | tail -7 | cut -c1-200
python tools/run_shard_nccl.py --config 4 2>&1 | tail -1
