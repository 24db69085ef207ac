python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu18.log
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/final_sweep_cfg2_it2 python tools/run_mine.py --config 2 --iters 2 --repeat 1 > gpurun_out/ncu18.log 2>&1
python tools/run_mine.py --config 1 --repeat 3 | cut -c1-300
