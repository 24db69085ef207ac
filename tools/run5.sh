set -x
mkdir -p gpurun_out/golden
python tools/make_golden.py tests/golden > gpurun_out/make_golden.log 2>&1; cp tests/golden/*.npz gpurun_out/golden/; tail -5 gpurun_out/make_golden.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/bench_ours.err gpurun_out/bench_ref.err
