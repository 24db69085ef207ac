python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu8.log
python tools/gpu_probe2.py --case v10 b25 v47 --levels full > gpurun_out/probe8.txt 2>&1; python tools/summarize_probe.py gpurun_out/probe8.txt | grep -E "==|exact|speedup|mismatch"
python tools/run_mine.py --config 2 > gpurun_out/mine8_cfg2.json 2>&1
python tools/run_mine.py --config 2 --opt packed=0 --repeat 1 > gpurun_out/mine8_cfg2_nopacked.json 2>&1
python tools/run_mine.py --config 3 --repeat 1 > gpurun_out/mine8_cfg3.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/mine8_sweep_cfg2 python tools/run_mine.py --config 2 --iters 2 --repeat 1 > gpurun_out/ncu_mine8.log 2>&1
cat gpurun_out/mine8_cfg2.json gpurun_out/mine8_cfg2_nopacked.json gpurun_out/mine8_cfg3.json | cut -c1-420
