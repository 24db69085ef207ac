python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu9.log
python tools/gpu_probe2.py --case v10 b25 v47 --levels steps full > gpurun_out/probe9.txt 2>&1; python tools/summarize_probe.py gpurun_out/probe9.txt | grep -E "==|exact|speedup|mismatch"
python tools/run_mine.py --config 2 > gpurun_out/mine9_cfg2.json 2>&1
python tools/run_mine.py --config 2 --opt memo=0 --repeat 1 > gpurun_out/mine9_cfg2_nomemo.json 2>&1
python tools/run_mine.py --config 2 --opt packed=0 --repeat 1 > gpurun_out/mine9_cfg2_nopacked.json 2>&1
python tools/run_mine.py --config 3 --repeat 1 > gpurun_out/mine9_cfg3.json 2>&1
python tools/run_mine.py --config 4 --repeat 1 > gpurun_out/mine9_cfg4.json 2>&1
cat gpurun_out/mine9_cfg2.json gpurun_out/mine9_cfg2_nomemo.json gpurun_out/mine9_cfg2_nopacked.json gpurun_out/mine9_cfg3.json gpurun_out/mine9_cfg4.json | cut -c1-420
