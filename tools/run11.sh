python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu11.log
python tools/run_mine.py --config 2 > gpurun_out/mine11_cfg2.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches11.csv python tools/run_mine.py --config 2 --repeat 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 0 -c 1 -o gpurun_out/mine11_sweep_it1 python tools/run_mine.py --config 2 --iters 1 --repeat 1 > gpurun_out/ncu11a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 8 -c 1 -o gpurun_out/mine11_sweep_it5 python tools/run_mine.py --config 2 --iters 5 --repeat 1 > gpurun_out/ncu11b.log 2>&1
cat gpurun_out/mine11_cfg2.json | cut -c1-400
