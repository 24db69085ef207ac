python tools/run_ref.py --config 4 2>&1 | tail -1 | tee gpurun_out/ref_cfg4.json | cut -c1-300
python tools/run_ref.py --config 5 2>&1 | tail -1 | tee gpurun_out/ref_cfg5.json | cut -c1-300
python tools/run_mine.py --config 4 --repeat 2 | cut -c1-200
python -m pytest tests/test_gpu_live_reference.py tests/test_gpu_properties_fullsize.py -x -q -m gpu 2>&1 | tail -3
