#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_rng_stateful.py tests/test_gpu_view_shard.py tests/test_gpu_fused_sweep.py tests/test_gpu_batch_driver.py tests/test_gpu_properties_fullsize.py -q -m gpu --tb=short 2>&1 | tail -120 > $out/pytest_gpu_exp4.log; tail -100 $out/pytest_gpu_exp4.log
