#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_rng_stateful.py tests/test_gpu_view_shard.py tests/test_gpu_fused_sweep.py tests/test_gpu_batch_driver.py -q -m gpu --tb=short -x 2>&1 | tail -80 > $out/pytest_gpu_exp4.log; tail -60 $out/pytest_gpu_exp4.log
timeout 900 python tools/packed_probe.py --config 2 2>&1 | tail -1 > $out/packed_probe_cfg2.json; cut -c1-1500 $out/packed_probe_cfg2.json
