#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
GPM_PREPASS=1 timeout 400 python -m pytest tests/test_gpu_parity_golden.py tests/test_gpu_live_reference.py tests/test_gpu_fused_sweep.py tests/test_gpu_properties_fullsize.py tests/test_gpu_rng_stateful.py tests/test_gpu_batch_driver.py tests/test_gpu_view_shard.py -q -m gpu --tb=short 2>&1 | tail -12 > $out/pytest_gpu_last.log; tail -6 $out/pytest_gpu_last.log
timeout 200 python tools/run_mine.py --config 2 --repeat 3 --opt prepass=1 2>&1 | tail -1 > $out/last_cfg2.json; python -c "
import json; d=json.load(open('$out/last_cfg2.json')); print('cfg2', [round(r['sweep_ms'],1) for r in d['runs']], d['mean_cost'], d['runs'][-1]['stats'])"
GPM_PREPASS=1 timeout 300 python bench.py --steps 3 --warmup 3 2>$out/bench.err | tail -1 > $out/bench_ours_1gpu_last.json; python -c "
import json; d=json.load(open('$out/bench_ours_1gpu_last.json')); print(d['value'], d['e2e']['value'], d['work'])"
timeout 200 python tools/run_mine.py --config 2 --repeat 2 --opt prepass=0 2>&1 | tail -1 > $out/last_cfg2_off.json; python -c "
import json; d=json.load(open('$out/last_cfg2_off.json')); print('cfg2 prepass=0', [round(r['sweep_ms'],1) for r in d['runs']], d['mean_cost'], d['runs'][-1]['stats'])"
