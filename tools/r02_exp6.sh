#!/bin/bash
# 1-GPU validation of the identity memo build: whole GPU suite, timings, DRAM traffic per launch, warps-per-block at cfg 3
set -u
out=gpurun_out; mkdir -p $out
timeout 1800 python -m pytest tests -q -m gpu --tb=short --durations=5 2>&1 | tail -40 > $out/pytest_gpu_exp6.log; tail -25 $out/pytest_gpu_exp6.log
for cfg in 2 3 4; do
  timeout 600 python tools/run_mine.py --config $cfg --repeat 3 2>&1 | tail -1 > $out/exp6_cfg${cfg}.json
  python -c "
import json; d=json.load(open('$out/exp6_cfg${cfg}.json')); print('cfg$cfg', [round(r['sweep_ms'],1) for r in d['runs']], d['mean_cost'], d['runs'][-1]['stats'])"
done
for nw in 8 12; do
  timeout 600 python tools/run_mine.py --config 3 --repeat 2 --opt nwarps=$nw 2>&1 | tail -1 > $out/exp6_cfg3_nw$nw.json
  python -c "
import json; d=json.load(open('$out/exp6_cfg3_nw$nw.json')); print('cfg3 nwarps=$nw', [round(r['sweep_ms'],1) for r in d['runs']])"
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $out/launches_cfg2_idmemo.csv python tools/run_mine.py --config 2 --repeat 1 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$out/launches_cfg2_idmemo.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value')
d={}
for r in rows[h+1:]:
    if len(r)>vi: d.setdefault((int(r[0]), r[ki]),{})[r[mi]]=float(r[vi].replace(',',''))
sw=[v for k,v in sorted(d.items()) if 'k_sweep' in k[1]]
print("k_sweep launches", len(sw), "avg DRAM MB/launch", sum(v['dram__bytes_read.sum']+v['dram__bytes_write.sum'] for v in sw)/len(sw)/1e6, "sum ms", sum(v['gpu__time_duration.sum'] for v in sw)/1e6)
PY
timeout 900 python bench.py --steps 3 --warmup 3 2>$out/bench.err | tail -1 > $out/bench_ours_1gpu.json; python -c "
import json; d=json.load(open('$out/bench_ours_1gpu.json')); print(d['value'], d['e2e'], d['roofline']['binding_unit'], d['work'], d['cpu_baseline'])"
