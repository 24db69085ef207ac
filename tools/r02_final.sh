#!/bin/bash
# Round-2 closing evidence on one B200 (everything lands in gpurun_out/): bash tools/r02_final.sh
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 > $out/pytest_gpu_final.log; tail -3 $out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$out/bench_ref.err | tail -1 > $out/bench_reference_1gpu.json; cut -c1-200 $out/bench_reference_1gpu.json
timeout 900 python bench.py 2>$out/bench.err | tail -1 > $out/bench_ours_1gpu.json; cut -c1-300 $out/bench_ours_1gpu.json
# the hard scene (occluders, texture-less band, sensor noise), both arms
timeout 600 python bench.py --scene hard --steps 3 --warmup 3 2>>$out/bench.err | tail -1 > $out/bench_ours_hard.json; python -c "
import json; d=json.load(open('$out/bench_ours_hard.json')); print('hard ours', d['value'], d['work'])"
timeout 600 python bench.py --impl reference --scene hard --steps 2 --warmup 1 2>>$out/bench_ref.err | tail -1 > $out/bench_reference_hard.json; python -c "
import json; d=json.load(open('$out/bench_reference_hard.json')); print('hard ref', d['value'])"
# the other BASELINE configurations through bench.py, both arms (raw lines for profiles/)
for cfg in 3 4; do
  timeout 900 python bench.py --config $cfg --steps 2 --warmup 3 --no-ablation 2>>$out/bench.err | tail -1 > $out/bench_ours_cfg$cfg.json
  timeout 900 python bench.py --impl reference --config $cfg --steps 1 --warmup 1 2>>$out/bench_ref.err | tail -1 > $out/bench_reference_cfg$cfg.json
  python -c "
import json; a=json.load(open('$out/bench_ours_cfg$cfg.json')); b=json.load(open('$out/bench_reference_cfg$cfg.json')); print('cfg$cfg ours', a['value'], 'ref', b['value'], 'frac', a['roofline']['binding_unit']['frac'])"
done
timeout 900 python bench.py --config 5 --steps 1 --warmup 3 --no-ablation 2>>$out/bench.err | tail -1 > $out/bench_ours_cfg5.json; python -c "
import json; a=json.load(open('$out/bench_ours_cfg5.json')); print('cfg5 ours', a['value'], a['e2e']['value'], a['roofline']['binding_unit'])"
# launch lists: the bench command, and the staged view-shard flow on one GPU (world 1) next to the fused sweep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-ablation > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 140 --csv --log-file $out/launches_shard_world1_cfg4.csv python tools/run_shard_nccl.py --config 4 --repeat 1 > /dev/null 2>&1
python - <<PY
import csv
for f in ("$out/launches_bench.csv", "$out/launches_shard_world1_cfg4.csv"):
    try:
        rows=list(csv.reader(open(f)))
        h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
        hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
        agg={}
        for r in rows[h+1:]:
            if len(r)>vi:
                k=r[ki].split('(')[0][-40:]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))/1e6
        print(f); [print('   %-42s %4d launches %9.3f ms'%(k,v[0],v[1])) for k,v in sorted(agg.items(), key=lambda x:-x[1][1])]
    except Exception as e: print(f, 'failed', e)
PY
