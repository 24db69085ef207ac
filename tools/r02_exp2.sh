#!/bin/bash
# round-2 experiment 2: TMA tile staging + fused shard stages + new bench line; ncu launch list and per-iteration captures
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > $out/pytest_gpu_exp2.log; tail -3 $out/pytest_gpu_exp2.log
for o in "tma=0" "tma=1"; do
  timeout 300 python tools/run_mine.py --config 2 --repeat 3 --opt $o 2>&1 | tail -1 > $out/exp2_cfg2_$o.json
  python -c "
import json; d=json.load(open('$out/exp2_cfg2_$o.json')); print('cfg2 $o', [round(r['sweep_ms'],1) for r in d['runs']], d['mean_cost'])"
done
# the sharded flow on one GPU (world 1): fused stages vs the single-context sweep
timeout 600 python tools/run_shard_nccl.py --config 4 2>&1 | tail -1 | tee $out/exp2_shard_cfg4_1gpu.json | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 2>$out/bench.err | tail -1 > $out/bench_ours_1gpu.json; cut -c1-1500 $out/bench_ours_1gpu.json
# sanitizer on sizes that are not multiples of 32 (ADVICE: window staging past the padded image)
for sz in "--rows 70 --cols 100" "--rows 64 --cols 96 --color"; do
  echo "== memcheck $sz" >> $out/sanitizer_exp2.txt
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/run_mine.py --config 2 $sz --views 4 --iters 2 --repeat 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|Error" | head -5 >> $out/sanitizer_exp2.txt
done
echo "== initcheck 70x100" >> $out/sanitizer_exp2.txt
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/run_mine.py --config 2 --rows 70 --cols 100 --views 4 --iters 2 --repeat 1 2>&1 | grep -E "ERROR SUMMARY|Uninitialized|Error" | head -5 >> $out/sanitizer_exp2.txt
cat $out/sanitizer_exp2.txt
# launch list of the run (per-launch times under ncu are serialised: only the shares are meaningful)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file $out/launches_cfg2.csv python tools/run_mine.py --config 2 --repeat 1 > /dev/null 2>&1
# full captures of the dominant kernel: black launch of iterations 1, 2, 5, 8 (k_sweep launches 0, 2, 8, 14)
for it in 0 2 8 14; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip $it --launch-count 1 -f -o $out/k_sweep_cfg2_l$it python tools/run_mine.py --config 2 --repeat 1 > /dev/null 2>&1
  ncu -i $out/k_sweep_cfg2_l$it.ncu-rep --page raw --csv > $out/k_sweep_cfg2_l${it}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $out/k_sweep_cfg2_l${it}_raw.csv > $out/ncu_k_sweep_cfg2_launch$it.txt
  grep -E "time_duration|data_pipe_tex_wavefronts|dram__bytes|issue_active" $out/ncu_k_sweep_cfg2_launch$it.txt
done
ncu -i $out/k_sweep_cfg2_l2.ncu-rep --page source --csv > $out/k_sweep_cfg2_l2_source.csv 2>/dev/null
rm -f $out/k_sweep_cfg2_l0.ncu-rep $out/k_sweep_cfg2_l8.ncu-rep $out/k_sweep_cfg2_l14.ncu-rep
