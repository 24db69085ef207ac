#!/bin/bash
# Round-end evidence run on one B200 (everything lands in gpurun_out/):  bash tools/final_round.sh
set -u
out=gpurun_out; mkdir -p $out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > $out/pytest_gpu.log; tail -2 $out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 2>$out/bench_ref.err | tail -1 > $out/bench_reference_1gpu.json; cut -c1-220 $out/bench_reference_1gpu.json
python bench.py 2>$out/bench.err | tail -1 > $out/bench_ours_1gpu.json; cut -c1-420 $out/bench_ours_1gpu.json
for v in "--color" "--neighbours 20"; do
  tag=$(echo $v | tr -d ' -')
  python bench.py $v --steps 3 --warmup 3 2>>$out/bench.err | tail -1 > $out/bench_ours_$tag.json; cut -c1-160 $out/bench_ours_$tag.json
  python bench.py --impl reference $v --steps 2 --warmup 1 2>>$out/bench_ref.err | tail -1 > $out/bench_reference_$tag.json; cut -c1-200 $out/bench_reference_$tag.json
done
# launch list of the bench command (per-launch times under ncu are serialised: only the shares are meaningful)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 > $out/bench_under_ncu.log 2>&1
# one full capture of the dominant kernel: iteration 2, black (3rd k_sweep launch), gray and float4
ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 2 --launch-count 1 -f -o $out/k_sweep_cfg2 python tools/run_mine.py --config 2 --repeat 1 > /dev/null 2>&1
ncu -i $out/k_sweep_cfg2.ncu-rep --page raw --csv > $out/k_sweep_cfg2_raw.csv 2>/dev/null; python tools/ncu_summary.py $out/k_sweep_cfg2_raw.csv > $out/ncu_k_sweep_cfg2.txt; grep -E "time_duration|data_pipe_tex_wavefronts|dram__bytes" $out/ncu_k_sweep_cfg2.txt
ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 2 --launch-count 1 -f -o $out/k_sweep_cfg2_color python tools/run_mine.py --config 2 --repeat 1 --color > /dev/null 2>&1
ncu -i $out/k_sweep_cfg2_color.ncu-rep --page raw --csv > $out/k_sweep_cfg2_color_raw.csv 2>/dev/null; python tools/ncu_summary.py $out/k_sweep_cfg2_color_raw.csv > $out/ncu_k_sweep_cfg2_color.txt; grep -E "time_duration|data_pipe_tex_wavefronts|dram__bytes" $out/ncu_k_sweep_cfg2_color.txt
rm -f $out/*.ncu-rep
