python tools/gpu_probe2.py --case small v10 b25 v47 > gpurun_out/probe4.txt 2>&1
python tools/run_mine.py --config 2 > gpurun_out/mine4_cfg2.json 2>&1
for v in lb256x3 lb256x2 lb384x2 lb128x6; do GIPUMA_B200_LIB=$PWD/variants/gpurun_variants_$v.so python tools/run_mine.py --config 2 --repeat 2 > gpurun_out/mine4_cfg2_$v.json 2>&1; done
python tools/run_mine.py --config 3 --repeat 1 > gpurun_out/mine4_cfg3.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 2 -c 1 -o gpurun_out/mine4_sweep_cfg2 python tools/run_mine.py --config 2 --iters 2 --repeat 1 > gpurun_out/ncu_mine4.log 2>&1
python tools/summarize_probe.py gpurun_out/probe4.txt | grep -E "==|exact|speedup|mismatch"
cat gpurun_out/mine4_cfg2*.json gpurun_out/mine4_cfg3.json | cut -c1-400
