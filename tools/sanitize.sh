#!/bin/bash
# compute-sanitizer over the three kernel families (gray, float4, fused 20-neighbour) on a small scene.
# Usage on a B200:  bash tools/sanitize.sh > gpurun_out/sanitizer.txt 2>&1
run() {   # tool, extra run_mine.py args...
    tool=$1; shift
    echo "== compute-sanitizer --tool $tool  run_mine.py $*"
    compute-sanitizer --tool "$tool" --error-exitcode 9 python tools/run_mine.py --config 2 --rows 64 --cols 96 --views 4 --iters 2 --repeat 1 "$@" 2>&1 \
        | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Uninitialized|hazard|Error" | head -8
}
for tool in memcheck racecheck initcheck; do
    run $tool
    run $tool --color
    run $tool --opt neighbours=20
    run $tool --color --opt neighbours=20
done
