#!/usr/bin/env bash
# oracle/build_ref.sh — build the pinned reference (the parity oracle / "reference CUDA path").
#
# TEST / BENCH INFRASTRUCTURE.  Compiles the reference's gipuma.cu FROM WHERE IT LIES under
# /root/reference (never copied into the repo) together with oracle/harness/hx_harness.cu into
#   oracle/_ref/libhx_ref.so     stock source, <= 32 source views (costVector[32], gipuma.cu:736)
#   oracle/_ref/libhx_ref64.so   pin P3: a sed-patched scratch copy (temp dir, deleted after the build) with costVector[64] for V > 32
# Pins P1 (fixed curand seed) and P2 (zeroed gs.cs) are applied by macro inside the harness TU,
# see hx_harness.cu.  Flags mirror the reference's CMakeLists.txt:23 (-O3 --use_fast_math) with
# the gencode replaced by sm_100a (the reference stops at compute_75).  The reference's own
# CMake build is unusable here (FindCUDA removed in CMake 4, sm_30 rejected by nvcc 12.9).
# Outputs go only to oracle/_ref/ (git-ignored, NOT gpurun-ignored: travels to the GPU box).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${GIPUMA_REFERENCE:-/root/reference}"
out="$here/_ref"
if [ ! -f "$ref/gipuma.cu" ]; then
    echo "build_ref.sh: $ref/gipuma.cu not found (expected on the GPU box) — keeping prebuilt files" >&2
    exit 0
fi
mkdir -p "$out"
scratch="$(mktemp -d "${TMPDIR:-/tmp}/hx_ref_src.XXXXXX")"      # the sed-patched copy never stays in the repo tree
trap 'rm -rf "$scratch"' EXIT
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 --use_fast_math -std=c++14 -gencode arch=compute_100a,code=sm_100a -lineinfo
       -Xcompiler -fPIC -shared -w -I"$here/harness/shim" -I"$ref" -I"$here/harness")

stamp="$out/.stamp_ref"
src_sig="$(cat "$ref/gipuma.cu" "$here/harness/hx_harness.cu" "$here/harness/hx_api.h" "$0" | sha1sum | cut -d' ' -f1)"
if [ -f "$stamp" ] && [ "$(cat "$stamp")" = "$src_sig" ] && [ -f "$out/libhx_ref.so" ] && [ -f "$out/libhx_ref64.so" ]; then
    echo "build_ref.sh: up to date"
    exit 0
fi

echo "build_ref.sh: libhx_ref.so (stock, 32 views)"
"$NVCC" "${FLAGS[@]}" -DHX_BACKEND_REFERENCE -DHX_GIPUMA_CU="\"$ref/gipuma.cu\"" -DHX_MAX_VIEWS=32 \
    -o "$out/libhx_ref.so" "$here/harness/hx_harness.cu"

echo "build_ref.sh: libhx_ref64.so (pin P3: costVector[64])"
sed 's/float costVector\[32\];/float costVector[64];/' "$ref/gipuma.cu" > "$scratch/gipuma_v64.cu"
grep -q 'costVector\[64\]' "$scratch/gipuma_v64.cu"
"$NVCC" "${FLAGS[@]}" -DHX_BACKEND_REFERENCE -DHX_GIPUMA_CU="\"$scratch/gipuma_v64.cu\"" -DHX_MAX_VIEWS=64 \
    -o "$out/libhx_ref64.so" "$here/harness/hx_harness.cu"
echo "$src_sig" > "$stamp"
echo "build_ref.sh: done"
