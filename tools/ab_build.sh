#!/usr/bin/env bash
# Developer helper: build a variant of libpar_b200.so with extra nvcc flags into
# pixel-art-raytracer_b200/build/variants/<name>/libpar_b200.so (travels to the GPU box; select it
# with PAR_B200_LIB=<path>).   tools/ab_build.sh six "-DPAR_TILE_MIN_CTAS=6 -DPAR_TILE_LIST_CAP=256 ..."
set -euo pipefail
name=${1:?variant name}; extra=${2:-}
root="$(cd "$(dirname "$0")/.." && pwd)"; pkg="$root/pixel-art-raytracer_b200"
out="$pkg/build/variants/$name"; mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
objs=()
for src in "$pkg"/csrc/*.cu "$pkg"/host/host_scene.cpp; do
    obj="$out/$(basename "${src%.*}").o"
    "$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC \
        -I"$root/include" -I"$pkg/csrc" $extra -c "$src" -o "$obj" &
    objs+=("$obj")
done
wait
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a "${objs[@]}" -o "$out/libpar_b200.so" -ldl
echo "$out/libpar_b200.so"
