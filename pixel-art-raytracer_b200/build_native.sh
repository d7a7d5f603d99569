#!/usr/bin/env bash
# Build libpar_b200.so (CUDA kernels + C ABI + host-side scene helpers) for sm_100a, in-tree.
# nvcc cross-compiles without a GPU.  -fmad=false: the reference is built for baseline
# x86-64 (no FMA), so no fp32 multiply-add may be contracted.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
root="$(dirname "$here")"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
out="$here/par_b200/libpar_b200.so"
mkdir -p "$here/build"
FLAGS=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false
       -Xcompiler -fPIC -Xcompiler -Wall -I"$root/include" -I"$here/csrc" ${PAR_NVCC_EXTRA:-})
objs=()
for src in "$here"/csrc/*.cu "$here"/host/host_scene.cpp; do
    obj="$here/build/$(basename "${src%.*}").o"
    if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ -n "$(find "$here/csrc" "$root/include" -newer "$obj" \( -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
        "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj"
    fi
    objs+=("$obj")
done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a "${objs[@]}" -o "$out" -ldl
# headless C++ host: stand-in for the reference's main loop on top of the C ABI
CXXB="$( [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++ )"
"$CXXB" -std=c++20 -O2 -I"$root/include" "$here/host/par_headless.cpp" -L"$here/par_b200" -lpar_b200 \
    -Wl,-rpath,'$ORIGIN/../par_b200' -o "$here/build/par_headless"
echo "built $out and $here/build/par_headless"
