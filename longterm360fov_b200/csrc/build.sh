#!/usr/bin/env bash
# Build libfov360.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v)
objs=()
pids=()
for src in "$HERE"/*.cu; do
  obj="$HERE/build/$(basename "${src%.cu}").o"
  objs+=("$obj")
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "$HERE/fov_common.cuh" -nt "$obj" || "$HERE/fov_internal.h" -nt "$obj" || "$HERE/tc_common.cuh" -nt "$obj" || "$HERE/../../include/fov360.h" -nt "$obj" || "$HERE/../../include/fov_debug.h" -nt "$obj" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" > "$obj.log" 2>&1 || { cat "$obj.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$OUT/libfov360.so" "${objs[@]}" -lcudart -ldl
echo "built $OUT/libfov360.so"
# stand-alone bring-up harness for the tensor-core kernels (tests/cuda/tc_selftest.cu)
SELF="$HERE/../../tests/cuda/tc_selftest.cu"
if [[ -f "$SELF" ]]; then
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 "$SELF" -o "$HERE/build/tc_selftest" \
      -L"$OUT" -lfov360 -Xlinker -rpath -Xlinker '$ORIGIN/../../lib' -lcudart
  echo "built $HERE/build/tc_selftest"
fi
