#!/usr/bin/env bash
# Build libnuwa_b200.so for sm_100a (in-tree, so the .so travels with the repo snapshot).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I. ${NUWA_NVCC_EXTRA:-}"
mkdir -p build
objs=()
pids=()
for f in *.cu; do
  o="build/${f%.cu}.o"
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ kernels.h -nt "$o" ] || [ ../../include/nuwa_b200.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o libnuwa_b200.so "${objs[@]}"
echo "built $(pwd)/libnuwa_b200.so"
