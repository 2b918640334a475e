#!/bin/bash
# Build libcnhead_sm100.so in-tree (sm_100a only).  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       -I"$HERE/../../include" -I"$HERE" "$@")
pids=()
for f in api detloss softmax_stat decode raster; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$OUT/libcnhead_sm100.so" "$HERE"/obj/{api,detloss,softmax_stat,decode,raster}.o -cudart static
echo "built $OUT/libcnhead_sm100.so"
