#!/bin/bash
# Builds libsnb.so for sm_100a in-tree (shennong_b200/_build/libsnb.so).
# nvcc cross-compiles without a GPU; the .so travels with the gpurun snapshot.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_build"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
       -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
       ${SNB_EXTRA_NVCC_FLAGS})
for f in tables.cc io.cc features.cu post.cu pitch.cu peer.cu resample.cu; do
  o="$OUT/${f%.*}.o"
  if [ ! -f "$o" ] || [ "$HERE/$f" -nt "$o" ] || [ "$HERE/snb_internal.h" -nt "$o" ] \
     || [ "$HERE/device_utils.cuh" -nt "$o" ] || [ "$HERE/../../include/snb.h" -nt "$o" ]; then
    echo "nvcc $f"
    "$NVCC" "${FLAGS[@]}" -x cu -c "$HERE/$f" -o "$o" $SNB_PTXAS
  fi
done
"$NVCC" "${FLAGS[@]}" -shared -o "$OUT/libsnb.so" "$OUT"/tables.o "$OUT"/io.o "$OUT"/features.o "$OUT"/post.o "$OUT"/pitch.o "$OUT"/peer.o "$OUT"/resample.o
echo "built $OUT/libsnb.so"
