#!/usr/bin/env bash
# Builds the product library (hand-written CUDA for sm_100a behind the C ABI of include/cugs_b200.h)
# in-tree: cuda_gaussian_splatting_b200/libcugs_b200.so. No torch headers -> seconds per file.
set -euo pipefail
cd "$(dirname "$0")"
SRC=${SRCDIR:-cuda_gaussian_splatting_b200/csrc}
OUT=cuda_gaussian_splatting_b200
LIBNAME=${LIBNAME:-libcugs_b200.so}
OBJ=${OBJDIR:-build/obj}
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${EXTRA_NVCC_FLAGS:-} -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default"
pids=()
for f in preprocess binning radix_sort tile_binning blend train_ops grad_exchange density_control trainer api; do
  if [ ! -f "$OBJ/$f.o" ] || [ "$SRC/$f.cu" -nt "$OBJ/$f.o" ] || [ "$SRC/common.cuh" -nt "$OBJ/$f.o" ] || [ include/cugs_b200.h -nt "$OBJ/$f.o" ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$SRC/$f.cu" -o "$OBJ/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC -Wno-deprecated-gpu-targets -shared -o "$OUT/$LIBNAME" $OBJ/preprocess.o $OBJ/binning.o $OBJ/radix_sort.o $OBJ/tile_binning.o $OBJ/blend.o $OBJ/train_ops.o $OBJ/grad_exchange.o $OBJ/density_control.o $OBJ/trainer.o $OBJ/api.o -lcudart
echo "built $OUT/$LIBNAME"
