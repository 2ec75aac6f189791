#!/bin/bash
# Builds libnfftb200.so for sm_100a (in-tree, next to the sources' parent package directory).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libnfftb200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
NCCL_INC=${NCCL_INC:-/usr/include}
FLAGS="-I$NCCL_INC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ ${NFFTB_EXTRA_FLAGS}"
mkdir -p "$HERE/_obj"
pids=()
for f in plan.cu sort.cu deconv.cu spread.cu interp.cu comm.cu oned.cu twod.cu toeplitz.cu sdc.cu lean.cu; do
  ( $NVCC $FLAGS -c "$HERE/$f" -o "$HERE/_obj/${f%.cu}.o" ) &
  pids+=($!)
done
( $NVCC $FLAGS -x cu -c "$HERE/tables.cpp" -o "$HERE/_obj/tables.o" ) &
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT" "$HERE"/_obj/*.o -lcufft -lcudart -ldl -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo "built $OUT"
