#!/bin/bash
# Incremental build of libnfftb200.so: recompile only the named sources (e.g. `scripts/rebuild.sh spread interp`), relink.
set -e
HERE="$(cd "$(dirname "$0")/../nfft.jl_b200/csrc" && pwd)"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-I/usr/include -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -Wno-deprecated-gpu-targets ${NFFTB_EXTRA_FLAGS}"
# a change of a header every object sees (the plan struct!) means every object must be rebuilt
set -- "$@"
for o in "$HERE"/_obj/*.o; do
  if [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/../../include/nfftb200.h" -nt "$o" ]; then set -- plan sort deconv spread interp comm oned twod toeplitz sdc lean tables; break; fi
done
echo "compiling: $*"
pids=()
for n in "$@"; do
  if [ "$n" = tables ]; then ( $NVCC $FLAGS -x cu -c "$HERE/tables.cpp" -o "$HERE/_obj/tables.o" ) & else ( $NVCC $FLAGS -c "$HERE/$n.cu" -o "$HERE/_obj/$n.o" ) & fi
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -Wno-deprecated-gpu-targets -shared -o "$HERE/../libnfftb200.so" "$HERE"/_obj/*.o -lcufft -lcudart -ldl -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo "relinked"
