#!/bin/bash
# Experimental build of one translation unit with extra -D flags, linked into build/variants/NAME.so (same ABI; select
# it with NFFTB200_LIB=build/variants/NAME.so).  usage: scripts/variant.sh NAME UNIT "-DX=1 -DY=2"
set -e
NAME=$1; UNIT=$2; DEFS=$3
ROOT="$(cd "$(dirname "$0")/.." && pwd)"; HERE="$ROOT/nfft.jl_b200/csrc"
mkdir -p "$ROOT/build/variants/obj_$NAME"
NVCC=/usr/local/cuda/bin/nvcc
$NVCC -I/usr/include -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -Wno-deprecated-gpu-targets $DEFS \
  -c "$HERE/$UNIT.cu" -o "$ROOT/build/variants/obj_$NAME/$UNIT.o"
OBJS=$(ls "$HERE"/_obj/*.o | grep -v "/$UNIT.o")
$NVCC -Wno-deprecated-gpu-targets -shared -o "$ROOT/build/variants/$NAME.so" $OBJS "$ROOT/build/variants/obj_$NAME/$UNIT.o" -lcufft -lcudart -ldl -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo "built build/variants/$NAME.so"
