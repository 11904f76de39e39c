#!/bin/bash
# usage: bash tests/ubench/build_variant.sh <name> <file.cu> [-DMACRO=VALUE ...]  ->  .variants/libgbops_<name>.so (run with GBOPS_LIB=...)
# One translation unit is rebuilt with the extra flags and linked against the objects of the regular build.
set -e
NAME=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
cd "$ROOT/graspbalance_b200/csrc"
mkdir -p "$ROOT/.variants"
OBJ="$ROOT/.variants/${SRC%.cu}_$NAME.o"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden "$@" -c "$SRC" -o "$OBJ"
OTHERS=$(ls *.o | grep -v "^${SRC%.cu}.o$")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/.variants/libgbops_$NAME.so" $OTHERS "$OBJ"
echo "$ROOT/.variants/libgbops_$NAME.so"
