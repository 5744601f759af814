#!/bin/bash
# Development aid: build an experimental variant of the kernels into pypybox2d_b200/csrc/variants/libb2b_<name>.so
# (git-ignored).  usage: tools/build_variant.sh <name> [-DMACRO=value ...];  tools/variant_sweep.py times them.
# All three units are compiled with the macros (the layout functions in b2b_step.cuh are inline: mixing objects built
# with different macros would let the linker pick either definition).  SRC=<dir> builds another copy of the sources.
set -e
cd "$(dirname "$0")/../pypybox2d_b200/csrc"
OUT=$(pwd)/variants
INC=$(cd ../../include && pwd)
name=$1; shift
mkdir -p $OUT
T=$(mktemp -d)
cp ${SRC:-.}/*.cu ${SRC:-.}/*.cuh ${SRC:-.}/*.h ${SRC:-.}/*.inl $T/
mkdir -p $T/../../include 2>/dev/null || true
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$INC $ARCH"
( cd $T && sed -i "s#\"../../include/b2b.h\"#\"$INC/b2b.h\"#" *.cu *.cuh *.h *.inl 2>/dev/null;
  nvcc "$@" $FLAGS -c b2b_f32.cu -o f32.o &
  nvcc "$@" $FLAGS -fmad=false -c b2b_f64.cu -o f64.o &
  nvcc "$@" $FLAGS -c b2b_api.cu -o api.o & wait
  nvcc $ARCH -shared -o $OUT/libb2b_$name.so f32.o f64.o api.o -cudart static )
rm -rf $T
echo built variants/libb2b_$name.so
