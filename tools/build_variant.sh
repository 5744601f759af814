#!/bin/bash
# Development aid: build an experimental fp32 variant of the kernels into pypybox2d_b200/csrc/variants/libb2b_<name>.so
# (git-ignored).  usage: tools/build_variant.sh <name> [-DMACRO=value ...];  tools/variant_sweep.py times them.
set -e
cd "$(dirname "$0")/../pypybox2d_b200/csrc"
name=$1; shift
mkdir -p variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc "$@" -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include $ARCH -c b2b_f32.cu -o variants/b2b_f32_$name.o
nvcc $ARCH -shared -o variants/libb2b_$name.so variants/b2b_f32_$name.o b2b_f64.o b2b_api.o -cudart static
rm -f variants/b2b_f32_$name.o
echo built variants/libb2b_$name.so
