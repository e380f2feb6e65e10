#!/bin/bash
# Builds an experimental variant of librtb200.so with extra -D flags into ray-tracing-v06_b200/variants/<name>.so
# usage: tools/build_variant.sh <name> [-DFOO=1 ...]      (select at run time with RTB_LIB=<path>)
set -e
cd "$(dirname "$0")/../ray-tracing-v06_b200"
name=$1; shift
mkdir -p variants build/var_$name
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off -w $*"
$NV -x cu -c csrc/rtb_scene.cpp -o build/var_$name/rtb_scene.o &
$NV -c csrc/rtb_kernels.cu -o build/var_$name/rtb_kernels.o &
$NV -c csrc/rtb_render.cu -o build/var_$name/rtb_render.o &
$NV -c csrc/rtb_lbvh.cu -o build/var_$name/rtb_lbvh.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name.so build/var_$name/*.o -cudart static
echo built variants/$name.so
