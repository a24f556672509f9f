#!/bin/bash
# tools/build_variant.sh NAME [nvcc flags...]: A/B build of one kernel translation unit (KERNEL_TU, default
# bsx_map_se.cu = the SE WGBS kernel) with extra flags -> variants/NAME.so
# (run the bench against it with BSMAP_B200_LIB=variants/NAME.so); variants/ is git-ignored.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
tu=${KERNEL_TU:-bsx_map_se.cu}
mkdir -p variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -x cu -c bsmap_b200/csrc/$tu -o variants/$name.se.o
objs=$(ls bsmap_b200/build/*.o | grep -v "$tu.o")
nvcc -shared -o variants/$name.so variants/$name.se.o $objs -gencode arch=compute_100a,code=sm_100a -lpthread -lz
rm variants/$name.se.o
echo variants/$name.so
