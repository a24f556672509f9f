#!/bin/bash
# tools/build_variant2.sh NAME "TU1 TU2 ..." [nvcc flags...]: A/B build of several translation units with extra flags -> variants/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; tus=$2; shift; shift
mkdir -p variants
objs=""; skip=""
for tu in $tus; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -x cu -c bsmap_b200/csrc/$tu -o variants/$name.$tu.o &
done
wait
for tu in $tus; do objs="$objs variants/$name.$tu.o"; done
rest=$(ls bsmap_b200/build/*.o | grep -v -F "$(for tu in $tus; do echo /$tu.o; done)")
nvcc -shared -o variants/$name.so $objs $rest -gencode arch=compute_100a,code=sm_100a -lpthread -lz
rm -f $objs
echo variants/$name.so
