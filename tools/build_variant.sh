#!/bin/bash
# Another build of the library with extra compile-time knobs, for A/B measurements through CMCB200_LIB:
#   tools/build_variant.sh v512 -DCMCB_SORT_THREADS=512 -DCMCB_SORT_ITEMS=8   ->  tools/_build/v512/libcmcb200.so
set -e
name=$1; shift
src=cmc_cosmic_b200/csrc
out=tools/_build/$name
mkdir -p $out
for f in api sort potential orbit shard lagrange velocities dynamics sigma snapshot; do
	nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC "$@" -c $src/$f.cu -o $out/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libcmcb200.so $out/*.o -lcudart
ls -la $out/libcmcb200.so
