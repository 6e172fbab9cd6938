#!/bin/bash
# Radix pass tile shape: 256 threads x 16 items (default build) against 512 x 8 and 384 x 12 (tools/build_variant.sh): sort parity with each
# variant, stage timers at N = 1e7 / 1e6
mkdir -p gpurun_out
F="\[[^]]*\] step [0-9]\|total [0-9.]* ms\|keys_hist=[0-9.]*\|radix_passes=[0-9.]*\|record_gather=[0-9.]*"
: > gpurun_out/ab_r02c_sorttile.log
for v in default v512x8 v384x12; do
	if [ $v = default ]; then unset CMCB200_LIB; else export CMCB200_LIB=$PWD/tools/_build/$v/libcmcb200.so; fi
	echo "== $v" | tee -a gpurun_out/ab_r02c_sorttile.log
	if [ $v != default ]; then timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort" 2>&1 | tail -1 | tee -a gpurun_out/ab_r02c_sorttile.log; fi
	timeout 300 python tools/perf_orbit.py 1e7 3 "" 2>&1 | grep -o "$F" | paste - - - - - | tee -a gpurun_out/ab_r02c_sorttile.log
	timeout 200 python tools/perf_orbit.py 1e6 3 "" 2>&1 | grep -o "$F" | paste - - - - - | tee -a gpurun_out/ab_r02c_sorttile.log
done
