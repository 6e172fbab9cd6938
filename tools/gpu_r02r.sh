#!/bin/bash
# Radix pass: tile aggregate published before the ranking (+ windowed look-back) -- parity of the sort tests, stage timers at N = 1e7 / 1e6
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort or step" 2>&1 | tail -3 | tee gpurun_out/pytest_r02c_early.log
F="\[[^]]*\] step [0-9]\|total [0-9.]* ms\|keys_hist=[0-9.]*\|radix_passes=[0-9.]*\|record_gather=[0-9.]*"
timeout 400 python tools/perf_orbit.py 1e7 3 "" 2>&1 | grep -o "$F" | paste - - - - - | tee gpurun_out/ab_r02c_earlyagg.log
timeout 200 python tools/perf_orbit.py 1e6 3 "" 2>&1 | grep -o "$F" | paste - - - - - | tee -a gpurun_out/ab_r02c_earlyagg.log
