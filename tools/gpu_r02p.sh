#!/bin/bash
# Radix pass ranking by per-bit votes (default) against MATCH.ANY (CMCB200_SORT_RANK=0): parity of the sort / potential / step tests with the
# new build, then the A/B of the stage timers at N = 1e7 and 1e6
set -o pipefail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort or potential or step or lagrange or energy_totals or sigma" 2>&1 | tail -5 | tee gpurun_out/pytest_r02c_sort.log
timeout 400 python tools/perf_orbit.py 1e7 3 "CMCB200_SORT_RANK=0" "CMCB200_SORT_RANK=1" 2>&1 | grep -o "\[[^]]*\] step [0-9]\|total [0-9.]* ms\|keys_hist=[0-9.]*\|radix_passes=[0-9.]*\|record_gather=[0-9.]*" | paste - - - - - | tee gpurun_out/ab_r02c_sortrank.log
timeout 200 python tools/perf_orbit.py 1e6 3 "CMCB200_SORT_RANK=0" "CMCB200_SORT_RANK=1" 2>&1 | grep -o "\[[^]]*\] step [0-9]\|total [0-9.]* ms\|keys_hist=[0-9.]*\|radix_passes=[0-9.]*\|record_gather=[0-9.]*" | paste - - - - - | tee -a gpurun_out/ab_r02c_sortrank.log
