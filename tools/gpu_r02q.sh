#!/bin/bash
# Sort keys written by the orbit step (cmcb200_dev_step; CMCB200_STEP_KEYS=0 switches it off): parity, A/B of the stage timers, and an
# ncu --set full capture with source of the radix pass after the vote-based ranking (passes 5, 6 and the last one)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_r02c_keys.log
F="\[[^]]*\] step [0-9]\|total [0-9.]* ms\|keys_hist=[0-9.]*\|radix_passes=[0-9.]*\|orbit_prepare=[0-9.]*\|orbit_tries=[0-9.]*"
timeout 400 python tools/perf_orbit.py 1e7 3 "CMCB200_STEP_KEYS=0" "CMCB200_STEP_KEYS=1" 2>&1 | grep -o "$F" | paste - - - - - - | tee gpurun_out/ab_r02c_stepkeys.log
timeout 200 python tools/perf_orbit.py 1e6 3 "CMCB200_STEP_KEYS=0" "CMCB200_STEP_KEYS=1" 2>&1 | grep -o "$F" | paste - - - - - - | tee -a gpurun_out/ab_r02c_stepkeys.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sort_onesweep_pass -s 5 -c 3 -o gpurun_out/prof_r02c_sort -f python tools/perf_orbit.py 1e7 1 > gpurun_out/ncu_r02c_sort.log 2>&1
tail -3 gpurun_out/ncu_r02c_sort.log
