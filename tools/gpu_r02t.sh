#!/bin/bash
# Record gather through shared memory (whole 128-byte lines out) against the direct stores (CMCB200_GATHER_STAGED=0): parity, stage timers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py -m gpu -x -q -k "sort or step or resident or restart" 2>&1 | tail -3 | tee gpurun_out/pytest_r02c_gather.log
F="\[[^]]*\] step [0-9]\|total [0-9.]* ms\|radix_passes=[0-9.]*\|record_gather=[0-9.]*\|binaries=[0-9.]*"
timeout 400 python tools/perf_orbit.py 1e7 3 "CMCB200_GATHER_STAGED=0" "CMCB200_GATHER_STAGED=1" 2>&1 | grep -o "$F" | paste - - - - - | tee gpurun_out/ab_r02c_gather.log
timeout 200 python tools/perf_orbit.py 1e6 3 "CMCB200_GATHER_STAGED=0" "CMCB200_GATHER_STAGED=1" 2>&1 | grep -o "$F" | paste - - - - - | tee -a gpurun_out/ab_r02c_gather.log
