#!/bin/bash
# ncu --set full of the orbit kernels after the sample ladder
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"orbit_prepare_kernel|orbit_try_kernel" -c 4 -f -o gpurun_out/prof_r02_orbit python tools/perf_orbit.py 1e7 1 > gpurun_out/ncu_full_r02_orbit.log 2>&1
tail -3 gpurun_out/ncu_full_r02_orbit.log
ls -la gpurun_out/prof_r02_orbit.ncu-rep
