#!/bin/bash
# ncu --set full of the two record gathers (direct stores / staged through shared memory with cp.async)
mkdir -p gpurun_out
CMCB200_GATHER_STAGED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sort_gather_records -c 1 -o gpurun_out/prof_r02c_gather_staged -f python tools/perf_orbit.py 1e7 1 "CMCB200_GATHER_STAGED=1" > gpurun_out/ncu_r02c_gather.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sort_gather_records -c 1 -o gpurun_out/prof_r02c_gather_direct -f python tools/perf_orbit.py 1e7 1 "CMCB200_GATHER_STAGED=0" >> gpurun_out/ncu_r02c_gather.log 2>&1
tail -2 gpurun_out/ncu_r02c_gather.log
