#!/bin/bash
# round 2: sample ladder of the index bisections -- A/B of the stage timers, then the whole GPU suite
python tools/perf_orbit.py 1e7 3 "CMCB200_ORBIT_LADDERS=1" "CMCB200_ORBIT_LADDERS=9" 2>&1 | tee gpurun_out/slad_ab_r02_n1e7.log | cut -c1-420
python tools/perf_orbit.py 1e6 3 "CMCB200_ORBIT_LADDERS=1" "CMCB200_ORBIT_LADDERS=9" 2>&1 | tee gpurun_out/slad_ab_r02_n1e6.log | cut -c1-420
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/gputest_r02_1gpu.log
