#!/bin/bash
# round 2: FP32 screening of the index bisections -- A/B of the stage timers, then the whole GPU suite
python tools/perf_orbit.py 1e7 3 "CMCB200_ORBIT_LADDERS=1" "CMCB200_ORBIT_LADDERS=5" 2>&1 | tee gpurun_out/screen_ab_r02_n1e7.log | cut -c1-400
python tools/perf_orbit.py 1e6 3 "CMCB200_ORBIT_LADDERS=1" "CMCB200_ORBIT_LADDERS=5" 2>&1 | tee gpurun_out/screen_ab_r02_n1e6.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/gputest_r02_1gpu.log
