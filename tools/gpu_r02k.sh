#!/bin/bash
# rejection rounds: lookup kernel split off (bit 5), ladder lines whole (bit 4) or sector by sector -- A/B of the stage timers, then the GPU suite
python tools/perf_orbit.py 1e7 3 "CMCB200_ORBIT_LADDERS=9" "CMCB200_ORBIT_LADDERS=25" "CMCB200_ORBIT_LADDERS=41" "CMCB200_ORBIT_LADDERS=57" 2>&1 | grep "step 2" | tee gpurun_out/split_ab_r02_n1e7.log | cut -c1-420
python tools/perf_orbit.py 1e6 3 "CMCB200_ORBIT_LADDERS=9" "CMCB200_ORBIT_LADDERS=41" "CMCB200_ORBIT_LADDERS=57" 2>&1 | grep "step 2" | tee gpurun_out/split_ab_r02_n1e6.log | cut -c1-420
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/gputest_r02_1gpu.log
