#!/bin/bash
# A/B through perf_orbit.py's knob strings (it clears CMCB200_* from the environment and applies its arguments):
# FP32-arithmetic screening of the sample ladder's coarsest level (CMCB200_ORBIT_LADDERS bit 6: 9 = off, 73 = on), resident CTAs per SM of
# stage A, the L2's DRAM fetch granularity
python tools/perf_orbit.py 1e7 3 "CMCB200_ORBIT_LADDERS=9" "CMCB200_ORBIT_LADDERS=73" "CMCB200_ORBIT_LADDERS=73,CMCB200_ORBIT_CTAS=5" "CMCB200_ORBIT_LADDERS=73,CMCB200_ORBIT_CTAS=6" "CMCB200_ORBIT_LADDERS=9,CMCB200_ORBIT_CTAS=5" "CMCB200_L2_FETCH=32" "CMCB200_L2_FETCH=128" 2>&1 | grep "step 2" | grep -o "\[[^]]*\]\|total [0-9.]* ms\|orbit_prepare=[0-9.]*\|orbit_tries=[0-9.]*\|record_gather=[0-9.]*\|exact=[0-9]*" | paste - - - - - -
python tools/perf_orbit.py 1e6 3 "CMCB200_ORBIT_LADDERS=9" "CMCB200_ORBIT_LADDERS=73" 2>&1 | grep "step 2" | grep -o "\[[^]]*\]\|total [0-9.]* ms\|orbit_prepare=[0-9.]*" | paste - - -
