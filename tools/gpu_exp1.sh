#!/bin/bash
# knob sweep for the orbit step at N = 1e7 + ncu of two variants of the prepare kernel
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv > $OUT/pcie.txt 2>&1
timeout 900 python tools/perf_orbit.py 1e7 2 "CMCB200_ORBIT_LADDERS=0" "" "CMCB200_QLAD_FACTOR_LOG2=3" "CMCB200_QLAD_MAX_SHIFT=10,CMCB200_QLAD_MAX_WIDTH_LOG2=14" \
  "CMCB200_QLAD_MAX_SHIFT=7,CMCB200_QLAD_MAX_WIDTH_LOG2=11" "CMCB200_QLAD_MAX_SHIFT=4,CMCB200_QLAD_MAX_WIDTH_LOG2=9" "CMCB200_QLAD_MAX_SHIFT=10" \
  "CMCB200_ORBIT_LADDERS=0,CMCB200_ORBIT_CTAS=5" "CMCB200_ORBIT_LADDERS=0,CMCB200_ORBIT_CTAS=6" "CMCB200_ORBIT_CTAS=5" > $OUT/exp1.log 2>&1
grep "step 1" $OUT/exp1.log | sed 's/status=.*| / | /' | cut -c1-330
timeout 600 env CMCB200_ORBIT_LADDERS=0 ncu --set full --clock-control none --import-source on -k regex:"orbit_prepare_kernel" -c 1 -f -o $OUT/prof_prepare_noladder python tools/perf_orbit.py 1e7 1 > $OUT/ncu_nl.log 2>&1
