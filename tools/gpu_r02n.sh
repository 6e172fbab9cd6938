#!/bin/bash
# A/B: the two index bisections of stage A with their sample-ladder descents in lockstep (CMCB_SLAD_PAIR) against one after the other,
# at 4 and 3 resident CTAs per SM
for v in p0 ""; do
	lib=$PWD/cmc_cosmic_b200/libcmcb200.so; [ -n "$v" ] && lib=$PWD/tmp_ab/libcmcb200_$v.so
	for ctas in 4 3; do
		echo "== variant '${v:-pair}' CTAs/SM $ctas"; CMCB200_ORBIT_CTAS=$ctas CMCB200_LIB=$lib python tools/perf_orbit.py 1e7 3 2>&1 | tail -1 | grep -o "total [0-9.]* ms\|orbit_prepare=[0-9.]*\|orbit_tries=[0-9.]*" | paste - - -
	done
	CMCB200_LIB=$lib python tools/perf_orbit.py 1e6 3 2>&1 | tail -1 | grep -o "total [0-9.]* ms\|orbit_prepare=[0-9.]*" | paste - -
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -x -q -m gpu -k "orbit or fullsize or find_root or whole_timestep" 2>&1 | tail -3
