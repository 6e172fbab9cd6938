#!/bin/bash
# A/B: secant-guided search on the coarsest level of the sample ladder (find_zero_Q_sampled): 0 (plain halving) / 5 / 8 (default) / 12 guided probes
for v in s0 s5 "" s12; do
	lib=$PWD/cmc_cosmic_b200/libcmcb200.so; [ -n "$v" ] && lib=$PWD/tmp_ab/libcmcb200_$v.so
	echo "== secant steps variant '${v:-default 8}'"; CMCB200_LIB=$lib python tools/perf_orbit.py 1e7 3 2>&1 | tail -1 | grep -o "total [0-9.]* ms\|orbit_prepare=[0-9.]*\|orbit_tries=[0-9.]*" | paste - - -
	CMCB200_LIB=$lib python tools/perf_orbit.py 1e6 3 2>&1 | tail -1 | grep -o "total [0-9.]* ms\|orbit_prepare=[0-9.]*" | paste - -
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -x -q -m gpu -k "orbit or fullsize or find_root or whole_timestep" 2>&1 | tail -3
