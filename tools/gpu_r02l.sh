#!/bin/bash
# A/B: dyn_pair_kernel at 1 / 2 / 3 resident CTAs per SM; the L2's DRAM fetch granularity for the whole step
OUT=gpurun_out
for mb in "" _mb2 _mb3; do
	echo "== dyn_pair min blocks '$mb'"; CMCB200_LIB=$PWD/cmc_cosmic_b200/libcmcb200$mb.so python tools/perf_sigma.py 1e7 equal 2>&1 | tail -1 | sed 's/.*dynamics_apply/dynamics_apply/'
done
for g in 32 64 128; do   # needs the CMCB200_L2_FETCH knob (cudaLimitMaxL2FetchGranularity in cmcb200_create), removed after this measurement: no effect
	echo "== L2 fetch granularity $g"; CMCB200_L2_FETCH=$g python tools/perf_orbit.py 1e7 3 2>&1 | tail -1
	CMCB200_L2_FETCH=$g python tools/perf_sigma.py 1e7 equal 2>&1 | tail -1 | cut -c1-120
done
