#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2: resident upload + chunked orbit step, sample ladder, compact jobs, Brent cluster,
# comp_multi_mass_percent (small case), the resident drop-in shim.  One gpurun call; logs go to gpurun_out/
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/memcheck_r02.log \
  python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py tests/test_dropin_resident.py -m gpu -x -q \
  -k "5000-True or find_root_vr_fires or multi_mass_percent_bit_exact and 30000 or zeros_stay or resident_shim and plummer-24000 or orbit_step_bit_exact and plummer" > $OUT/memcheck_r02_pytest.log 2>&1
echo "rc=$?" >> $OUT/memcheck_r02_pytest.log
tail -5 $OUT/memcheck_r02_pytest.log; grep -c "Invalid\|Error" $OUT/memcheck_r02.log; tail -5 $OUT/memcheck_r02.log
