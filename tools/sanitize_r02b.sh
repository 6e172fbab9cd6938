#!/bin/bash
# compute-sanitizer memcheck over the kernels of the second half of round 2: calc_sigma_r (all passes, shards with ghost stars, the sequential
# replay), dynamics_apply's pair loop (encounters, threebb compaction), simul_relax_new, the restart writer, the shim's dynamics binding.
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/memcheck_r02b.log \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_resident.py tests/test_dropin_resident.py -m gpu -x -q \
  -k "sigma_r_bit_exact and not 3000000 and not 2000000 or dynamics_apply_pairs and not large or simul_relax or restart_file or through_the_resident_shim and plummer" > $OUT/memcheck_r02b_pytest.log 2>&1
echo "rc=$?" >> $OUT/memcheck_r02b_pytest.log
tail -4 $OUT/memcheck_r02b_pytest.log; grep -c "Invalid\|Error" $OUT/memcheck_r02b.log; tail -4 $OUT/memcheck_r02b.log
