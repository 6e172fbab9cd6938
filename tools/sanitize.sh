#!/bin/bash
# compute-sanitizer memcheck over the small GPU parity cases (one gpurun call); the log goes to gpurun_out/
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/memcheck.log \
  python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q \
  -k "plummer-2000 or plummer-1000 or potential_edge or sort_bit_exact and not 1000003 and not 100000 or cosmic_plummer_n100 or lagrange_plummer_n4000 or sort_ties" > $OUT/memcheck_pytest.log 2>&1
echo "rc=$?" >> $OUT/memcheck_pytest.log
tail -5 $OUT/memcheck_pytest.log; grep -c "Invalid\|Error" $OUT/memcheck.log; tail -5 $OUT/memcheck.log
