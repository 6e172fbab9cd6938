#!/bin/bash
# compute-sanitizer memcheck over what round 2c changed: the radix passes with vote-based ranking (n = 1 ... 100000, ties, all stripped) and
# the step whose orbit kernels write the sort keys (new stars, three steps, both knob settings)
OUT=gpurun_out; mkdir -p $OUT
timeout 230 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file $OUT/memcheck_r02c.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "sort_bit_exact and not 1000003 or sort_ties or step_with_sort_keys" > $OUT/memcheck_r02c_pytest.log 2>&1
echo "rc=$?" >> $OUT/memcheck_r02c_pytest.log
tail -4 $OUT/memcheck_r02c_pytest.log; grep -c "Invalid\|Error" $OUT/memcheck_r02c.log; tail -3 $OUT/memcheck_r02c.log
