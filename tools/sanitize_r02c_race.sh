#!/bin/bash
# compute-sanitizer racecheck over the shared-memory phases of the radix pass with the vote-based ranking (counter chain, scatter, write-out)
OUT=gpurun_out; mkdir -p $OUT
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file $OUT/racecheck_r02c.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort_bit_exact and not 1000003 or sort_ties" > $OUT/racecheck_r02c_pytest.log 2>&1
echo "rc=$?" >> $OUT/racecheck_r02c_pytest.log
tail -3 $OUT/racecheck_r02c_pytest.log; tail -3 $OUT/racecheck_r02c.log
