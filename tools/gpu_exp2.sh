#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_exp2.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_exp2.log; tail -3 $OUT/pytest_gpu_exp2.log
timeout 600 python tools/perf_orbit.py 1e7 2 "" "CMCB200_ORBIT_LADDERS=0" > $OUT/exp2.log 2>&1
grep "step 1" $OUT/exp2.log | sed 's/status=.*| / | /' | cut -c1-330
timeout 300 python tools/perf_orbit.py 1e6 2 "" "CMCB200_ORBIT_LADDERS=0" > $OUT/exp2_1e6.log 2>&1
grep "step 1" $OUT/exp2_1e6.log | sed 's/status=.*| / | /' | cut -c1-330
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_exp2.json 2> $OUT/bench_exp2.err; cat $OUT/bench_exp2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"orbit_prepare_kernel|orbit_try_kernel" -c 2 -f -o $OUT/prof_exp2 python tools/perf_orbit.py 1e7 1 > $OUT/ncu_exp2.log 2>&1
