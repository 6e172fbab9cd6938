#!/bin/bash
# One gpurun call: GPU tests, stage timings for the tuning knobs, the bench line, the ncu launch list and one --set full capture.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
timeout 600 python tools/perf_orbit.py 1e7 3 "" "CMCB200_ORBIT_LADDERS=0" "CMCB200_ORBIT_LADDERS=3" > $OUT/perf_1e7_$TAG.log 2>&1
timeout 300 python tools/perf_orbit.py 1e6 3 "" "CMCB200_ORBIT_LADDERS=0" > $OUT/perf_1e6_$TAG.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv python tools/perf_orbit.py 1e7 2 > $OUT/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"orbit_prepare_kernel|orbit_try_kernel|sort_gather_records|sort_onesweep_pass|seq_scan_classify|seq_scan_finalize" -c 24 -f -o $OUT/prof_$TAG python tools/perf_orbit.py 1e7 1 > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log; cat $OUT/perf_1e7_$TAG.log; cat $OUT/bench_$TAG.json
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_$TAG.json 2> $OUT/bench_reference_$TAG.err; cat $OUT/bench_reference_$TAG.json | cut -c1-600
