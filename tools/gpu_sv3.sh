#!/bin/bash
# gpurun call after the energy scheme's velocity correction went in: the whole GPU suite, smoke, the timing of the energy passes inside
# full steps, one ncu --set full capture of their kernels, and the bench line of the unchanged hot path from the rebuilt library.
OUT=gpurun_out; TAG=${1:-sv3}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -4 $OUT/pytest_gpu_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
timeout 300 python tools/perf_energy.py 1e6 3 > $OUT/perf_energy_1e6_$TAG.log 2>&1
timeout 600 python tools/perf_energy.py 1e7 3 > $OUT/perf_energy_1e7_$TAG.log 2>&1; cat $OUT/perf_energy_1e7_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sv3_|energy_rnew_kernel" -c 12 -f -o $OUT/prof_energy_$TAG python tools/perf_energy.py 1e7 1 > $OUT/ncu_energy_$TAG.log 2>&1; tail -2 $OUT/ncu_energy_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json | cut -c1-700
