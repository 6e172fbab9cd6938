#!/bin/bash
# gpurun call for the energy scheme's velocity correction: its parity tests, then its timing inside full steps
OUT=gpurun_out; TAG=${1:-sv3}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "velocity or energy_scheme" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log
tail -25 $OUT/pytest_$TAG.log
timeout 300 python tools/perf_energy.py 1e6 3 > $OUT/perf_energy_1e6_$TAG.log 2>&1; cat $OUT/perf_energy_1e6_$TAG.log
timeout 600 python tools/perf_energy.py 1e7 3 > $OUT/perf_energy_1e7_$TAG.log 2>&1; cat $OUT/perf_energy_1e7_$TAG.log
