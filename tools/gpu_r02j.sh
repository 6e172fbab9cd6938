#!/bin/bash
# try rounds: sector-sparing ladder lines + compact job arrays -- stage timers, then the whole GPU suite
python tools/perf_orbit.py 1e7 3 2>&1 | tee gpurun_out/tries_r02_n1e7.log | cut -c1-420
python tools/perf_orbit.py 1e6 3 2>&1 | tee gpurun_out/tries_r02_n1e6.log | cut -c1-420
python tools/perf_energy.py 1e7 2>&1 | tail -12 | tee gpurun_out/perf_energy_r02_n1e7.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/gputest_r02_1gpu.log
