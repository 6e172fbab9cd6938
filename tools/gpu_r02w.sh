#!/bin/bash
# the whole GPU test suite with the final build of round 2c
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/gputest_r02c_1gpu.log; cat gpurun_out/gputest_r02c_1gpu.log
