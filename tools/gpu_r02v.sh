#!/bin/bash
# Round 2c (final build of the round): the whole GPU test suite, then the one-GPU artefacts (tools/gpu_round_r02.sh r02c)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/gputest_r02c_1gpu.log; cat gpurun_out/gputest_r02c_1gpu.log
bash tools/gpu_round_r02.sh r02c
