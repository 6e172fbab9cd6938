#!/bin/bash
# round 2: new parity tests, then the whole GPU suite with its log kept
timeout 1500 python -m pytest tests/test_gpu_fullsize_parity.py tests/test_dropin_resident.py tests/test_gpu_parity.py -x -q -k "full_size or resident_shim or find_root_vr" 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/gputest_r02_1gpu.log
