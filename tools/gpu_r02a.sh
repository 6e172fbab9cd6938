#!/bin/bash
# round 2, first call: resident-timestep tests, then the bench line
timeout 600 python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py -x -q -k "resident or host_step or tidally or zeros" 2>&1 | tail -15
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
tail -c 1500 gpurun_out/bench_r02a.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02a.json"))
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "parity", "cpu_baseline", "stage_ms_per_step")})
PY
