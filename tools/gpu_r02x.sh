#!/bin/bash
# 2 GPUs, final build of round 2c: the multi-GPU parity tests (log kept), then the sharded bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_dropin_shim.py -q 2>&1 | tail -6 | tee gpurun_out/gputest_r02c_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02c_n1e7_2gpu.json 2> gpurun_out/bench_r02c_n1e7_2gpu.err
tail -c 400 gpurun_out/bench_r02c_n1e7_2gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_r02c_n1e7_2gpu.json").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "parity", "stage_ms_per_step_max_over_ranks")}); print(d["e2e"]["ms_per_step"])
PY
