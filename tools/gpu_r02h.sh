#!/bin/bash
# 4 GPUs: the multi-GPU parity tests incl. the 4-rank cases, then the sharded bench line
timeout 1500 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -6 | tee gpurun_out/gputest_r02_4gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_r02_4gpu.json 2> gpurun_out/bench_r02_4gpu.err
tail -c 400 gpurun_out/bench_r02_4gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_r02_4gpu.json").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "parity")}); print({k: v for k, v in d["stage_ms_per_step_max_over_ranks"].items() if v > 0}); print({k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step")}); print(d["cpu_baseline"])
PY
