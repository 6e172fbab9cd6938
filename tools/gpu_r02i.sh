#!/bin/bash
# 8 GPUs: the sharded bench line (parity check inside), topology
nvidia-smi topo -m > gpurun_out/topo8.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_r02_8gpu.err
tail -c 400 gpurun_out/bench_r02_8gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_r02_8gpu.json").read().splitlines() if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "parity")}); print({k: v for k, v in d["stage_ms_per_step_max_over_ranks"].items() if v > 0}); print({k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step")}); print(d["cpu_baseline"])
PY
