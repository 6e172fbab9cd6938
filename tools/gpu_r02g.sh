#!/bin/bash
# 2 GPUs: the multi-GPU parity tests (both forms of the sharded sort, the multi-rank drop-in) with their log kept, then the sharded bench line
nvidia-smi topo -m > gpurun_out/topo2.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_sharded.py tests/test_dropin_shim.py -q 2>&1 | tail -15 | tee gpurun_out/gputest_r02_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err
tail -c 600 gpurun_out/bench_r02_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_2gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "parity", "stage_ms_per_step_max_over_ranks")}); print(d["e2e"])
PY
CMCB200_SHARD_HOST_BARRIERS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_2gpu_hostbar.json 2> gpurun_out/bench_r02_2gpu_hostbar.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_2gpu_hostbar.json"))
print("host barriers:", {k: d[k] for k in ("value", "ms_per_step")})
PY
