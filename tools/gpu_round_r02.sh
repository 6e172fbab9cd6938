#!/bin/bash
# Round 2, one gpurun call on ONE GPU: bench line + reference arm, the other workloads, ncu launch list of the bench, one --set full capture, smoke.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_${TAG}_n1e7_1gpu.json 2> $OUT/bench_${TAG}_n1e7_1gpu.err
for W in king_w5_n1e6 king_w7_n2e6_bin plummer_n1e5; do
	timeout 600 python bench.py --steps 10 --warmup 3 --workload $W > $OUT/bench_${TAG}_$W.json 2> $OUT/bench_${TAG}_$W.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}_bench.csv python bench.py --steps 2 --warmup 1 --no-parity --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"orbit_prepare_kernel|orbit_try_kernel|sort_gather_records|sort_onesweep_pass|seq_scan_classify|seq_scan_finalize" -c 24 -f -o $OUT/prof_$TAG python tools/perf_orbit.py 1e7 1 > $OUT/ncu_full_$TAG.log 2>&1
# second half of round 2: the N3 kernels and a whole run
timeout 300 python tools/perf_sigma.py 1e7 imf > $OUT/perf_sigma_$TAG.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sigma|dyn_|simul_relax" -c 16 --csv --log-file $OUT/ncu_sigma_launches_$TAG.csv python tools/perf_sigma.py 1e7 imf > /dev/null 2>&1
timeout 200 python tools/full_run.py 1e6 1.40 > $OUT/full_run_${TAG}_n1e6.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log
python - <<PY
import json
for f in ("bench_${TAG}_reference", "bench_${TAG}_n1e7_1gpu", "bench_${TAG}_king_w5_n1e6", "bench_${TAG}_king_w7_n2e6_bin", "bench_${TAG}_plummer_n1e5"):
    try:
        d = json.loads([l for l in open("$OUT/" + f + ".json").read().splitlines() if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("value"), "parity", d.get("parity", {}).get("ok"))
    except Exception as e:
        print(f, "FAILED", e)
PY
