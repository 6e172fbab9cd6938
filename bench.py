#!/usr/bin/env python
"""Throughput of the sort + potential + new-orbit path: stars*timesteps/s (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
    python bench.py --impl reference --gpus N ...            the reference's own CPU implementation of the path on all host cores

One "step" is one pass of the hot path over one cluster state: new-orbit step for every star -> r <- rnew ->
radius sort (keys, 424-byte record gather, binaries) -> potential rebuild -> E, J from the new potential.
`value` times it with the arrays resident in HBM (CUDA events on the launching stream); `e2e` times the same work
through the host-buffer C-ABI entry points (what the reference-side shim calls) including every host<->device copy.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "stars*timesteps/s (radius sort + potential rebuild + new-orbit step)"
UNIT = "stars*steps/s"
# Algorithmic HBM bytes per star (restated in DESIGN.md "Kernels and their rooflines").
BYTES_GATHER = 848.0          # sort_gather_records: 424 B read + 424 B written
BYTES_STEP = 996.0            # whole step
# orbit_prepare_kernel (the dominant kernel): the streaming share of 8d's 96 B/star that stage A moves (E, J read; star_r, star_m of the
# star itself read; r_peri, r_apo and the envelope written) -- the bisection probes are NOT streaming traffic and are not counted here;
# they show up in `traffic` (32-byte sectors fetched by ~2 x 24 dependent probes per star once the 16 B/star table exceeds L2).
BYTES_ORBIT_PREPARE = 56.0


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_traffic(kernel):
    """dram bytes per star of a kernel from the committed ncu capture (profiles/traffic_r01.json), if there is one"""
    p = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_star")
        except Exception:
            return None
    return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, reasons, smax = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}
        return out


def workload(N, seed=1234):
    from cmc_cosmic_b200 import models
    r, vr, vt, m = models.plummer(N, seed=seed)
    return r, vr, vt, np.ones(N)          # equal masses: star_m = obj_m * N_STAR = 1


# ---------------------------------------------------------------------------------------------- reference / CPU legs


def run_reference_cpu(N, steps, warmup, procs):
    """The reference's own sources (oracle/_ref) as `procs` forked ranks over the mini-MPI shim; TIMER-style brackets
    (barrier, wall clock) around new_orbits_calculate, qsorts_new, post_sort_comm, calc_potential_new.
    Returns (stars*steps/s, per-phase seconds) on rank 0; other ranks never return."""
    import refharness as rh
    r, vr, vt, sm = workload(N)
    R = rh.Ref()
    win = int(3.0 * (N / procs + 4096) * 424) + (1 << 22)
    if R.L.ref_spawn(procs, win) != 0:
        raise RuntimeError("mini-MPI spawn failed")
    rank = R.L.ref_rank()
    R.init(N, n_star_dim=int(2.5 * (N // procs + 4096)) if procs > 1 else N + 4096, n_table_dim=N + 4096, min_chunk=40)
    S = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start()))
    E = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End()))
    lo, hi = int(S[rank]), int(E[rank])
    nl = hi - lo + 1
    sr, smm = R.star_r(), R.star_m()
    sr[0] = 1e-20; sr[1:N + 1] = r; sr[N + 1] = 1e10; smm[1:N + 1] = sm
    for name, val in (("r", r), ("vr", vr), ("vt", vt), ("m", sm)):
        R.star_field(name)[1:nl + 1] = val[lo - 1:hi]
    R.star_field("id", np.int64)[1:nl + 1] = np.arange(lo, hi + 1)
    R.set_scalars(N, nl)
    R.L.ref_calc_potential_new()
    R.L.ref_compute_energy()
    t = [C.c_double() for _ in range(4)]
    tot, phases, stars = 0.0, np.zeros(4), 0
    for step in range(warmup + steps):
        n_now = int(R.scalars()["N_MAX"])
        R.L.ref_time_step(*[C.byref(x) for x in t])
        R.L.ref_compute_energy()
        if step >= warmup:
            ph = np.array([x.value for x in t])
            phases += ph; tot += ph.sum(); stars += n_now
    if R.L.ref_join() != 0 and rank == 0:
        raise RuntimeError("a reference rank failed")
    return stars / tot, phases / max(steps, 1)


def cpu_baseline_leg(N_sample, kind_pref="reference"):
    """bounded single-core sample for the `cpu_baseline` object of the GPU arm (in-process, no fork: CUDA is live)"""
    import refharness as rh
    if kind_pref == "reference" and rh.available():
        r, vr, vt, sm = workload(N_sample, seed=99)
        R = rh.Ref()
        R.init(N_sample, n_star_dim=N_sample + 4096)
        R.load_cluster(r, vr, vt, sm)
        R.L.ref_calc_potential_new()
        R.L.ref_compute_energy()
        t = [C.c_double() for _ in range(4)]
        best = None
        for _ in range(2):
            R.L.ref_time_step(*[C.byref(x) for x in t])
            R.L.ref_compute_energy()
            tot = sum(x.value for x in t)
            best = tot if best is None else min(best, tot)
        R.free()
        return {"value": N_sample / best, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"plummer N={N_sample}, 1 rank of the compiled reference (oracle/_ref), best of 2 steps, TIMER-style brackets"}
    # fall back to the restatement
    import common as cm
    r, vr, vt, sm = cm.cluster_state("plummer", N_sample, seed=99)
    t0 = time.perf_counter()
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N_sample + 1], r, vr, vt)
    out = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N_sample, N_sample, E, J, vr, vt, pot["Rtidal"])
    order = np.argsort(out["rnew"], kind="stable")
    cm.oracle_potential(out["rnew"][order], sm[order])
    dt = time.perf_counter() - t0
    return {"value": N_sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"plummer N={N_sample}, oracle restatement (keys-only sort), 1 step"}


# ---------------------------------------------------------------------------------------------- the CUDA arm


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from cmc_cosmic_b200.cluster import Cluster
    from cmc_cosmic_b200.sharded import ShardedCluster

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line (NCCL's version banner goes to stderr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N = args.n_stars
    dev = f"cuda:{local_rank}"

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream()
    r, vr, vt, sm = workload(N)           # every rank builds the same cluster (same seed) and keeps its radius shard
    with torch.cuda.stream(stream):
        if world == 1:
            c = Cluster(N, n_star_dim=N + 4096, device=local_rank, pinned=True, stream=stream.cuda_stream)
            c.load(r, vr, vt, sm)
        else:
            c = ShardedCluster(N, dist, device=local_rank, pinned=True, stream=stream)
            c.load_global(r, vr, vt, sm)
        c.calc_potential_new()
        c.upload()
        c.sync()
        clocks = ClockSampler(local_rank) if rank == 0 else None     # sampled under load: warm-up and timed steps
        # ---- resident arm: `value` ----
        for _ in range(args.warmup):
            c.dev_step(seed=args.seed)
        c.sync()
        c.profile(True)
        c.profile_read(reset=True)
        l0 = c.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier_sync()
        stars = 0
        e0.record(stream)
        for _ in range(args.steps):
            stars += c.clus.N_MAX                     # the whole cluster advances one step (all ranks together)
            c.dev_step(seed=args.seed)
        e1.record(stream)
        barrier_sync()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = c.launch_count() - l0
        prof = c.profile_read(reset=True)
        c.profile(False)
        value = stars / (ms * 1e-3)
        n_now, n_shard = c.clus.N_MAX, c.clus.N_MAX_NEW

        # ---- auxiliary: the same step inside the reference's timestep order WITH the energy-conservation scheme and the tidal stripping
        # around it (Cluster.dev_timestep), resident, CUDA events.  Not the metric: reported next to it.
        timestep_ms = None
        if world == 1:
            try:
                c.dev_timestep(seed=args.seed, tidally_strip=1)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c.sync()
                t0.record(stream)
                for _ in range(3):
                    c.dev_timestep(seed=args.seed, tidally_strip=1)
                t1.record(stream)
                c.sync()
                timestep_ms = t0.elapsed_time(t1) / 3
            except Exception as exc:            # never lose the bench line over the auxiliary number
                print(f"bench: auxiliary timestep measurement skipped: {exc}", file=sys.stderr)

        # ---- e2e arm: the same step through the public API with HOST buffers, copies included ----
        c.download()
        c.sync()
        e2e_t, e2e_stars, h2d, d2h = 0.0, 0, 0, 0
        e2e_calls = {}
        if world == 1:
            # the drop-in form: the four host-buffer C-ABI entry points the reference-side shim binds
            for it in range(max(1, min(args.steps, 3)) + 1):
                n = c.clus.N_MAX_NEW
                c.compute_energy()
                if it == 1:
                    c.transfer_bytes(reset=True)      # the library counts what its host-buffer entry points move
                    c.profile(True)                   # stage timers (CUDA events) split the calls into copies and kernels
                    c.profile_read(reset=True)
                t0 = time.perf_counter()
                c.new_orbits_calculate(seed=args.seed + it, want_status=False)
                t1 = time.perf_counter()
                # host code of the caller between the calls (ComputeIntermediateEnergy's transfer, reference cmc_utils.c:380-390): untimed
                for dst, src in (("r", "rnew"), ("vr", "vrnew"), ("vt", "vtnew")):
                    c.field(dst)[1:n + 1] = c.field(src)[1:n + 1]
                t2 = time.perf_counter()
                c.qsorts_new()
                c.post_sort_comm()
                t2b = time.perf_counter()
                c.calc_potential_new()
                t3 = time.perf_counter()
                if it > 0:  # the first pass is warm-up (page faults, pinning)
                    e2e_t += (t1 - t0) + (t3 - t2)
                    e2e_stars += n
                    n_e2e = it
                    for k, v in (("new_orbits_calculate", t1 - t0), ("qsorts_new+post_sort_comm", t2b - t2), ("calc_potential_new", t3 - t2b)):
                        e2e_calls[k] = e2e_calls.get(k, 0.0) + v
            h2d, d2h = (x / n_e2e for x in c.transfer_bytes())
            e2e_stage = {k: round(v[0] / n_e2e, 2) for k, v in c.profile_read(reset=True).items() if v[0] > 0}
            c.profile(False)
            note = ("new_orbits_calculate + qsorts_new + post_sort_comm + calc_potential_new through the host-buffer C ABI, pinned host arrays "
                    "(orbit step: 128 B in / 64 B out per star as strided DMA pipelined with the kernels; sort: whole 424 B records both ways), "
                    "wall clock; the caller's r<-rnew transfer between the calls is host code of the reference and is not timed")
        else:
            # sharded: every rank uploads its shard and the tables from pinned host memory, steps, and downloads them again
            for it in range(max(1, min(args.steps, 3)) + 1):
                barrier_sync()
                t0 = time.perf_counter()
                c.upload()
                c.dev_step(seed=args.seed + it)
                c.download()
                c.sync()
                barrier_sync()
                dt = max_over_ranks(time.perf_counter() - t0)
                if it > 0:
                    e2e_t += dt
                    e2e_stars += c.clus.N_MAX
                    nt = c.clus.N_MAX + 2
                    h2d = d2h = c.clus.N_MAX_NEW * 424 + 3 * 8 * nt
            note = "per rank: upload(shard of star[] + tables) -> dev_step -> download, pinned host arrays, wall clock, max over ranks; bytes are per rank"
        e2e = {"value": e2e_stars / e2e_t, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "note": note}
        if e2e_calls:
            e2e["ms_per_call"] = {k: round(1e3 * v / n_e2e, 2) for k, v in e2e_calls.items()}
            e2e["device_ms_per_step"] = e2e_stage
        clk = clocks.stop() if clocks else None
        c.close()

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()

    def roof(stage, kernel, bytes_per_star, n_units, extra=None):
        t_ms, n_l = prof[stage]
        avg = t_ms / max(n_l, 1)
        ach = bytes_per_star * n_units / (avg * 1e-3) / 1e9 if avg > 0 else None
        tr = profile_traffic(kernel)
        d = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
             "traffic": (tr * n_units if tr else None), "peak_source": peak_src, "algorithmic_bytes_per_star": bytes_per_star, "avg_launch_ms": avg,
             # the DRAM bytes ncu counted for this kernel (profiles/traffic_r01.json) over the launch time measured here, against the same peak:
             # how close the kernel is to the memory system's limit on the traffic it actually causes
             "traffic_frac": (tr * n_units / (avg * 1e-3) / 1e9 / peak) if (tr and avg > 0) else None}
        d.update(extra or {})
        return d

    roofs = {
        "orbit_prepare": roof("orbit_prepare", "orbit_prepare_kernel", BYTES_ORBIT_PREPARE, n_shard,
                              {"note": "2 x log2(N) dependent probes of the 16 N-byte (r, Phi) table per star: bound by DRAM sector traffic (`traffic`, ncu: ~49 % of "
                                       "peak DRAM throughput in 32-byte random sectors, L2 hit rate 38 %), not by the 56 B/star it streams"}),
        "record_gather": roof("record_gather", "sort_gather_records" if world == 1 else "shard_push_records", BYTES_GATHER, n_shard),
    }
    dominant = max(roofs, key=lambda k: roofs[k]["avg_launch_ms"])
    stage_ms = {k: round(v[0] / max(args.steps, 1), 4) for k, v in prof.items()}
    cpu = cpu_baseline_leg(args.cpu_sample) if world == 1 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (Plummer sphere, equal masses, Henon units; numpy generator in cmc_cosmic_b200/models.py)",
        "config": {"workload": f"plummer_n{N:.0e}".replace("+0", "").replace("+", ""), "n_stars": N, "record_bytes": 424,
                   "rng": "philox4x32-10 per star (production mode)", "l2": "inputs larger than L2 (star array %.2f GB, r/Phi table %.0f MB)" % (N * 424 / 1e9, N * 16 / 1e6),
                   "partition": "1 GPU" if world == 1 else f"radius-sharded over {world} GPUs (findLimits, MIN_CHUNK_SIZE=40); records exchanged by peer stores over NVLink, "
                                                           "tables replicated, potential rebuilt on every rank",
                   "step": "E,J from Phi + orbits (index bisections) -> rejection rounds -> radix sort on rnew + 424 B record gather with r<-rnew -> "
                           "potential (exact sequential rounding)"},
        "clocks": clk,
        "e2e": e2e,
        "gpu_launches": int(launches),
        # the dominant kernel by time (largest average launch among the single-kernel stages) ...
        "roofline": roofs[dominant],
        # ... and both candidates: the 424-byte record gather (HBM streaming) and stage A of the orbit step (index bisections: its probes are
        # 32-byte sector reads, so its streaming-bytes fraction is small by construction; `traffic` (ncu, profiles/) shows what bounds it)
        "roofline_streaming": roofs["record_gather"],
        "roofline_orbit": roofs["orbit_prepare"],
        "step_bytes_per_star": BYTES_STEP, "step_achieved_gbs": BYTES_STEP * n_now / (ms / args.steps * 1e-3) / 1e9,
        "stage_ms_per_step": stage_ms,
        "cpu_baseline": cpu,
        "timestep_with_energy_scheme_ms": timestep_ms,
    }
    if world > 1:
        line["stage_ms_per_step_rank"] = 0
        dist.destroy_process_group()
    print(json.dumps(line))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import refharness as rh
    if not rh.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcmc_ref.so was not built (needs /root/reference at build time)"}))
        return
    cores = len(os.sched_getaffinity(0))
    procs = min(cores, 64)
    # bounded sample: the whole workload if ~10 s per step suffices (2.5e5 stars/s/core measured), else proportionally fewer stars
    N = args.ref_sample or int(min(args.n_stars, 2.5e5 * procs * 0.5 * 10.0))
    val, phases = run_reference_cpu(N, args.steps, min(args.warmup, 1), procs)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": float(phases.sum() * 1e3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (same Plummer generator as the CUDA arm)",
        "config": {"workload": f"plummer_n{args.n_stars:.0e}".replace("+0", "").replace("+", ""), "sample_n_stars": N,
                   "phases_s": {"t_orb": phases[0], "t_sort": phases[1], "t_postsort_comm": phases[2], "t_pot_cal": phases[3]}},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "reference",
                         "sample": f"plummer N={N} per step (bounded sample of the N={args.n_stars} workload), {procs} forked ranks of the compiled "
                                   "reference over a shared-memory mini-MPI, TIMER-style brackets"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--n-stars", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--ref-sample", type=int, default=0, help="stars per step of the reference arm (0: derive from the core count)")
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
