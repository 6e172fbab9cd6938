#!/usr/bin/env python
"""Throughput of the sort + potential + new-orbit path: stars*timesteps/s (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
    python bench.py --impl reference --gpus N ...            the reference's own CPU implementation of the path on all host cores

One "step" is one pass of the hot path over one cluster state: new-orbit step for every star -> r <- rnew ->
radius sort (keys, 424-byte record gather, binaries) -> potential rebuild -> E, J from the new potential.
`value` times it with the arrays resident in HBM (CUDA events on the launching stream); `e2e` times the same work on
HOST arrays through the C ABI the resident mode of the reference-side shim calls (cmcb200_resident_begin -> cmcb200_dev_* ->
cmcb200_resident_end: every record crosses the host link once each way per timestep), wall clock, all copies inside.
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
    """dram bytes per star of a kernel from the committed ncu capture (profiles/traffic_r02.json, else the round-1 file), if there is one"""
    for name in ("traffic_r02c.json", "traffic_r02b.json", "traffic_r02.json", "traffic_r01.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                v = json.load(open(p)).get(kernel, {}).get("dram_bytes_per_star")
                if v:
                    return v
            except Exception:
                pass
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


# BASELINE.json `configs`: [1] Plummer 1e5, [2] King W0=5 1e6, [3] Plummer 1e7 (the configuration the metric is quoted on; default),
# [4] King W0=7 2e6 with IMF masses and 10 % primordial binaries (464-byte binary_t records follow their stars)
WORKLOADS = {
    "plummer_n1e7": dict(kind="plummer", N=10_000_000, seed=1234),
    "plummer_n1e5": dict(kind="plummer", N=100_000, seed=1234),
    "king_w5_n1e6": dict(kind="king", w0=5.0, N=1_000_000, seed=12345),
    "king_w7_n2e6_bin": dict(kind="king", w0=7.0, N=2_000_000, seed=12345, imf=True, bin_frac=0.1),
}


def workload_name(args):
    w = WORKLOADS[args.workload]
    if args.n_stars and args.n_stars != w["N"]:
        return f"{args.workload.rsplit('_n', 1)[0]}_n{args.n_stars:.0e}".replace("+0", "").replace("+", "") + ("_bin" if w.get("bin_frac") else "")
    return args.workload


def workload(args, N=None, seed=None):
    """(r, vr, vt, star_m, binind or None, n_binaries): radius-sorted, star_m = obj_m * N_STAR"""
    from cmc_cosmic_b200 import models
    w = WORKLOADS[args.workload]
    N = int(N or args.n_stars or w["N"])
    seed = w["seed"] if seed is None else seed
    if w["kind"] == "plummer":
        r, vr, vt, m = models.plummer(N, seed=seed)
    else:
        r, vr, vt, m = models.king(N, w0=w["w0"], seed=seed)
    if w.get("imf"):
        m = models.kroupa_like_masses(N, seed=seed + 1)
    binind, nb = None, 0
    if w.get("bin_frac"):
        rng = np.random.default_rng(seed + 2)
        nb = int(w["bin_frac"] * N)
        binind = np.zeros(N, np.int64)
        binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    return r, vr, vt, m * N, binind, nb


def config_for(args, world, N):
    """the same dictionary on both arms (the driver compares them)"""
    return {"workload": workload_name(args), "n_stars": N, "record_bytes": 424,
            "partition": "1 rank per GPU" if world == 1 else f"radius-sharded over {world} GPUs (findLimits, MIN_CHUNK_SIZE=40)",
            "l2": ("inputs larger than L2" if N * 424 / world > 2 * 126e6 else "inputs NOT larger than L2 (a parity-size workload, not the headline)") +
                  " (star array %.2f GB, r/Phi table %.0f MB)" % (N * 424 / 1e9, N * 16 / 1e6),
            "step": "new_orbits_calculate -> r <- rnew -> qsorts_new + post_sort_comm -> calc_potential_new (E, J from the new potential between steps)"}


# ---------------------------------------------------------------------------------------------- reference / CPU legs


def run_reference_cpu(args, N, steps, warmup, procs):
    """The reference's own sources (oracle/_ref) as `procs` forked ranks over the mini-MPI shim; TIMER-style brackets
    (barrier, wall clock) around new_orbits_calculate, qsorts_new, post_sort_comm, calc_potential_new.
    Returns (stars*steps/s, per-phase seconds) on rank 0; other ranks never return."""
    import refharness as rh
    r, vr, vt, sm, binind, nb = workload(args, N)
    R = rh.Ref()
    win = int(3.0 * (N / procs + 4096) * 424) + (1 << 22)
    if R.L.ref_spawn(procs, win) != 0:
        raise RuntimeError("mini-MPI spawn failed")
    rank = R.L.ref_rank()
    R.init(N, n_star_dim=int(2.5 * (N // procs + 4096)) if procs > 1 else N + 4096, n_table_dim=N + 4096, min_chunk=40,
           n_bin_dim=(int(2.5 * (nb // procs + 1024)) if procs > 1 else nb + 64) if nb else None)
    S = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start()))
    E = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End()))
    lo, hi = int(S[rank]), int(E[rank])
    nl = hi - lo + 1
    sr, smm = R.star_r(), R.star_m()
    sr[0] = 1e-20; sr[1:N + 1] = r; sr[N + 1] = 1e10; smm[1:N + 1] = sm
    for name, val in (("r", r), ("vr", vr), ("vt", vt), ("m", sm)):
        R.star_field(name)[1:nl + 1] = val[lo - 1:hi]
    R.star_field("id", np.int64)[1:nl + 1] = np.arange(lo, hi + 1)
    nbl = 0
    if nb:      # this rank's binaries, renumbered 1.. in star order (what the reference's load-time distribution leaves)
        b = binind[lo - 1:hi]
        loc = np.zeros(nl, np.int64)
        nbl = int((b > 0).sum())
        loc[b > 0] = np.arange(1, nbl + 1)
        R.star_field("binind", np.int64)[1:nl + 1] = loc
        R.binary_bytes()[1:nbl + 1, R.binary_offset("inuse")] = 1
    R.set_scalars(N, nl, N_BINARY=nbl)
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


def cpu_baseline_single_core(args, N_sample):
    """bounded single-core sample, in-process (no fork: CUDA is live) -- the fall-back of cpu_baseline_all_cores"""
    import refharness as rh
    if not rh.available():
        return None
    r, vr, vt, sm, binind, nb = workload(args, N_sample, seed=99)
    R = rh.Ref()
    R.init(N_sample, n_star_dim=N_sample + 4096, n_bin_dim=nb + 64 if nb else None)
    R.load_cluster(r, vr, vt, sm, binind=binind)
    R.set_scalars(N_sample, N_sample, N_BINARY=nb)
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
            "sample": f"{workload_name(args)} cut to N={N_sample}, 1 rank of the compiled reference (oracle/_ref), best of 2 steps, TIMER-style brackets"}


def cpu_baseline_all_cores(args):
    """`cpu_baseline` of the CUDA arm's line: the reference arm itself (all host cores, forked ranks) on a bounded sample, run as a child
    process so that nothing is forked from the process that holds the CUDA context"""
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", args.workload,
               "--n-stars", str(args.n_stars), "--ref-budget-s", "25"]
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
        cb = line["cpu_baseline"]
        cb["ms_per_step"] = line["ms_per_step"]
        return cb
    except Exception as exc:
        print(f"bench: all-core cpu_baseline leg failed ({exc}); single-core sample instead", file=sys.stderr)
        return cpu_baseline_single_core(args, args.cpu_sample)


# ---------------------------------------------------------------------------------------------- the CUDA arm


def parity_check(world, rank, local_rank, dist, stream):
    """Before anything is timed: one step of a 2e5-star Plummer cluster with RECORDED draws through the path this run is about to time
    (1 GPU: Cluster; N GPUs: the radius-sharded ShardedCluster), every rank's records, the replicated tables and N_MAX compared bit
    for bit with the CPU oracle run as one rank (oracle/ as the checker only).  Returns {"checked", "ok", ...}."""
    import torch
    import common as cm
    from cmc_cosmic_b200.cluster import Cluster
    from cmc_cosmic_b200.sharded import ShardedCluster
    N = 200_000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=4321)
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vr, vt, pot["Rtidal"], seed=3)
    order = np.argsort(want["rnew"], kind="stable")
    n_new = int((want["rnew"] < cm.SF_INFINITY).sum())
    order = order[:n_new]
    pot2 = cm.oracle_potential(want["rnew"][order], sm[order], N_STAR=N)
    bad = []
    with torch.cuda.stream(stream):
        if world == 1:
            c = Cluster(N, n_star_dim=N + 64, device=local_rank, stream=stream.cuda_stream)
            c.load(r, vr, vt, sm)
            lo, hi = 1, N
            draws, off = want["draws"], want["draw_off"]
        else:
            c = ShardedCluster(N, dist, device=local_rank, stream=stream)
            c.load_global(r, vr, vt, sm)
            lo, hi = int(c.Start[rank]), int(c.End[rank])
            off = want["draw_off"][lo - 1:hi + 1] - want["draw_off"][lo - 1]
            draws = want["draws"][want["draw_off"][lo - 1]:want["draw_off"][hi]]
        c.calc_potential_new()
        if not np.array_equal(c.star_phi[:N + 2].view(np.uint64), pot["phi"][:N + 2].view(np.uint64)):
            bad.append("initial potential")
        c.upload()
        c.dev_step(draws=draws, draw_off=off)
        c.dev_compute_energy()
        c.download()
        c.sync()
        lo2, nl = int(c.Start[rank if world > 1 else 0]), c.clus.N_MAX_NEW
        if c.clus.N_MAX != n_new:
            bad.append(f"N_MAX {c.clus.N_MAX} != {n_new}")
        else:
            sl = order[lo2 - 1:lo2 - 1 + nl]
            for f, w in (("r", want["rnew"][sl]), ("vr", want["vrnew"][sl]), ("vt", want["vtnew"][sl]), ("m", sm[sl]), ("r_peri", want["r_peri"][sl]),
                         ("r_apo", want["r_apo"][sl])):
                if not np.array_equal(np.ascontiguousarray(c.field(f)[1:nl + 1]).view(np.uint64), np.ascontiguousarray(w).view(np.uint64)):
                    bad.append(f"star[].{f}")
            if not np.array_equal(c.field("id")[1:nl + 1], sl + 1):
                bad.append("sort permutation (star[].id)")
            for f, w in (("star_r", pot2["star_r"]), ("star_m", pot2["star_m"]), ("star_phi", pot2["phi"])):
                if not np.array_equal(getattr(c, f)[1:n_new + 1].view(np.uint64), w[1:n_new + 1].view(np.uint64)):
                    bad.append(f)
        c.close()
    ok = 0 if bad else 1
    if world > 1:
        t = torch.tensor([ok], dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = int(t.item())
    if bad:
        print(f"bench: parity check FAILED on rank {rank}: {bad}", file=sys.stderr)
    return {"checked": True, "ok": bool(ok), "n_stars": N, "ranks": world,
            "what": "one step with recorded draws: every rank's records (r, vr, vt, m, r_peri, r_apo, id), star_r / star_m / star_phi and N_MAX "
                    "bit-identical to the CPU oracle run as one rank"}


def wait_for_rank0(world, rank, tag, fn):
    """rank 0 runs fn() while the other ranks sleep (a collective would spin on their cores and disturb a CPU measurement)"""
    flag = os.path.join(tempfile.gettempdir(), f"cmcb200_bench_{os.environ.get('MASTER_PORT', '0')}_{tag}")
    if rank == 0:
        try:
            return fn()
        finally:
            open(flag, "w").write("done")
    while not os.path.exists(flag):
        time.sleep(0.25)
    return None


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
        for f in os.listdir(tempfile.gettempdir()):             # flags of an earlier run on the same port
            if rank == 0 and f.startswith(f"cmcb200_bench_{os.environ.get('MASTER_PORT', '0')}_"):
                os.unlink(os.path.join(tempfile.gettempdir(), f))
        dist.barrier()
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
    parity = parity_check(world, rank, local_rank, dist, stream) if not args.no_parity else {"checked": False}
    r, vr, vt, sm, binind, nb = workload(args)           # every rank builds the same cluster (same seed) and keeps its radius shard
    with torch.cuda.stream(stream):
        if world == 1:
            c = Cluster(N, n_star_dim=N + 4096, n_bin_dim=nb + 64 if nb else None, device=local_rank, pinned=True, stream=stream.cuda_stream)
            c.load(r, vr, vt, sm, binind=binind)
            if nb:
                c.bfield("inuse")[1:nb + 1] = 1
        else:
            c = ShardedCluster(N, dist, device=local_rank, pinned=True, stream=stream, n_bin_dim=int(2.5 * (nb // world + 1024)) if nb else None)
            c.load_global(r, vr, vt, sm, binind_global=binind)
            if nb:
                c.bfield("inuse")[1:c.N_b_local + 1] = 1
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
        # stage timers of ALL ranks: the maximum over the ranks per stage names the limiter of a sharded step
        stage_ms_max = {k: max_over_ranks(v[0] / max(args.steps, 1)) for k, v in prof.items()}

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
        # ---- auxiliary 2: the WHOLE timestep of src/cmc/cmc.c:286-353 resident -- dynamics_apply (calc_sigma_r + two-body relaxation of every
        # neighbouring pair) in front of the above.  With it nothing of a PlummerSphere.ini timestep reads the records on the host.
        timestep_dyn_ms, dyn_stage = None, None
        if world == 1:
            try:
                c.dev_timestep(seed=args.seed, tidally_strip=1, dynamics_dt=1e-3)
                c.profile(True)
                c.profile_read(reset=True)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c.sync()
                t0.record(stream)
                for _ in range(3):
                    c.dev_timestep(seed=args.seed, tidally_strip=1, dynamics_dt=1e-3)
                t1.record(stream)
                c.sync()
                timestep_dyn_ms = t0.elapsed_time(t1) / 3
                pr = c.profile_read(reset=True)
                c.profile(False)
                dyn_stage = {"calc_sigma_r_ms": round(pr["sigma"][0] / 3, 4), "dynamics_apply_pairs_ms": round(pr["dynamics"][0] / 3, 4),
                             "encounter_pairs": int(c.last_dynamics.n_encounters), "relaxed_pairs": int(c.last_dynamics.Nrel)}
            except Exception as exc:
                print(f"bench: auxiliary timestep-with-dynamics measurement skipped: {exc}", file=sys.stderr)

        # ---- e2e arm: the same step on HOST arrays through the C ABI of the shim's resident mode, every copy inside the timed region.
        # Per step and rank: cmcb200_resident_begin (tables + this rank's records + binaries up, asynchronous) -> orbit step chunk by chunk
        # behind the arriving records -> r <- rnew + sort -> potential -> cmcb200_resident_end (records, binaries, tables down).  The same
        # definition at every N: each rank moves its own shard over its own host link.  E, J are the caller's, as in the reference: the
        # host computes them from the new potential between the steps (untimed on both arms).
        c.download()
        c.sync()
        e2e_t, e2e_stars, n_e2e = 0.0, 0, 0
        e2e_stage = {}
        n_rep = max(2, min(args.steps, 4))
        for it in range(n_rep + 1):
            c.compute_energy()
            if it == 1:     # the first pass is warm-up (page faults, pinning)
                c.transfer_bytes(reset=True)
                c.profile(True)
                c.profile_read(reset=True)
            barrier_sync()
            t0 = time.perf_counter()
            n_before = c.clus.N_MAX
            c.host_step(seed=args.seed + it)
            dt = time.perf_counter() - t0
            barrier_sync()
            dt = max_over_ranks(dt)
            if it > 0:
                e2e_t += dt
                e2e_stars += n_before
                n_e2e += 1
        h2d, d2h = (x / n_e2e for x in c.transfer_bytes())
        e2e_stage = {k: round(v[0] / n_e2e, 2) for k, v in c.profile_read(reset=True).items() if v[0] > 0}
        c.profile(False)
        e2e = {"value": e2e_stars / e2e_t, "unit": UNIT, "ms_per_step": 1e3 * e2e_t / n_e2e, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "device_ms_per_step": e2e_stage,
               "note": "per rank: cmcb200_resident_begin(records of the shard + binaries + star_r/m/phi, pinned host arrays, asynchronous) -> "
                       "cmcb200_dev_new_orbits chunk by chunk behind the upload -> cmcb200_dev_transfer_and_sort -> cmcb200_dev_calc_potential_new -> "
                       "cmcb200_resident_end(records + binaries + tables); wall clock, max over ranks; bytes are per rank, counted by the library"}
        # the same with the caller's declaration of what it wrote since the last step (here: E and J, which is all the host code between two
        # steps of this benchmark -- and of the reference arm -- touches).  Auxiliary: the headline e2e sends every byte.
        try:
            from cmc_cosmic_b200.cluster import Cluster as _C
            t_d, n_d = 0.0, 0
            for it in range(3):
                c.compute_energy()
                barrier_sync()
                t0 = time.perf_counter()
                c.host_step(seed=args.seed + 10 + it, dirty_words=_C.dirty_mask("E", "J"))
                dt = max_over_ranks(time.perf_counter() - t0)
                barrier_sync()
                if it > 0:
                    t_d += dt; n_d += 1
            e2e["dirty_words_E_J_only_ms_per_step"] = 1e3 * t_d / n_d
        except Exception as exc:
            print(f"bench: dirty-word e2e variant skipped: {exc}", file=sys.stderr)
        clk = clocks.stop() if clocks else None
        c.close()

    if world > 1:
        dist.barrier()
    cpu = None
    if not args.no_cpu_baseline:
        cpu = wait_for_rank0(world, rank, "cpu", lambda: cpu_baseline_all_cores(args)) if world > 1 else cpu_baseline_all_cores(args)
    if rank != 0:
        if world > 1:
            dist.barrier()
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
             # the DRAM bytes ncu counted for this kernel (profiles/traffic_r0x.json) over the launch time measured here, against the same peak:
             # how close the kernel is to the memory system's limit on the traffic it actually causes
             "traffic_frac": (tr * n_units / (avg * 1e-3) / 1e9 / peak) if (tr and avg > 0) else None}
        d.update(extra or {})
        return d

    roofs = {
        "orbit_prepare": roof("orbit_prepare", "orbit_prepare_kernel", BYTES_ORBIT_PREPARE, n_shard,
                              {"note": "2 x log2(N) dependent probes of the (r, Phi) table per star: bound by DRAM / L2 sector traffic (`traffic`, ncu), "
                                       "not by the 56 B/star it streams"}),
        "record_gather": roof("record_gather", "sort_gather_records" if world == 1 else "shard_push_records", BYTES_GATHER, n_shard),
    }
    dominant = max(roofs, key=lambda k: roofs[k]["avg_launch_ms"])
    stage_ms = {k: round(v[0] / max(args.steps, 1), 4) for k, v in prof.items()}
    step_gbs = BYTES_STEP * n_now / (ms / args.steps * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (numpy generators in cmc_cosmic_b200/models.py: Plummer / King spheres in Henon units, equal or power-law masses; "
                "not COSMIC's elson.py / king.py, which cannot travel to the GPU box -- the same arrays feed both arms)",
        "config": config_for(args, world, N),
        "rng": "philox4x32-10 per star (production mode); parity runs use recorded draws",
        "clocks": clk,
        "e2e": e2e,
        "parity": parity,
        "gpu_launches": int(launches),
        # the dominant kernel by time (largest average launch among the single-kernel stages) ...
        "roofline": roofs[dominant],
        # ... and both candidates: the 424-byte record gather (HBM streaming) and stage A of the orbit step (index bisections: its probes are
        # 32-byte sector reads, so its streaming-bytes fraction is small by construction; `traffic` (ncu, profiles/) shows what bounds it)
        "roofline_streaming": roofs["record_gather"],
        "roofline_orbit": roofs["orbit_prepare"],
        # the whole step against the same peak: SURVEY 8d's 996 algorithmic bytes per star and step
        "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak / world, "algorithmic_bytes_per_star": BYTES_STEP,
                          "note": "996 B/star x N / step time, per GPU"},
        "step_bytes_per_star": BYTES_STEP, "step_achieved_gbs": step_gbs,
        "stage_ms_per_step": stage_ms,
        "stage_ms_per_step_max_over_ranks": {k: round(v, 4) for k, v in stage_ms_max.items()},
        "cpu_baseline": cpu,
        "timestep_with_energy_scheme_ms": timestep_ms,
        "timestep_with_dynamics_ms": timestep_dyn_ms,
        "dynamics_stage": dyn_stage,
    }
    if world > 1:
        dist.barrier()
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
    # bounded sample: the whole workload if the run fits the time budget (2.5e5 stars/s/core measured, half of it kept as margin), else
    # proportionally fewer stars per step
    n_steps = args.steps + args.warmup
    N = args.ref_sample or int(min(args.n_stars, 2.5e5 * procs * 0.5 * args.ref_budget_s / max(n_steps, 1)))
    val, phases = run_reference_cpu(args, N, args.steps, args.warmup, procs)
    cfg = config_for(args, args.gpus, args.n_stars)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(phases.sum() * 1e3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (the CUDA arm's generators, same seeds)",
        "config": cfg,
        "rng": "taus113 (the reference's own generator)",
        "reference_detail": {"sample_n_stars": N, "ranks": procs,
                             "phases_s": {"t_orb": phases[0], "t_sort": phases[1], "t_postsort_comm": phases[2], "t_pot_cal": phases[3]}},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "reference",
                         "sample": f"{workload_name(args)}: N={N} stars per step" + ("" if N == args.n_stars else f" (bounded sample of the N={args.n_stars} workload)") +
                                   f", {procs} forked ranks of the compiled reference over a shared-memory mini-MPI, TIMER-style brackets, {args.steps} steps after {args.warmup} warm-up"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="plummer_n1e7", choices=sorted(WORKLOADS))
    ap.add_argument("--n-stars", type=int, default=0, help="override the workload's star count")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--ref-sample", type=int, default=0, help="stars per step of the reference arm (0: derive from the core count and the time budget)")
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="wall-clock budget of the reference arm's steps")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check in front of the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    if not args.n_stars:
        args.n_stars = WORKLOADS[args.workload]["N"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
