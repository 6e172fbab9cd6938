"""Run by tests/test_host.py in a subprocess: P forked ranks of the compiled reference (mini-MPI shim) do one sample sort +
post_sort_comm + calc_potential_new; every rank checks its shard and the replicated tables against the oracle.  TEST-ONLY."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np

import common as cm
import oraclelib as ol
import refharness as rh


def main(procs, N):
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=21, masses="imf")
    rng = np.random.default_rng(4)
    rnew = r * np.exp(rng.normal(0.0, 0.5, N))            # shuffled radii: the sort has work to do
    R = rh.Ref()
    assert R.L.ref_spawn(procs, int(3.0 * (N / procs + 4096) * 424) + (1 << 22)) == 0
    rank = R.L.ref_rank()
    R.init(N, n_star_dim=int(2.5 * (N // procs + 4096)), n_table_dim=N + 4096, min_chunk=40)
    S = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start())).copy()
    E = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End())).copy()
    S0, E0 = ol.find_limits(N, 40, procs)
    ok = np.array_equal(S, S0) and np.array_equal(E, E0)
    lo, hi = int(S[rank]), int(E[rank])
    nl = hi - lo + 1
    R.star_field("r")[1:nl + 1] = rnew[lo - 1:hi]
    R.star_field("m")[1:nl + 1] = sm[lo - 1:hi]
    R.star_field("id", np.int64)[1:nl + 1] = np.arange(lo, hi + 1)
    R.set_scalars(N, nl)
    R.L.ref_qsorts_new()
    R.L.ref_post_sort_comm()
    R.L.ref_calc_potential_new()
    order = np.argsort(rnew, kind="stable")
    S = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start())).copy()
    E = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End())).copy()
    lo, hi = int(S[rank]), int(E[rank])
    ok = ok and int(R.scalars()["N_MAX_NEW"]) == hi - lo + 1
    ok = ok and np.array_equal(R.star_field("id", np.int64)[1:hi - lo + 2], order[lo - 1:hi] + 1)
    want = cm.oracle_potential(rnew[order], sm[order], N)
    ok = ok and np.array_equal(R.star_r()[1:N + 1], rnew[order]) and np.array_equal(R.star_phi()[:N + 2].view(np.uint64), want["phi"][:N + 2].view(np.uint64))
    # comp_mass_percent (reference cmc_utils.c:872-985): rank-local cumulative sums + Exscan + the owner's values sent to rank 0
    vrs, vts = vr[order], vt[order]
    R.star_field("vr")[1:hi - lo + 2] = vrs[lo - 1:hi]
    R.star_field("vt")[1:hi - lo + 2] = vts[lo - 1:hi]
    pcs = np.array([1e-4, 0.001, 0.01, 0.1, 0.3, 0.5, 0.5000001, 0.9, 0.99, 0.999999, 1.5])
    Mtot = R.scalars()["Mtotal"]
    got = R.comp_mass_percent(pcs, initial_total_mass=0.97)
    if rank == 0:
        sr, smt = R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy()
        parts, prefix = [], np.zeros(4)
        for q in range(procs):   # the oracle as every rank in turn; MPI_Exscan = running sum of the lower ranks' totals in rank order
            o = ol.comp_mass_percent(sr, smt, N, N, vrs[S[q] - 1:E[q]], vts[S[q] - 1:E[q]], pcs, Mtot, first_global=int(S[q]), prefix=prefix.copy(),
                                     is_rank0=(q == 0), initial_total_mass=0.97)
            parts.append(o)
            prefix = o["totals"].copy() if q == 0 else prefix + o["totals"]
        for f in ol.MASS_PC_FIELDS[:4]:
            ok = ok and np.array_equal(got[f].view(np.uint64), parts[0][f].view(np.uint64))
        for f in ol.MASS_PC_FIELDS[4:]:
            merged = np.full(len(pcs), -1.0)
            for o in parts:
                merged[o["owner"] == 1] = o[f][o["owner"] == 1]
            ok = ok and np.array_equal(got[f].view(np.uint64), merged.view(np.uint64))
        ok = ok and parts[0]["filled"] == len(pcs) - 1 and got["mass_r"][-1] == -1.0     # 150 % of the mass is never reached
        ok = ok and len({int(np.argmax([o["owner"][i] for o in parts])) for i in range(len(pcs) - 1)}) >= min(procs, 2)  # several owners
    # calc_sigma_r (reference cmc_dynamics_helper.c:2821-2975): every rank starts its running window sums from the ghost stars its
    # neighbours sent (ring of MPI_Send / MPI_Recv)
    for mave in (False, True):
        a, sg = R.calc_sigma_r(20, hi - lo + 1, mave)
        a0, sg0 = ol.calc_sigma_r(20, R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), N, N, vrs, vts, first_global=lo, n_local=hi - lo + 1, mave=mave)
        ok = ok and np.array_equal(a[1:].view(np.uint64), a0[1:].view(np.uint64)) and np.array_equal(sg[1:].view(np.uint64), sg0[1:].view(np.uint64))
    # write_snapshot (reference cmc_io.c:2960-3345) with the HDF5 calls recorded: every rank's rows (rank 0 through H5TBmake_table, the others
    # through H5TBappend_records) against the host build of csrc/snapshot_core.h for its shard -- get_global_idx with Start[myid] > 1, the
    # calc_average_mass_sqr windows of binaries at both ends of a shard
    import hostemul as he
    from cmc_cosmic_b200 import layout
    nl = hi - lo + 1
    rng3 = np.random.default_rng(1000 + rank)
    kb = 12
    spots = np.unique(np.concatenate(([1, 2, 19, 20, 21, nl - 20, nl - 19, nl - 1, nl], rng3.integers(1, nl + 1, 3))))[:kb]
    bi = np.zeros(nl, np.int64)
    bi[spots - 1] = np.arange(1, len(spots) + 1)
    R.star_field("binind", np.int64)[1:nl + 1] = bi
    bb = R.binary_bytes()
    bb[1:len(spots) + 1] = rng3.uniform(0.1, 5.0, (len(spots), R.sizeof_binary // 8)).view(np.uint8).reshape(len(spots), R.sizeof_binary)
    sb = R.star_bytes()
    layout.star_field(sb, "se_k")[1:nl + 1] = rng3.integers(0, 16, nl)
    for k in ("E", "J", "rad"):
        R.star_field(k)[1:nl + 1] = rng3.uniform(0.01, 1.0, nl)
    sgm = np.concatenate(([0.0], rng3.uniform(0.2, 0.9, nl)))
    for bh_only in (False, True):
        rows, _ = R.write_snapshot(3.0e38, 5.0e18, sigma=sgm, avekernel=20, bh_only=bh_only)
        nanb = 0xFFF8000000000000
        # the WHOLE tables: for the last p stars of the last rank calc_average_mass_sqr's window is not clamped (its N_LIMIT is the global
        # clus.N_MAX, its indices are local) and get_global_idx maps the overshoot to the table entries behind N_MAX + 1
        mine = he.snapshot_rows(R.star_bytes().copy(), R.binary_bytes().copy(), R.star_r().copy(), R.star_m().copy(), R.star_phi().copy(),
                                sgm, nl, nl, lo, N, N, 3.0e38, 5.0e18, 20, nanb, bh_only=bh_only)
        ok = ok and len(rows) == len(mine) and (len(rows) == nl if not bh_only else 0 < len(rows) < nl)
        for name in layout.SNAPSHOT_FIELDS:
            a, b = rows[name], mine[name]
            same = np.array_equal(a.view(np.uint64), b.view(np.uint64)) if a.dtype == np.float64 else np.array_equal(a, b)
            ok = ok and same
    R.star_field("binind", np.int64)[1:nl + 1] = 0
    # set_velocities3 (reference cmc_utils.c:254-352): the excess-energy carry runs through every rank's shard and on to the next rank
    # (MPI_Send / MPI_Recv, repeated while the all-reduced leftover is > 0)
    phi = R.star_phi()[:N + 2].copy()
    f = cm.sv3_fields(rnew[order], vrs, vts, phi[1:N + 1], seed=N + procs)
    nl = hi - lo + 1
    for k, v in f.items():
        R.star_field(k, np.int64 if k == "interacted" else np.float64)[1:nl + 1] = v[lo - 1:hi]
    R.L.ref_set_Eoops(0.0)
    R.L.ref_energy_conservation3()
    t = ol.tables(R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), phi, N, N, 0.0)
    shards = [(int(S[q]), int(E[q] - S[q] + 1)) for q in range(procs)]
    io, eoops, passes = ol.set_velocities3(t, shards, [{k: v[S[q] - 1:E[q]] for k, v in f.items()} for q in range(procs)])
    for k in ol.SV3_IO:
        ok = ok and np.array_equal(io[rank][k].view(np.uint64), R.star_field(k)[1:nl + 1].view(np.uint64))
    ok = ok and R.L.ref_get_Eoops() == eoops[rank] and passes >= (2 if procs > 2 else 1)
    # tidally_strip_stars (reference cmc_evolution_thr.c:443-645): per-rank passes, the stripped mass all-reduced per pass, the escaper sums
    # all-reduced at the end
    rng2 = np.random.default_rng(77)
    Eg = -rng2.uniform(0.05, 1.0, N)
    Eg[rng2.choice(N, N // 100, replace=False)] = -rng2.uniform(0.0, 2.8e-6, N // 100)      # around alpha * Phi(Rtidal) ~ -1.4e-6
    fg = dict(E=Eg, J=rng2.uniform(0.0, 1.0, N), Eint=rng2.uniform(0.0, 1e-7, N), r_apo=rnew[order] * 2.0, rnew=rnew[order].copy(),
              vrnew=vrs.copy(), vtnew=vts.copy())
    for k, v in fg.items():
        R.star_field(k)[1:nl + 1] = v[lo - 1:hi]
    R.star_field("binind", np.int64)[1:nl + 1] = 0
    smt = R.star_m()[:N + 2].copy()            # before the call: destroy_obj sets star_m[g] = DBL_MIN for the stars THIS rank strips
    Rt, TML, DT = R.tidally_strip_stars(1)
    sc = R.scalars()
    parts = [{k: v[S[q] - 1:E[q]].copy() for k, v in fg.items()} for q in range(procs)]
    states = [dict(Mtotal=sc["Mtotal"], TidalMassLoss=0.0, OldTidalMassLoss=0.0, Eescaped=0.0, Jescaped=0.0, Eintescaped=0.0, Ebescaped=0.0) for _ in range(procs)]
    ol.tidally_strip_loop_ranks([(lambda thr, q=q: ol.tidally_strip_select(1, thr, parts[q])) for q in range(procs)],
                                lambda x: ol.lib().orc_potential(C.byref(t), x),
                                [(lambda i, q=q: smt[S[q] + i - 1]) for q in range(procs)],
                                [(lambda i, q=q: dict(E=parts[q]["E"][i - 1], J=parts[q]["J"][i - 1], Eint=parts[q]["Eint"][i - 1], binind=0)) for q in range(procs)],
                                [None] * procs, states, N, 1.0e6, 1, 0.11)
    me = states[rank]
    ok = ok and (Rt, TML) == (me["Rtidal"], me["TidalMassLoss"]) and me["passes"] >= 2 and len(me["stripped"]) > 3
    for k in ("Eescaped", "Jescaped", "Eintescaped", "Etidal"):
        ok = ok and sc[k] == me[k]
    for k in ("rnew", "vrnew", "vtnew"):
        ok = ok and np.array_equal(R.star_field(k)[1:nl + 1].view(np.uint64), parts[rank][k].view(np.uint64))
    if not ok:
        print(f"rank {rank}: MISMATCH", flush=True)
        sys.exit(3)
    rc = R.L.ref_join()
    if rank == 0:
        print("MULTIRANK OK" if rc == 0 else "MULTIRANK FAIL", flush=True)
        sys.exit(0 if rc == 0 else 4)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
