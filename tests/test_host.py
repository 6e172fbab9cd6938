"""CPU-side checks: the C-ABI library loads and exports what include/cmcb200.h declares, partition arithmetic, the exact-sum and x87
machinery (host builds of the headers the kernels are made of), and the multi-rank reference harness."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import common as cm
import oraclelib as ol
import refharness as rh
import hostemul as he

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_EMUL = os.path.join(ROOT, "tests", "host_emul")


def test_library_loads_and_exports_every_declared_symbol():
    from cmc_cosmic_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "cmcb200.h")).read()
    declared = sorted(set(re.findall(r"\b(cmcb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = lib.load()
    for sym in declared:
        assert hasattr(L, sym), f"libcmcb200.so does not export {sym}"
    assert set(declared) == set(lib.EXPORTS), "lib.EXPORTS and include/cmcb200.h disagree"


def test_no_cpu_fallback_when_library_is_missing(monkeypatch):
    from cmc_cosmic_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libcmcb200.so")
    with pytest.raises(lib.CmcB200Error):
        lib.load()


def test_product_package_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cmc_cosmic_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"oraclelib|refharness|cmc_oracle|libcmc_ref|hostemul", txt):
                    bad.append(f)
    assert not bad, f"product files reference the oracle: {bad}"


@pytest.mark.parametrize("N,blk,procs", [(10000, 20, 1), (10000, 20, 2), (100000, 40, 8), (99999, 40, 7), (1000003, 40, 16), (39, 40, 4), (0, 40, 3),
                                         (10 ** 7, 40, 8), (2 * 10 ** 6 + 17, 40, 5)])
def test_find_limits_matches_oracle(N, blk, procs):
    from cmc_cosmic_b200.partition import find_limits
    S, E = find_limits(N, blk, procs)
    S0, E0 = ol.find_limits(N, blk, procs)
    assert np.array_equal(S, S0) and np.array_equal(E, E0)
    # the partition tiles 1..N with multiples of blk on every rank but the last
    assert S[0] == 1 and E[-1] == N and np.all(S[1:] == E[:-1] + 1) and np.all((E[:-1] - S[:-1] + 1) % blk == 0)


def _run(exe, *args):
    subprocess.check_call(["make", "-s", "-C", HOST_EMUL])
    p = subprocess.run([os.path.join(HOST_EMUL, "_build", exe), *args], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "PASS" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_exact_sequential_sum_classification_is_bit_exact():
    """exactsum.h: the parallel evaluation of y_i = fl(y_{i-1} +- a_i) (run sums + sequentially resolved events) reproduces the
    sequential roundings on equal-mass / IMF / adversarial chains (tests/host_emul/test_exactsum.cpp)"""
    _run("test_exactsum")


def test_x87_emulation_matches_hardware_long_double():
    """x87emu.h against the host's real 80-bit long double, incl. function_q right at its sign changes (tests/host_emul/test_x87emu.cpp)"""
    _run("test_x87emu", "300000")


@pytest.mark.parametrize("kind,N,masses,deep", [("plummer", 20000, "equal", False), ("king", 20000, "equal", True), ("plummer", 10000, "imf", True),
                                               ("plummer", 200000, "equal", False)])
def test_kernel_core_on_host_matches_oracle(kind, N, masses, deep):
    """orbit_core.h compiled for the host, with the ladders the kernels use (deep = the build with tiny top levels, so these small tables get
    3- and 4-level ladders; N = 200000 gets a 3-level Q ladder with the production constants) and without them"""
    import hostemul
    hostemul.use_deep_ladders(deep)
    try:
        r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 3, masses=masses)
        pot = cm.oracle_potential(r, sm)
        E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
        Rt = 3.5 if masses == "imf" else pot["Rtidal"]
        want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vr, vt, Rt, seed=5)
        for no_ladders in (False, True):
            out = hostemul.get_positions(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0, np.arange(1, N + 1), E, J, vr, vt,
                                         Rt, want["draws"], csr_off=want["draw_off"], no_ladders=no_ladders)
            if no_ladders:
                assert out["q_levels"] == 0 and out["r_levels"] == 0
            else:
                assert out["q_levels"] >= (3 if deep or N >= 200000 else 1) and out["r_levels"] >= (3 if deep else 1), (out["q_levels"], out["r_levels"])
            assert np.array_equal(out["status"], want["status"])
            has = (want["status"] != ol.ST_MASSLESS) & (want["status"] != ol.ST_UNBOUND)
            assert np.array_equal(out["kmin"][has], want["kmin"][has]) and np.array_equal(out["kmax"][has], want["kmax"][has])
            for f in ("rnew", "vrnew", "vtnew"):
                cm.assert_bits_equal(out[f], want[f], f)
            ok = want["status"] == ol.ST_OK
            for f in ("X", "r_peri", "r_apo"):
                cm.assert_bits_equal(out[f][ok], want[f][ok], f)
    finally:
        hostemul.use_deep_ladders(False)


@pytest.mark.skipif(not rh.available(), reason="oracle/_ref/libcmc_ref.so not built")
@pytest.mark.parametrize("procs,N", [(2, 10000), (4, 50000), (3, 10007)])
def test_multirank_reference_equals_single_rank_oracle(procs, N):
    """P forked ranks of the compiled reference (sample sort + load balance + allgather + replicated potential) give the global result
    the oracle computes in one piece -- this is the contract the radius-sharded GPU path has to meet"""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_multirank_check.py"), str(procs), str(N)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MULTIRANK OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.parametrize("deep", [False, True])
@pytest.mark.parametrize("N", [1, 2, 13, 61, 62, 63, 64, 255, 256, 257, 4095, 4096, 4097, 40951, 40959, 40960, 65535, 65536, 70001, 655359, 655361, 1048577])
def test_ladder_find_zero_r_equals_reference_bisection(N, deep):
    """the r ladder used by potential() on the device returns FindZero_r's index for every radius: between, on and beyond the table
    entries, with ties, at every block boundary of every level (deep: the build with a 10-sample top level -> up to 5 levels)"""
    import hostemul
    hostemul.use_deep_ladders(deep)
    try:
        rng = np.random.default_rng(N)
        r = np.sort(rng.uniform(1e-3, 50.0, N))
        if N > 20:
            r[N // 3:N // 3 + 5] = r[N // 3]                      # ties
        sr = np.zeros(N + 2)
        sr[0], sr[1:N + 1], sr[N + 1] = cm.ZERO, r, cm.SF_INFINITY
        probes = np.concatenate([rng.uniform(r[0], r[-1] * 1.01, 4000), r[rng.integers(0, N, min(N, 2000))], r[:40], r[-40:],
                                 np.nextafter(r[:40], 0), np.nextafter(r[-40:], 1e9), [r[-1] * 2, 1e9, r[0]]])
        edges = np.arange(8, N + 2, 8)
        for d in (-1, 0, 1):                                       # radii right at the block boundaries of every level
            idx = np.clip(edges + d, 1, N)
            probes = np.concatenate([probes, sr[idx][:3000], sr[idx][-3000:], np.nextafter(sr[idx][:3000], 1e9)])
        probes = probes[probes >= r[0]]                            # potential() handles r < star_r[1] before the search
        a, b = hostemul.find_zero_r_both(sr, N, probes)
        plan = hostemul.ladder_plan(N + 1)
        if N + 1 < 64:
            assert len(plan["r_shift"]) == 0 and np.all(b == -1)   # small tables keep the plain bisection
            return
        assert plan["r_shift"][-1] == 3 and np.all(np.diff(plan["r_shift"]) == -4) and plan["r_count"][0] <= (10 if deep else 5120)
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, (bad[:5], probes[bad[:5]], a[bad[:5]], b[bad[:5]])
    finally:
        hostemul.use_deep_ladders(False)


@pytest.mark.parametrize("kind,N", [("plummer", 30000), ("king", 30000)])
def test_laddered_find_zero_Q_replays_the_reference_bisection(kind, N):
    """find_zero_Q answered from the Q ladder (block bounds prove the sign of most probes) returns the index of the plain bisection for random
    (E, J) -- bound, unbound, nearly circular, radial -- on both brackets, and with a central mass that makes Phi_0 > Phi_1"""
    import hostemul
    hostemul.use_deep_ladders(True)
    try:
        rng = np.random.default_rng(N)
        r, vr, vt, sm = cm.cluster_state(kind, N, seed=N)
        for cenma in (0.0, 0.05 * N):
            pot = cm.oracle_potential(r, sm, cenma_m=cenma)
            phi = pot["phi"]
            for _ in range(1500):
                g = int(rng.integers(1, N + 1))
                scale = rng.choice([1.0, 0.3, 1e-3, 3.0])
                v2 = scale * rng.uniform(0, 2.0) * abs(phi[g])
                E = phi[g] + 0.5 * v2 * rng.uniform(0, 1)
                J = pot["star_r"][g] * np.sqrt(v2) * rng.uniform(0, 1)
                for (k0, k1) in ((0, g), (g, N + 1)):
                    ka, _ = hostemul.find_zero_Q(pot["star_r"], pot["star_m"], phi, N, N, g, E, J, k0, k1, use_ladder=True)
                    kb, _ = hostemul.find_zero_Q(pot["star_r"], pot["star_m"], phi, N, N, g, E, J, k0, k1, use_ladder=False)
                    assert ka == kb, (g, E, J, k0, k1, ka, kb)
    finally:
        hostemul.use_deep_ladders(False)


def test_kernel_core_on_host_handles_stars_created_this_step():
    """new stars (local index beyond the old shard, global index N_MAX + 2 .., position in star[].rnew): orbit_core.h against the oracle"""
    import hostemul
    N, M = 20000, 150
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=12)
    pot = cm.oracle_potential(r, sm)
    sr, smt, phi, E, J, vra, vta, g_idx, is_new, rnew_in = cm.with_new_stars(r, vr, vt, sm, pot, M, seed=4)
    want = cm.oracle_orbits(sr, smt, phi, N, N, E, J, vra, vta, pot["Rtidal"], seed=8, g_idx=g_idx, is_new=is_new, rnew_in=rnew_in)
    out = hostemul.get_positions(sr, smt, phi, N, N, 0.0, g_idx, E, J, vra, vta, pot["Rtidal"], want["draws"], csr_off=want["draw_off"],
                                 is_new=is_new, rnew_in=rnew_in)
    assert np.array_equal(out["status"], want["status"]) and (want["status"][N:] == ol.ST_OK).sum() > M // 2
    assert np.array_equal(out["kmin"], want["kmin"]) and np.array_equal(out["kmax"], want["kmax"])
    for f in ("rnew", "vrnew", "vtnew", "r_peri", "r_apo"):
        cm.assert_bits_equal(out[f], want[f], f)


@pytest.mark.parametrize("deep", [False, True])
@pytest.mark.parametrize("N,blk,first,nl", [(30000, 1024, 1, None), (30000, 64, 1, None), (5000, 8, 1, None), (20000, 1024, 4001, 9000), (3000, 1, 1, None)])
def test_set_velocities3_block_decomposition_equals_serial_loop(N, blk, first, nl, deep):
    """velocities_core.h (prepare / per-block carries / serial pass over the blocks / apply) against the restatement of the reference's
    serial loop (cmc_utils.c:268-311), bit for bit: chains inside a block, across blocks, a whole shard, a leftover at the end; then one
    pass of the second loop (:323-343) from a received excess"""
    import hostemul
    hostemul.use_deep_ladders(deep)
    try:
        r, vr, vt, sm = cm.cluster_state("plummer", N, seed=N + blk, masses="imf")
        pot = cm.oracle_potential(r, sm, N)
        f = cm.sv3_fields(r, vr, vt, pot["phi"][1:N + 1], seed=blk)
        nl = nl or N
        f = {k: v[first - 1:first - 1 + nl] for k, v in f.items()}
        t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
        g = np.arange(first, first + nl)
        want, flags, left = ol.set_velocities3_first(t, g, f)
        got, kind, left2, cnt = hostemul.set_velocities3(pot["star_r"], pot["star_m"], pot["phi"], N, N, first, f, blk=blk)
        assert cnt["failed"] == 0 and cnt["walked"] > 0 and (cnt["replayed"] > 0 or blk >= nl or first > 1)
        for k in ol.SV3_IO:
            cm.assert_bits_equal(got[k], want[k], k)
        assert left2 == left and np.array_equal((kind & 4) != 0, flags == 1) and np.array_equal((kind & 1) != 0, f["interacted"] == 0)
        for Ein in (0.0, left * 3.0, 1e-9):
            a = {k: v.copy() for k, v in want.items()}
            b = {k: v.copy() for k, v in want.items()}
            ea = ol.set_velocities3_second(t, g, f["interacted"], flags, Ein, a)
            eb = hostemul.set_velocities3_second(pot["star_m"], N, first, kind, Ein, b)
            assert ea == eb
            for k in ("vr", "vt"):
                cm.assert_bits_equal(b[k], a[k], k + " second loop")
    finally:
        hostemul.use_deep_ladders(False)


@pytest.mark.parametrize("N,p,first,nl,masses", [(30000, 20, 1, None, "imf"), (30000, 3, 1, None, "imf"), (20000, 20, 1, None, "equal"),
                                                 (30000, 20, 8001, 9000, "imf"), (30000, 20, 21001, 9000, "imf"), (30000, 20, 1, 9000, "imf")])
def test_calc_sigma_r_as_one_signed_chain(N, p, first, nl, masses):
    """sigma_core.h: calc_sigma_r (reference cmc_dynamics_helper.c:2821-2975) written as one list of signed operations reproduces the
    restatement of the reference's loop bit for bit -- applied in order, and through the exact-scan decomposition with signed terms
    (exactsum.h: seq_classify_signed) the device path will use; whole cluster, first / middle / last shard, both output modes"""
    import hostemul
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=N + p, masses=masses)
    sr, smt = cm.make_tables(r, sm)
    events = 0
    for mave in (False, True):
        a0, s0 = ol.calc_sigma_r(p, sr, smt, N, N, vr, vt, first_global=first, n_local=nl, mave=mave)
        for exact_scan in (False, True):
            a1, s1, ev = hostemul.calc_sigma_r(p, sr, smt, N, N, vr, vt, first_global=first, n_local=nl, mave=mave, exact_scan=exact_scan)
            cm.assert_bits_equal(a1[1:], a0[1:], "r / mean mass")
            cm.assert_bits_equal(s1[1:], s0[1:], "sigma")
            events = max(events, ev)
    # the first window fills from zero (2p events per chain); beyond that every operation that crosses a binade boundary is one -- all of
    # them for the mean-mass chain of N = 20000 equal masses, whose window sum alternates between 39/N and 40/N around 2^-9
    assert 2 * p <= events <= 2 * (2 * p + 2 * ((nl or N) - 1))
    if masses == "equal":
        assert events > 2 * N


def test_find_root_vr_fires_in_the_device_core_and_matches_the_oracle():
    """calc_orbit_new's Brent fallback (reference cmc_orbit.c:459-474, 485-500; cmc_nr.c:500-603) and the helpers it
    reaches (calc_vr_within_interval ...) on a cluster built so that it fires hundreds of times (common.brent_cluster): the per-star
    code the CUDA kernels run (csrc/orbit_core.h, compiled for the host) against the oracle, bit for bit."""
    import hostemul as he
    N = 20000
    r, vr, vt, sm, pot, E, J, planted = cm.brent_cluster(N)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N)
    draws = ol.Taus113(3).draws(200 * N)
    want = ol.get_positions(t, np.arange(1, N + 1), E, J, vr, vt, pot["Rtidal"], draws)
    assert want["n_brent"] > 150 and (want["status"] == ol.ST_NTRY).sum() == 0
    ok = want["status"] == ol.ST_OK
    assert (ok & planted).sum() > 100          # most of the fallbacks end in a regular (non-circular) orbit
    used = draws[:want["draw_off"][-1]]
    for no_ladders in (False, True):
        got = he.get_positions(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0, np.arange(1, N + 1), E, J, vr, vt, pot["Rtidal"], used,
                               csr_off=want["draw_off"], no_ladders=no_ladders)
        assert got["n_brent"] == want["n_brent"]
        for k in ("status", "kmin", "kmax"):
            assert np.array_equal(got[k], want[k]), k
        for k in ("rnew", "vrnew", "vtnew"):
            cm.assert_bits_equal(got[k], want[k], k)
        for k in ("r_peri", "r_apo", "X"):
            cm.assert_bits_equal(got[k][ok], want[k][ok], k)


@pytest.mark.parametrize("kind,N,masses,cenma", [("plummer", 120000, "equal", 0.0), ("king", 40000, "imf", 0.0), ("plummer", 20000, "imf", 50.0), ("plummer", 3000, "equal", 0.0)])
def test_sample_ladder_index_bisections_bit_exact(kind, N, masses, cenma):
    """find_zero_Q through the sample ladder (csrc/orbit_core.h: find_zero_Q_sampled -- the device's default): the signs of all probes far from
    the sign change are PROVEN from a few sampled table entries (concavity of Q over the entries + a rigorous margin), the probes next to it
    are evaluated exactly, the reference's probe sequence is replayed.  k_min / k_max and everything computed from them must equal the
    oracle's bit for bit -- on ordinary clusters, with a central mass, with stars at (and within rounding of) their turning points, where a
    bracket end sits inside the margin and the search must fall back, and with ladders of 2 .. 4 levels (the 'deep' build)."""
    import hostemul as he
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=3, masses=masses)
    pot = cm.oracle_potential(r, sm, cenma_m=cenma)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, cenma)
    for turning in (False, True):
        v = vr.copy()
        g = np.arange(1, N + 1)
        if turning:
            v[::3] = 0.0
            v[1::3] *= 1e-7
            g = np.arange(1, N + 1, max(1, N // 4000))        # their linear scans are O(N) each in the oracle
        E, J = cm.energies(pot["phi"][1:N + 1], r, v, vt)
        d = ol.Taus113(3).draws(200 * len(g))
        want = ol.get_positions(t, g, E[g - 1], J[g - 1], v[g - 1], vt[g - 1], pot["Rtidal"], d)
        used = d[:want["draw_off"][-1]]
        ok = want["status"] == ol.ST_OK
        for deep in (False, True):
            he.use_deep_ladders(deep)
            try:
                got = he.get_positions(pot["star_r"], pot["star_m"], pot["phi"], N, N, cenma, g, E[g - 1], J[g - 1], v[g - 1], vt[g - 1], pot["Rtidal"], used,
                                       csr_off=want["draw_off"], sampled=True)
            finally:
                he.use_deep_ladders(False)
            assert np.array_equal(got["status"], want["status"])
            assert np.array_equal(got["kmin"][ok], want["kmin"][ok]) and np.array_equal(got["kmax"][ok], want["kmax"][ok])
            for k in ("rnew", "vrnew", "vtnew"):
                cm.assert_bits_equal(got[k], want[k], k)


def test_sample_ladder_where_find_root_vr_fires():
    import hostemul as he
    N = 20000
    r, vr, vt, sm, pot, E, J, planted = cm.brent_cluster(N)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N)
    d = ol.Taus113(3).draws(200 * N)
    want = ol.get_positions(t, np.arange(1, N + 1), E, J, vr, vt, pot["Rtidal"], d)
    got = he.get_positions(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0, np.arange(1, N + 1), E, J, vr, vt, pot["Rtidal"], d[:want["draw_off"][-1]],
                           csr_off=want["draw_off"], sampled=True)
    ok = want["status"] == ol.ST_OK
    assert got["n_brent"] == want["n_brent"] > 150 and np.array_equal(got["status"], want["status"])
    assert np.array_equal(got["kmin"][ok], want["kmin"][ok]) and np.array_equal(got["kmax"][ok], want["kmax"][ok])
    for k in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(got[k], want[k], k)


@pytest.mark.parametrize("case,first,nl", [("relax_only", 1, None), ("encounters", 1, None), ("no_relaxation", 1, None), ("encounters", 5001, 7001)])
def test_dynamics_pair_core_equals_oracle(case, first, nl):
    """csrc/dynamics_core.h (the per-pair code of dynamics.cu's kernel) built for the host, where sin / cos / log / pow are glibc's, against
    the oracle's restatement of dynamics_apply's pair loop: bit for bit, incl. a shard that starts in the middle of the cluster"""
    d = cm.dynamics_case(case, first=first, nl=nl)
    N = d["N"]
    sr, sm, phi = d["pot"]["star_r"], d["pot"]["star_m"], d["pot"]["phi"]
    t = ol.tables(sr, sm, phi, N, N)
    dp = ol.dyn_params(**{k: v for k, v in d["cfg"].items() if k in dict(ol.DynParams._fields_)})
    f0 = {k: v.copy() for k, v in d["f"].items()}
    want = ol.dynamics_apply_pairs(t, first, d["dt"], dp, d["draws"], f0, d["bin_a"])
    f1 = {k: v.copy() for k, v in d["f"].items()}
    got = he.dynamics_pairs(sr, sm, phi, N, N, first, d["dt"], d["cfg"], d["draws"], f1, d["bin_a"])
    for k in ("vr", "vt", "E", "J", "Y"):
        cm.assert_bits_equal(f1[k], f0[k], k)
    assert np.array_equal(got["status"], want["status"])
    for k in ("rperi", "W", "rcm"):
        cm.assert_bits_equal(got[k], want[k], k)
    if case == "encounters":
        assert (want["status"] == 1).sum() > 10 and (want["status"] == 2).sum() > 10


SIGMA_SCAN_CASES = [
    # kind, N, masses, p, first, n_local, band_scale
    ("plummer", 30000, "imf", 20, 1, None, 1.0),
    ("king", 20000, "equal", 20, 1, None, 1.0),
    ("plummer", 5000, "imf", 3, 1, None, 1.0),          # BH_AVEKERNEL
    ("plummer", 200, "imf", 20, 1, None, 1.0),          # less than one tile
    ("plummer", 60000, "imf", 20, 20001, 20000, 1.0),   # a middle shard: starts from an empty window p stars before its first star
    ("plummer", 60000, "imf", 20, 40001, 20000, 1.0),   # the last shard: the window stops sliding at the end
    ("plummer", 60000, "imf", 20, 1, 20000, 1.0),       # the first shard
    ("plummer", 300000, "imf", 20, 1, None, 1.0),
    ("plummer", 40000, "imf", 20, 1, None, 3.0e6),      # inflated bands: dozens of hard segments, hard tiles walked segment by segment
    ("king", 40000, "imf", 20, 1, None, 1.0e12),        # bands wider than the values: `broken`, the whole chain replayed by one thread
]


@pytest.mark.parametrize("kind,N,masses,p,first,nl,band_scale", SIGMA_SCAN_CASES)
def test_calc_sigma_r_segment_tables_equal_serial_loop(kind, N, masses, p, first, nl, band_scale):
    """calc_sigma_r as sigma.cu evaluates it (csrc/sigma_scan.h: per-segment tables over the low mantissa bits of the incoming value, tile
    tables, one serial walk over the tiles, replay) on the host against the oracle's serial loop: both outputs, bit for bit"""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + p, masses=masses)
    sr, smt = cm.make_tables(r, sm)
    for mave in (False, True):
        want_a, want_s = ol.calc_sigma_r(p, sr, smt, N, N, vr, vt, first_global=first, n_local=nl, mave=mave)
        a, sg, cnt = he.sigma_scan(p, sr, smt, N, N, vr, vt, first_global=first, n_local=nl, mave=mave, band_scale=band_scale)
        cm.assert_bits_equal(a[1:], want_a[1:], "r / mean mass")
        cm.assert_bits_equal(sg[1:], want_s[1:], "sigma")
        assert cnt["inconsistent"] == 0
    if band_scale == 1.0:
        assert cnt["broken"] == 0 and cnt["hard_segments"] <= 2, cnt
    elif band_scale < 1e10:
        assert cnt["broken"] == 0 and cnt["hard_segments"] > 50 and cnt["hard_tiles"] > 10, cnt
    else:
        assert cnt["broken"] == 1


CI_INPUT = "/root/reference/ci/input.hdf5"


@pytest.mark.skipif(not os.path.exists(CI_INPUT), reason="the reference tree (ci/input.hdf5) is not here")
def test_input_hdf5_reader_reads_the_reference_ci_input():
    """cmc_cosmic_b200/hdf5_input.py (the file cmc_read_hdf5_file reads, reference src/libs/fitslib.c:209-448) on the reference's
    own ci/input.hdf5: the attributes, both tables, and the columns equal to what the committed fixture ci_input.npz was generated from (a raw-
    byte scan of the same file by tests/golden/make_golden.py -- two independent readers)."""
    from cmc_cosmic_b200 import hdf5_input
    d = hdf5_input.read_input(CI_INPUT)
    assert d["NOBJ"] == 10000 and d["NBINARY"] == 1967 and d["RTID"] == 1.0e6 and d["RVIR"] == 1.0 and abs(d["Z"] - 0.014) < 1e-15
    o, b = d["obj"], d["bin"]
    n, nb = d["NOBJ"], d["NBINARY"]
    g = np.load(os.path.join(cm.GOLDEN, "ci_input.npz"))
    cm.assert_bits_equal(o["r"][1:n + 1], g["r"], "r")
    cm.assert_bits_equal(o["vr"][1:n + 1], g["vr"], "vr")
    cm.assert_bits_equal(o["vt"][1:n + 1], g["vt"], "vt")
    cm.assert_bits_equal(o["m"][1:n + 1] * n, g["star_m"], "star_m")
    assert np.array_equal(o["binind"][1:n + 1].astype(np.int64), g["binind"])
    assert np.array_equal(o["id"][1:n + 1], np.arange(1, n + 1)) and o["r"][0] == np.finfo(float).tiny and o["r"][n + 1] == 1e40
    assert abs(o["m"].sum() - 1.0) < 1e-12                      # total mass in units of the cluster mass (fitslib convention)
    assert np.array_equal(np.sort(o["binind"][o["binind"] > 0]), np.arange(1, nb + 1)) and np.array_equal(b["index"][1:nb + 1], np.arange(1, nb + 1))
    assert np.all(b["a"][1:nb + 1] > 0) and np.all((b["e"][1:nb + 1] >= 0) & (b["e"][1:nb + 1] < 1))
    k = np.nonzero(o["binind"] > 0)[0]                          # a binary's object carries the sum of the component masses
    np.testing.assert_allclose(o["m"][k], (b["m1"] + b["m2"])[o["binind"][k].astype(int)], rtol=1e-12)


def test_input_hdf5_reader_refuses_what_it_does_not_parse(tmp_path):
    from cmc_cosmic_b200 import hdf5_input
    p = tmp_path / "x.hdf5"
    p.write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(hdf5_input.Hdf5FormatError):
        hdf5_input.read_input(str(p))
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + bytes(200))    # superblock version 2
    with pytest.raises(hdf5_input.Hdf5FormatError):
        hdf5_input.read_input(str(p))


@pytest.mark.parametrize("kind", ["hover_at_power_of_two", "zeros", "wide_range", "steep_trend", "rising_trend", "spikes"])
@pytest.mark.parametrize("n,first", [(30011, 1), (9000, 4001)])
def test_calc_sigma_r_segment_tables_on_adversarial_inputs(kind, n, first):
    """The replay through segment / group / tile tables must give the serial loop's bits whatever the values do: sums hovering on a power of
    two, chains that reach exactly zero, terms over ten decades, monotone trends over forty binades, isolated spikes.  What the classification
    cannot prove is replayed operation by operation -- the counters show that these inputs do reach those paths."""
    rng = np.random.default_rng(len(kind) * 1000 + n)
    N = 60000
    m, vr, vt = cm.adversarial_sigma_inputs(kind, N, rng)
    r = np.sort(rng.random(N)) + 0.01
    sr, smt = cm.make_tables(r, m)
    nl = n
    want_a, want_s = ol.calc_sigma_r(20, sr, smt, N, N, vr, vt, first_global=first, n_local=nl, mave=True)
    a, sg, cnt = he.sigma_scan(20, sr, smt, N, N, vr, vt, first_global=first, n_local=nl, mave=True)
    cm.assert_bits_equal(a[1:], want_a[1:], f"{kind}: mean mass")
    both_nan = np.isnan(sg[1:]) & np.isnan(want_s[1:])           # 0 / 0 where a whole window is at rest: NaN on both sides
    cm.assert_bits_equal(np.where(both_nan, 0.0, sg[1:]), np.where(both_nan, 0.0, want_s[1:]), f"{kind}: sigma")
    assert cnt["inconsistent"] == 0, cnt
    if kind in ("hover_at_power_of_two",):
        assert cnt["hard_segments"] > 50, cnt
    if kind in ("wide_range", "rising_trend", "spikes"):
        assert cnt["wide_segments"] + cnt["hard_tiles"] + cnt["broken"] > 0, cnt
