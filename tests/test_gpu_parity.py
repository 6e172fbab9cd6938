"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on identical inputs.

Bars (BASELINE.json north_star): sort permutation / integer indices bit-exact; Phi rel. err <= 1e-12; r_min, r_max and the
new r rel. err <= 1e-10 given the same draws.  The implementation aims higher -- every FP64 output is compared BIT FOR BIT.
"""
import os

import numpy as np
import pytest

import common as cm
import oraclelib as ol
from cmc_cosmic_b200.cluster import Cluster
from cmc_cosmic_b200.layout import star_field, binary_field, STAR_DTYPE_SIZE, BINARY_DTYPE_SIZE

pytestmark = pytest.mark.gpu


def _potential_case(r, star_m, N_STAR=None, cenma_m=0.0, cenma_m_new=0.0, N_MAX_in=None):
    n = len(r)
    want = cm.oracle_potential(r, star_m, N_STAR, cenma_m, cenma_m_new, N_MAX_in=N_MAX_in)
    c = Cluster(N_STAR or n, n_star_dim=n + 64)
    c.star_r[:n + 2], c.star_m[:n + 2] = cm.make_tables(r, star_m)
    c.clus.N_MAX = n if N_MAX_in is None else N_MAX_in
    c.cenma.m, c.cenma.m_new = cenma_m, cenma_m_new
    c.calc_potential_new()
    nm = want["N_MAX"]
    assert c.clus.N_MAX == nm
    assert c.Mtotal == want["Mtotal"] and c.Rtidal == want["Rtidal"] and c.cenma.m == want["cenma_m"]
    cm.assert_bits_equal(c.star_phi[:nm + 2], want["phi"][:nm + 2], "star_phi")
    assert c.star_r[nm + 1] == cm.SF_INFINITY
    ev, fb = c.potential_stats()
    c.close()
    return ev, fb


@pytest.mark.parametrize("kind,N,masses", [("plummer", 1000, "equal"), ("plummer", 100000, "equal"), ("king", 300000, "equal"),
                                           ("plummer", 100000, "imf"), ("plummer", 1 << 20, "equal"), ("plummer", 2049, "imf")])
def test_potential_bit_exact(kind, N, masses):
    r, vr, vt, sm = cm.cluster_state(kind, N, masses=masses)
    ev, fb = _potential_case(r, sm)
    assert fb == 0 and 0 < ev < 2000


def test_potential_edge_cases():
    rng = np.random.default_rng(5)
    # central mass, pending central mass, N_STAR != N, stripped stars at the end (N_MAX shrinks)
    r = np.sort(rng.uniform(0.01, 50.0, 5000)); sm = rng.uniform(0.2, 3.0, 5000)
    _potential_case(r, sm, N_STAR=6000, cenma_m=12.5, cenma_m_new=0.75)
    r2 = r.copy(); r2[-37:] = cm.SF_INFINITY
    _potential_case(r2, sm, N_STAR=6000)
    # tiny systems
    _potential_case(np.array([0.5]), np.array([1.0]))
    _potential_case(np.array([0.5, 0.7, 3.0]), np.array([1.0, 1.0, 1.0]))
    # N_MAX_in smaller than the table
    _potential_case(r, sm, N_MAX_in=4000)
    # adversarial for the exact-sum machinery: many exact ties -> many events, possibly the sequential fallback
    sm3 = np.where(rng.uniform(size=5000) < 0.5, 2.0 ** -20, 1.0 + 2.0 ** -30)
    _potential_case(r, sm3)


def _sort_case(n, n_bin, n_stripped, seed, ties=False):
    rng = np.random.default_rng(seed)
    nbd = n_bin + 8
    c = Cluster(n, n_star_dim=n + 64, n_bin_dim=nbd)
    c.star[:] = rng.integers(0, 256, c.star.shape, dtype=np.uint8)        # every byte of the record must travel
    c.binary[:] = rng.integers(0, 256, c.binary.shape, dtype=np.uint8)
    r = rng.uniform(1e-3, 1e2, n) if not ties else rng.integers(1, 50, n).astype(np.float64)
    strip = rng.choice(n, n_stripped, replace=False)
    r[strip] = cm.SF_INFINITY
    c.field("r")[1:n + 1] = r
    c.field("m")[1:n + 1] = rng.uniform(0.1, 2.0, n)
    binind = np.zeros(n, np.int64)
    live = np.setdiff1d(np.arange(n), strip)
    who = rng.choice(live, min(n_bin, live.size), replace=False)
    binind[who] = rng.permutation(np.arange(1, who.size + 1))
    c.field("binind")[1:n + 1] = binind
    c.clus.N_MAX_NEW = n
    ref_star = c.star[1:n + 1].copy(); ref_bin = c.binary[1:].copy()
    n_new, nb, perm = ol.sort_stars(ref_star, n, STAR_DTYPE_SIZE, 0, 152, ref_bin, nbd, BINARY_DTYPE_SIZE)
    c.qsorts_new()
    assert c.clus.N_MAX == n_new and c.clus.N_MAX_NEW == n_new and c.N_b_local == nb
    assert np.array_equal(c.star[1:n_new + 1], ref_star[:n_new]), "sorted star records differ"
    assert np.array_equal(c.binary[1:], ref_bin), "binary array differs"
    assert np.array_equal(c.star_r[1:n_new + 1], star_field(ref_star, "r")[:n_new])
    assert np.array_equal(c.star_m[1:n_new + 1], star_field(ref_star, "m")[:n_new])
    c.close()


@pytest.mark.parametrize("n,n_bin,n_stripped", [(1, 0, 0), (2, 1, 1), (4095, 100, 3), (4097, 0, 0), (100000, 9000, 41), (1000003, 50000, 1000)])
def test_sort_bit_exact(n, n_bin, n_stripped):
    _sort_case(n, n_bin, n_stripped, seed=n)


def test_sort_ties_and_all_stripped():
    _sort_case(50000, 100, 10, seed=3, ties=True)     # massive ties: stable order == the oracle's tie rule
    _sort_case(100, 0, 100, seed=4)                   # everything stripped


def _orbit_case(kind, N, masses="equal", seed=1234, tidal=None, cenma_m=0.0):
    r, vr, vt, sm = cm.cluster_state(kind, N, masses=masses, seed=seed)
    pot = cm.oracle_potential(r, sm, cenma_m=cenma_m)
    phi, sr, smt = pot["phi"], pot["star_r"], pot["star_m"]
    E, J = cm.energies(phi[1:N + 1], r, vr, vt)
    Rtidal = tidal if tidal is not None else pot["Rtidal"]
    want = cm.oracle_orbits(sr, smt, phi, N, N, E, J, vr, vt, Rtidal, seed=seed, cenma_m=cenma_m)
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.star_phi[:N + 2] = phi[:N + 2]
    c.field("E")[1:N + 1], c.field("J")[1:N + 1] = E, J
    c.Rtidal = Rtidal
    c.cenma.m = cenma_m
    status = c.new_orbits_calculate(draws=want["draws"], draw_off=want["draw_off"])
    assert np.array_equal(status, want["status"])
    ok = want["status"] == ol.ST_OK
    has_orbit = (want["status"] != ol.ST_MASSLESS) & (want["status"] != ol.ST_UNBOUND)
    assert np.array_equal(c.last_kmin[has_orbit], want["kmin"][has_orbit]) and np.array_equal(c.last_kmax[has_orbit], want["kmax"][has_orbit])
    for f in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(c.field(f)[1:N + 1], want[f], f)
    for f in ("X", "r_peri", "r_apo"):
        cm.assert_bits_equal(c.field(f)[1:N + 1][ok], want[f][ok], f)
    tid = want["status"] == ol.ST_TIDAL
    cm.assert_bits_equal(c.field("r_apo")[1:N + 1][tid], want["r_apo"][tid], "r_apo(tidal)")
    assert c.max_r == want["max_rad"]
    assert c.last_orbit.n_draws == len(want["draws"])
    res = c.last_orbit
    c.close()
    return want, res


@pytest.mark.parametrize("kind,N,masses", [("plummer", 2000, "equal"), ("plummer", 100000, "equal"), ("king", 200000, "equal"), ("plummer", 50000, "imf")])
def test_orbit_step_bit_exact(kind, N, masses):
    _orbit_case(kind, N, masses)


def test_orbit_step_deep_ladders_bit_exact(monkeypatch):
    """N = 700000: three levels in both ladders (Q: shifts 10/7/4, r: 11/7/3), every probe of the index bisections answered from the
    coarsest level that proves its sign -- kmin / kmax / r_peri / r_apo / rnew still the reference's bit for bit"""
    monkeypatch.setenv("CMCB200_ORBIT_LADDERS", "3")
    _orbit_case("plummer", 700000, seed=99)


def test_orbit_step_with_central_mass(monkeypatch):
    """cenma.m > 0 makes star_phi[0] > star_phi[1] (cmc_utils.c:727): the Q-ladder blocks carry the true min / max of Phi, not its ends"""
    _orbit_case("king", 120000, seed=17, cenma_m=0.02 * 120000)
    monkeypatch.setenv("CMCB200_ORBIT_LADDERS", "3")
    _orbit_case("king", 120000, seed=17, cenma_m=0.02 * 120000)


@pytest.mark.parametrize("knobs", [{"CMCB200_ORBIT_LADDERS": "0"}, {"CMCB200_ORBIT_LADDERS": "3"}, {"CMCB200_ORBIT_LADDERS": "2", "CMCB200_ORBIT_CTAS": "3"},
                                   {"CMCB200_ORBIT_LADDERS": "3", "CMCB200_QLAD_MAX_SHIFT": "7", "CMCB200_QLAD_MAX_WIDTH_LOG2": "11"}])
def test_orbit_step_tuning_knobs_do_not_change_results(knobs, monkeypatch):
    """every combination of the search structures (no ladders / r ladder [default] / Q ladder / both, partial Q ladders read in place) and
    the register variants of the orbit kernel is bit-exact"""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    _orbit_case("plummer", 300000, seed=5)


def test_host_buffer_orbit_step_is_chunk_pipelined_and_touches_only_its_fields():
    """cmcb200_new_orbits_calculate moves the first 128 bytes of every record in and bytes 64..127 out as strided DMA copies, pipelined over
    chunks of the star range: against the resident form (whole records uploaded, one range) every byte of star[] must agree -- including the
    fields the step must not touch (star[] starts as random bytes) and every status-dependent store (tidal strippings, planted unbound and
    massless stars); N is large enough for three chunks, pinned and pageable host arrays"""
    N = 1200000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=61)
    rng = np.random.default_rng(3)
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    E[rng.choice(N, 50, replace=False)] = 0.5             # unbound
    outs = []
    for mode in ("pinned", "pageable", "resident"):
        c = Cluster(N, n_star_dim=N + 64, pinned=(mode == "pinned"))
        c.star[:] = np.random.default_rng(1).integers(0, 256, c.star.shape, dtype=np.uint8)
        c.field("binind")[:] = 0
        c.load(r, vr, vt, sm)
        c.star_m[1 + 777] = 0.0                             # massless
        c.star_phi[:N + 2] = pot["phi"][:N + 2]
        c.field("E")[1:N + 1], c.field("J")[1:N + 1] = E, J
        c.Rtidal = 4.0
        if mode == "resident":
            c.upload()
            res = c.dev_new_orbits(seed=11, want_status=True)
            status = c.dev_read_status()[0]
            c.download()
            h2d = d2h = 0
        else:
            status = c.new_orbits_calculate(seed=11)
            res = c.last_orbit
            h2d, d2h = c.transfer_bytes()
        outs.append((status.copy(), c.star[:N + 1].copy(), res.max_r, res.n_draws, h2d, d2h))
        c.close()
    a, b, ref = outs
    assert set(np.unique(ref[0])) >= {ol.ST_OK, ol.ST_MASSLESS, ol.ST_UNBOUND, ol.ST_TIDAL}
    for x, name in ((a, "pinned"), (b, "pageable")):
        assert np.array_equal(x[0], ref[0]) and x[2] == ref[2] and x[3] == ref[3], name
        assert np.array_equal(x[1], ref[1]), f"star[] differs between the {name} host-buffer form and the resident form"
        assert x[4] == N * 128 + 3 * 8 * (N + 2) and x[5] == N * 64, (name, x[4], x[5])


def test_orbit_step_branches():
    # a small tidal radius exercises the TIDAL branch; unbound / massless / circular stars are planted by hand
    want, res = _orbit_case("plummer", 20000, tidal=3.0)
    assert (want["status"] == ol.ST_TIDAL).sum() > 10


def test_resident_pipeline_matches_host_buffer_form():
    """orbit -> transfer -> sort -> potential on resident arrays == the same through host buffers and the oracle."""
    N = 60000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=77)
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vr, vt, pot["Rtidal"], seed=9)
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.calc_potential_new()
    c.compute_energy()
    c.upload()
    c.dev_new_orbits(draws=want["draws"], draw_off=want["draw_off"])
    c.dev_transfer_new_positions()
    c.dev_qsorts_new()
    c.dev_calc_potential_new()
    c.dev_compute_energy()
    c.download()
    # oracle: same chain on the host
    rnew, vrn, vtn = want["rnew"], want["vrnew"], want["vtnew"]
    order = np.argsort(rnew, kind="stable")
    n_new = int((rnew < cm.SF_INFINITY).sum())
    assert c.clus.N_MAX == n_new
    cm.assert_bits_equal(c.star_r[1:n_new + 1], rnew[order][:n_new], "sorted r")
    pot2 = cm.oracle_potential(rnew[order][:n_new], sm[order][:n_new], N_STAR=N)
    cm.assert_bits_equal(c.star_phi[:n_new + 2], pot2["phi"][:n_new + 2], "phi after the step")
    E2, J2 = cm.energies(pot2["phi"][1:n_new + 1], rnew[order][:n_new], vrn[order][:n_new], vtn[order][:n_new])
    cm.assert_bits_equal(c.field("E")[1:n_new + 1], E2, "E")
    cm.assert_bits_equal(c.field("J")[1:n_new + 1], J2, "J")
    cm.assert_bits_equal(c.field("rOld")[1:n_new + 1], r[order][:n_new], "rOld")
    c.close()


def test_orbit_step_linear_scan_is_deferred_and_bit_exact():
    """Stars with v_r == 0 sit exactly on a turning point: Q at their own index is 0 up to rounding, and where it is <= 0 the reference
    runs get_positive_Q_index's linear scan from k = 0 (cmc_orbit.c:46-57).  The CUDA path queues those stars and runs the scans as a
    grid-wide find-first; the result must be the reference's, bit for bit."""
    N = 60000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=31)
    rng = np.random.default_rng(2)
    vr[rng.choice(N, 400, replace=False)] = 0.0
    pot = cm.oracle_potential(r, sm)
    phi, sr, smt = pot["phi"], pot["star_r"], pot["star_m"]
    E, J = cm.energies(phi[1:N + 1], r, vr, vt)
    want = cm.oracle_orbits(sr, smt, phi, N, N, E, J, vr, vt, pot["Rtidal"], seed=5)
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.star_phi[:N + 2] = phi[:N + 2]
    c.field("E")[1:N + 1], c.field("J")[1:N + 1] = E, J
    c.Rtidal = pot["Rtidal"]
    status = c.new_orbits_calculate(draws=want["draws"], draw_off=want["draw_off"])
    assert c.last_orbit.n_deferred > 20, "the test no longer reaches the deferred path"
    assert np.array_equal(status, want["status"])
    assert np.array_equal(c.last_kmin, want["kmin"]) and np.array_equal(c.last_kmax, want["kmax"])
    for f in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(c.field(f)[1:N + 1], want[f], f)
    assert c.max_r == want["max_rad"] and c.last_orbit.n_draws == len(want["draws"])
    c.close()


def test_orbit_step_philox_mode():
    """Production mode (no recorded draws): a counter-based Philox stream per star keyed by (seed, tcount, global index).  Deterministic,
    independent of launch geometry, distinct per seed/step, and statistically the same orbit sampling as the reference's taus113 run:
    every new radius inside its [r_peri, r_apo], mean number of rejection tries ~ 1.9, same radial distribution (KS distance)."""
    N = 200000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=8)
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vr, vt, pot["Rtidal"], seed=5)
    outs = []
    for seed, tcount in ((7, 1), (7, 1), (7, 2), (8, 1)):
        c = Cluster(N, n_star_dim=N + 64)
        c.load(r, vr, vt, sm)
        c.star_phi[:N + 2] = pot["phi"][:N + 2]
        c.field("E")[1:N + 1], c.field("J")[1:N + 1] = E, J
        c.Rtidal = pot["Rtidal"]
        c.tcount = tcount
        status = c.new_orbits_calculate(seed=seed)
        outs.append((c.field("rnew")[1:N + 1].copy(), c.field("r_peri")[1:N + 1].copy(), c.field("r_apo")[1:N + 1].copy(), status.copy(), c.last_orbit.n_draws))
        c.close()
    a, b, c2, d = outs
    assert np.array_equal(a[0], b[0]), "same (seed, tcount) must reproduce"
    assert not np.array_equal(a[0], c2[0]) and not np.array_equal(a[0], d[0])
    rn, rp, ra, st, nd = a
    assert np.array_equal(st, want["status"])                       # the branch taken does not depend on the draws here (no tidal cut)
    ok = st == ol.ST_OK
    cm.assert_bits_equal(rp[ok], want["r_peri"][ok], "r_peri")     # the orbit itself is deterministic
    cm.assert_bits_equal(ra[ok], want["r_apo"][ok], "r_apo")
    assert np.all((rn[ok] >= rp[ok] * (1 - 1e-12)) & (rn[ok] <= ra[ok] * (1 + 1e-12)))
    tries = (nd / ok.sum() - 1) / 2
    assert abs(tries - (len(want["draws"]) / ok.sum() - 1) / 2) < 0.02, tries
    xs, ys = np.sort(rn[ok]), np.sort(want["rnew"][ok])
    ks = np.max(np.abs(np.searchsorted(xs, ys, side="right") / xs.size - np.arange(1, ys.size + 1) / ys.size))
    assert ks < 1.36 * np.sqrt(2.0 / ok.sum()) * 1.5, ks          # two-sample KS at ~5 % with slack


def test_fused_step_equals_unfused_step():
    """dev_step(fused=True) -- E, J computed inside the orbit step, r <- rnew applied while the sort's gather moves the records -- against
    the same step made of separate passes: every byte of star[], binary[] and the tables after three steps (Philox draws, binaries)."""
    N, nb = 150000, 9000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=41, masses="imf")
    rng = np.random.default_rng(6)
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    outs = []
    for fused in (True, False):
        c = Cluster(N, n_star_dim=N + 64, n_bin_dim=nb + 8)
        c.star[:] = np.random.default_rng(1).integers(0, 256, c.star.shape, dtype=np.uint8)
        c.binary[:] = np.random.default_rng(2).integers(0, 256, c.binary.shape, dtype=np.uint8)
        c.field("rnew")[:] = 0.0
        c.load(r, vr, vt, sm, binind=binind)
        c.calc_potential_new()
        c.upload()
        for _ in range(3):
            c.dev_step(seed=5, fused=fused)
        c.dev_compute_energy()          # bring E, J of the fused run up to date for the comparison
        c.download()
        n = c.clus.N_MAX
        outs.append((n, c.star[:n + 1].copy(), c.binary.copy(), c.star_r[:n + 2].copy(), c.star_m[:n + 2].copy(), c.star_phi[:n + 2].copy(), c.N_b_local))
        c.close()
    a, b = outs
    assert a[0] == b[0] and a[6] == b[6]
    assert np.array_equal(a[1], b[1]), "star[]"
    assert np.array_equal(a[2], b[2]), "binary[]"
    for x, y, name in zip(a[3:6], b[3:6], ("star_r", "star_m", "star_phi")):
        cm.assert_bits_equal(x, y, name)


@pytest.mark.parametrize("kind,N,masses,cenma,bh", [("plummer", 1000, "equal", 0.0, False), ("plummer", 300000, "imf", 0.0, False),
                                                    ("king", 1000000, "equal", 0.0, False), ("plummer", 200000, "imf", 0.004, True)])
def test_lagrange_radii_bit_exact(kind, N, masses, cenma, bh):
    """comp_mass_percent on the resident arrays: enclosed-mass and the four cumulative kinetic sums carry the reference's
    sequential roundings, so radii, star counts, mean masses, densities and sums equal the oracle's bit for bit -- before and after a step
    of the resident pipeline; thresholds the cluster never reaches stay untouched; MASS_PC_BH_INCLUDE with a central mass"""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 1, masses=masses)
    pcs = np.array([1e-5, 1e-4, 0.001, 0.005, 0.01, 0.05, 0.1, 0.2, 0.3, 0.5, 0.5 + 1e-9, 0.7, 0.9, 0.99, 0.999999, 1.2])
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.cenma.m = cenma * N
    c.calc_potential_new()
    c.upload()
    for step in range(2):
        filled = c.dev_comp_mass_percent(pcs, bh_include=bh, minimum_r=1e-7, initial_total_mass=0.93)
        c.download()
        n = c.clus.N_MAX
        want = ol.comp_mass_percent(c.star_r[:n + 2].copy(), c.star_m[:n + 2].copy(), n, N, c.field("vr")[1:n + 1].copy(), c.field("vt")[1:n + 1].copy(),
                                    pcs, c.Mtotal, cenma_m=c.cenma.m, initial_total_mass=0.93, bh_include=bh, minimum_r=1e-7)
        assert filled == want["filled"] == len(pcs) - 1
        for f in ol.MASS_PC_FIELDS:
            cm.assert_bits_equal(getattr(c, f), want[f], f"{f} (step {step})")
        assert c.mass_r[-1] == -1.0 and np.all(np.diff(c.mass_r[:-1]) >= 0)
        cm.assert_bits_equal(c.mass_pc_totals, want["totals"], "totals")
        if bh:
            assert c.mass_r[0] == 1e-7 and c.no_star_r[0] == 0     # the central mass alone exceeds the first thresholds
        c.compute_energy(); c.upload()
        c.dev_step(seed=3)
    c.close()


@pytest.mark.parametrize("kind,N,masses,tidal", [("plummer", 50000, "equal", None), ("king", 800000, "equal", None), ("plummer", 120000, "imf", 3.0)])
def test_energy_scheme_lookups_at_new_radii_bit_exact(kind, N, masses, tidal):
    """SURVEY 8f N1 on the resident arrays: energy_conservation2 (Uoldrnew = potential(rnew) + MPI_PHI_S) and ComputeIntermediateEnergy
    (EI where rnew < 1e6, then the transfer) after an orbit step with recorded draws, against the oracle: every potential() lookup goes
    through the r ladder (3 levels at N = 800000) and must return the reference's interpolation bit for bit; escaped stars (tidal cut) keep
    their EI and get Uoldrnew = potential(SF_INFINITY) + self term"""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 7, masses=masses)
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    Rt = tidal if tidal is not None else pot["Rtidal"]
    want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vr, vt, Rt, seed=3)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
    g = np.arange(1, N + 1)
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.star_phi[:N + 2] = pot["phi"][:N + 2]
    c.field("E")[1:N + 1], c.field("J")[1:N + 1] = E, J
    c.field("EI")[1:N + 1] = -7.0
    c.Rtidal = Rt
    c.upload()
    c.dev_new_orbits(draws=want["draws"], draw_off=want["draw_off"])
    c.dev_energy_conservation2()
    c.dev_compute_intermediate_energy()
    c.download()
    cm.assert_bits_equal(c.field("Uoldrnew")[1:N + 1], ol.energy_conservation2(t, g, want["rnew"]), "Uoldrnew")
    cm.assert_bits_equal(c.field("EI")[1:N + 1], ol.intermediate_energy(t, g, want["rnew"], vr, vt, np.full(N, -7.0)), "EI")
    if tidal is not None:
        esc = want["rnew"] >= 1.0e6
        assert esc.sum() > 10 and np.all(c.field("EI")[1:N + 1][esc] == -7.0)
    cm.assert_bits_equal(c.field("r")[1:N + 1], want["rnew"], "r <- rnew")
    cm.assert_bits_equal(c.field("rOld")[1:N + 1], r, "rOld")
    cm.assert_bits_equal(c.field("vr")[1:N + 1], want["vrnew"], "vr <- vrnew")
    c.close()


def _sv3_cluster(kind, N, masses, first=1, nl=None, seed=0, **kw):
    """a sorted cluster with its potential on the device, the local stars first .. first + nl - 1 carrying cm.sv3_fields"""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 7, masses=masses)
    pot = cm.oracle_potential(r, sm)
    nl = nl or N
    f = {k: v[first - 1:first - 1 + nl].copy() for k, v in cm.sv3_fields(r, vr, vt, pot["phi"][1:N + 1], seed=seed, **kw).items()}
    if first + nl - 1 < N:                                   # a shard: its last star feeds the carry too, so the shard keeps a leftover
        f["interacted"][-1], f["vrold"][-1], f["vtold"][-1] = 0, 0.0, 0.0
        f["Uoldrold"][-1] = f["Uoldrnew"][-1] + pot["phi"][first + nl - 1] * 0.3
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.star_phi[:N + 2] = pot["phi"][:N + 2]
    rng = np.random.default_rng(seed + 1)
    c.star[1:nl + 1] = rng.integers(0, 256, (nl, c.star.shape[1]), dtype=np.uint8)     # every other byte of the records is noise
    for k, v in f.items():
        c.field(k)[1:nl + 1] = v
    c.clus.N_MAX_NEW, c.Start[0], c.End[0] = nl, first, first + nl - 1
    c.bin_slots_in_use = 0                                   # binind is noise here; no binary record travels
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
    return c, t, f, np.arange(first, first + nl)


def _only_these_fields_changed(before, c, nl, names):
    """every byte of star[1..nl] outside the named 8-byte fields is what it was"""
    from cmc_cosmic_b200 import layout
    mask = np.ones(c.star.shape[1], bool)
    for k in names:
        mask[layout.STAR_OFFSETS[k]:layout.STAR_OFFSETS[k] + 8] = False
    assert np.array_equal(before[1:nl + 1][:, mask], c.star[1:nl + 1][:, mask])


@pytest.mark.parametrize("kind,N,masses,first,nl", [("plummer", 50000, "imf", 1, None), ("king", 800000, "equal", 1, None), ("plummer", 4096, "equal", 1, None),
                                                    ("plummer", 3000, "imf", 1, None), ("plummer", 120000, "imf", 30001, 50001)])
def test_energy_scheme_velocity_correction_bit_exact(kind, N, masses, first, nl):
    """SURVEY 8f N1, the last potential() loop of the energy scheme, on the resident arrays: energy_conservation1 and
    energy_conservation3 -> set_velocities3 (reference cmc_utils.c:2153-2174, 254-363).  The excess-energy carry the reference runs serially
    from star to star is decomposed into per-star, per-block and one short serial pass (csrc/velocities_core.h) and must reproduce every
    rounding of the serial loop: chains inside a 1024-star block, across many blocks, a leftover at the end of the shard, a shard that
    starts in the middle of the cluster; then passes of the second loop from a received excess.  Nothing else in the records changes."""
    c, t, f, g = _sv3_cluster(kind, N, masses, first, nl, seed=N)
    nl = len(g)
    c.upload()
    before = c.star.copy()
    c.dev_energy_conservation1()
    c.download()
    for k, v in ol.energy_conservation1(t, g, f["vr"], f["vt"]).items():
        cm.assert_bits_equal(c.field(k)[1:nl + 1], v, k)
    _only_these_fields_changed(before, c, nl, ("vrold", "vtold", "Uoldrold"))
    for k, v in f.items():                                  # the scheme's inputs (energy_conservation1 just overwrote three of them)
        c.field(k)[1:nl + 1] = v
    c.upload()
    before = c.star.copy()
    left = c.dev_energy_conservation3()
    c.download()
    want, flags, left0 = ol.set_velocities3_first(t, g, f)
    assert (flags == 0).sum() > 3 and left0 > 0
    for k in ol.SV3_IO:
        cm.assert_bits_equal(c.field(k)[1:nl + 1], want[k], k)
    assert left == left0 and c.Eoops == 0.0
    _only_these_fields_changed(before, c, nl, ol.SV3_IO)
    for Ein in (left0 * 3.0 + 0.01, 0.0, 1e-9, float(np.sum(0.5 * t.madhoc * (want["vr"] ** 2 + want["vt"] ** 2)))):   # the last: more than the shard can absorb
        e0 = ol.set_velocities3_second(t, g, f["interacted"], flags, Ein, want)
        e1 = c.dev_set_velocities3_pass(Ein)
        c.download()
        assert e1 == e0
        for k in ol.SV3_IO:
            cm.assert_bits_equal(c.field(k)[1:nl + 1], want[k], k + f" after a pass from {Ein}")
    c.close()


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("kind,N,masses", [("plummer", 60000, "imf"), ("king", 300000, "equal")])
def test_whole_timestep_with_energy_scheme_bit_exact(kind, N, masses, fused):
    """The hot path inside the reference's timestep order (src/cmc/cmc.c:302-353), everything on the resident arrays with no host round trip:
    energy_conservation1 -> new_orbits_calculate -> energy_conservation2 -> ComputeIntermediateEnergy -> qsorts_new -> calc_potential_new
    -> energy_conservation3 (set_velocities3) -> ComputeEnergy, against the oracle run in the same order, bit for bit."""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 3, masses=masses)
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
    g = np.arange(1, N + 1)
    o1 = ol.energy_conservation1(t, g, vr, vt)
    want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vr, vt, pot["Rtidal"], seed=3)
    Uoldrnew = ol.energy_conservation2(t, g, want["rnew"])
    order = np.argsort(want["rnew"], kind="stable")
    n2 = int((want["rnew"] < cm.SF_INFINITY).sum())
    order = order[:n2]
    pot2 = cm.oracle_potential(want["rnew"][order], sm[order], N_STAR=N)
    t2 = ol.tables(pot2["star_r"], pot2["star_m"], pot2["phi"], n2, N, 0.0)
    f = dict(interacted=np.zeros(n2, np.int64), rOld=r[order], vrold=o1["vrold"][order], vtold=o1["vtold"][order], Uoldrold=o1["Uoldrold"][order],
             Uoldrnew=Uoldrnew[order], vr=want["vrnew"][order], vt=want["vtnew"][order], vrnew=want["vrnew"][order], vtnew=want["vtnew"][order])
    io, flags, left0 = ol.set_velocities3_first(t2, np.arange(1, n2 + 1), f)
    E2, J2 = cm.energies(pot2["phi"][1:n2 + 1], want["rnew"][order], io["vr"], io["vt"])

    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.star_phi[:N + 2] = pot["phi"][:N + 2]
    c.field("E")[1:N + 1], c.field("J")[1:N + 1] = E, J
    c.Rtidal = pot["Rtidal"]
    c.upload()
    if fused:       # the same passes, those that read the same data merged (Cluster.dev_timestep)
        c.dev_timestep(draws=want["draws"], draw_off=want["draw_off"])
    else:
        c.dev_energy_conservation1()
        c.dev_new_orbits(draws=want["draws"], draw_off=want["draw_off"])
        c.dev_energy_conservation2()
        c.dev_compute_intermediate_energy()
        c.dev_qsorts_new()
        c.dev_calc_potential_new()
        assert c.dev_energy_conservation3() == left0
        c.dev_compute_energy()
    c.download()
    assert c.clus.N_MAX == n2
    cm.assert_bits_equal(c.star_phi[:n2 + 2], pot2["phi"][:n2 + 2], "phi")
    assert np.array_equal(c.field("id")[1:n2 + 1], order + 1)
    for k in ("rOld", "vrold", "vtold", "Uoldrold", "Uoldrnew"):
        cm.assert_bits_equal(c.field(k)[1:n2 + 1], f[k], k)
    cm.assert_bits_equal(c.field("EI")[1:n2 + 1], ol.intermediate_energy(t, g, want["rnew"], vr, vt, np.zeros(N))[order], "EI")
    for k in ol.SV3_IO:
        cm.assert_bits_equal(c.field(k)[1:n2 + 1], io[k], k)
    cm.assert_bits_equal(c.field("E")[1:n2 + 1], E2, "E")
    cm.assert_bits_equal(c.field("J")[1:n2 + 1], J2, "J")
    # the scheme does what it is for: the corrected speeds conserve the total energy of the step far better than the uncorrected ones
    m2 = sm[order] / N
    def etot(vr_, vt_):
        return float(np.sum(m2 * (0.5 * (vr_ ** 2 + vt_ ** 2) + 0.5 * pot2["phi"][1:n2 + 1])))
    m1 = sm / N
    e_before = float(np.sum(m1 * (0.5 * (vr ** 2 + vt ** 2) + 0.5 * pot["phi"][1:N + 1]))) - float(np.sum((m1 * E)[want["rnew"] >= cm.SF_INFINITY]))
    assert abs(etot(io["vr"], io["vt"]) - e_before) < abs(etot(f["vr"], f["vt"]) - e_before) + 1e-9
    c.close()


def test_energy_scheme_velocity_correction_edge_cases():
    """no star takes part; no source at all (nothing to carry); every star a source; a failed potential() bracket is reported"""
    N = 5000
    for case in ("none", "no_source", "all_sources"):
        c, t, f, g = _sv3_cluster("plummer", N, "equal", seed=5, src_frac=0.0 if case == "no_source" else 0.004, long_chains=False)
        if case == "none":
            f["interacted"][:] = 1
        elif case == "no_source":
            f["vrold"], f["vtold"], f["Uoldrold"] = f["vr"] * 3.0, f["vt"] * 3.0, f["Uoldrnew"].copy()   # plenty of kinetic energy: every vnew2 > 0
            f["rOld"] = c.star_r[1:N + 1].copy()
        else:
            f["vrold"], f["vtold"] = f["vr"] * 0.0, f["vt"] * 0.0
            f["Uoldrold"] = f["Uoldrnew"] + t_phi(t, g) * 5.0
        for k, v in f.items():
            c.field(k)[1:N + 1] = v
        c.upload()
        left = c.dev_energy_conservation3()
        c.download()
        want, flags, left0 = ol.set_velocities3_first(t, g, f)
        if case == "no_source":
            assert left0 == 0.0 and flags.sum() == (f["interacted"] == 0).sum()
        if case == "all_sources":
            assert flags.sum() == 0 and left0 > 0
        for k in ol.SV3_IO:
            cm.assert_bits_equal(c.field(k)[1:N + 1], want[k], f"{case}: {k}")
        assert left == left0
        c.close()


def t_phi(t, g):
    import ctypes
    return np.ctypeslib.as_array(ctypes.cast(t.phi, ctypes.POINTER(ctypes.c_double)), shape=(t.N_MAX + 2,))[g].copy()


@pytest.mark.parametrize("treatment,orbit_r", [(1, 6.0), (0, 6.0), (1, 1.0e6)])
def test_tidally_strip_stars_bit_exact(treatment, orbit_r):
    """tidally_strip_stars (reference cmc_evolution_thr.c:443-645) on the resident records after an orbit step: the criterion on the device,
    the reference's loop and per-star sums in the host mirror, against the oracle's criterion + loop (pinned against the compiled
    reference in test_oracle_vs_ref.py::test_tidally_strip_stars): the same stars in the same passes, the same sums bit for bit, and
    potential() at arbitrary radii (cmcb200_dev_potential)"""
    import ctypes
    N, nb = 60000, 2500
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=41, masses="imf")
    pot = cm.oracle_potential(r, sm, orbit_r=orbit_r)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vr, vt, pot["Rtidal"], seed=3)
    rng = np.random.default_rng(6)
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    bvals = {k: np.concatenate([[0.0], rng.uniform(lo, hi, nb)]) for k, lo, hi in (("m1", 0.1, 5.0), ("m2", 0.1, 5.0), ("a", 1e-6, 1e-3),
                                                                                   ("Eint1", 0.0, 1e-7), ("Eint2", 0.0, 1e-7))}
    Eint = rng.uniform(0.0, 1e-7, N)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
    f = dict(r_apo=want["r_apo"].copy(), E=E, J=J, Eint=Eint, rnew=want["rnew"].copy(), vrnew=want["vrnew"].copy(), vtnew=want["vtnew"].copy())
    tml0 = float(np.sum((sm / N)[want["rnew"] >= cm.SF_INFINITY]))          # what the orbit step removed
    s = dict(Mtotal=pot["Mtotal"], TidalMassLoss=tml0, OldTidalMassLoss=0.0, Eescaped=0.25, Jescaped=0.5, Eintescaped=0.0, Ebescaped=0.0)
    ol.tidally_strip_loop(lambda thr: ol.tidally_strip_select(treatment, thr, f), lambda x: ol.lib().orc_potential(ctypes.byref(t), x),
                          lambda i: pot["star_m"][i], lambda i: dict(E=E[i - 1], J=J[i - 1], Eint=Eint[i - 1], binind=int(binind[i - 1])),
                          lambda k: {q: bvals[q][k] for q in bvals}, s, N, orbit_r, treatment, 0.11)

    c = Cluster(N, n_star_dim=N + 64, n_bin_dim=nb + 8, orbit_r=orbit_r)
    c.load(r, vr, vt, sm, binind=binind)
    c.star_phi[:N + 2] = pot["phi"][:N + 2]
    for k in ("E", "J", "Eint"):
        c.field(k)[1:N + 1] = f[k] if k != "E" else E
    for k in ("r_apo", "rnew", "vrnew", "vtnew"):
        c.field(k)[1:N + 1] = want[k]
    for k, v in bvals.items():
        c.bfield(k)[:nb + 1] = v
    c.Mtotal, c.TidalMassLoss, c.OldTidalMassLoss, c.Eescaped, c.Jescaped = pot["Mtotal"], tml0, 0.0, 0.25, 0.5
    c.upload()
    probe = np.concatenate([rng.uniform(r[0] * 0.5, r[-1] * 2.0, 5000), r[:50], [0.0, 1e9]])
    cm.assert_bits_equal(c.dev_potential(probe), np.array([ol.lib().orc_potential(ctypes.byref(t), float(x)) for x in probe]), "potential()")
    before = c.star.copy()
    stripped = c.dev_tidally_strip_stars(tidal_treatment=treatment, GAMMA=0.11)
    c.download()
    assert stripped == s["stripped"]
    for k in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(c.field(k)[1:N + 1], f[k], k)
    _only_these_fields_changed(before, c, N, ("rnew", "vrnew", "vtnew"))
    for k in ("Rtidal", "TidalMassLoss", "DTidalMassLoss", "Eescaped", "Jescaped", "Eintescaped", "Ebescaped", "Etidal"):
        assert getattr(c, k) == s[k], (k, getattr(c, k), s[k])
    if orbit_r < 100:
        assert s["passes"] >= 2 and len(stripped) > 5 and (treatment == 0 or (binind[np.array(stripped) - 1] > 0).sum() > 2)
    else:
        assert not stripped
    c.close()


@pytest.mark.parametrize("cenma,N", [(0.0, 60000), (25.0, 60000), (0.0, 1500000)])
def test_compute_energy_totals_bit_exact(cenma, N):
    """The second loop of ComputeEnergy (reference cmc_utils.c:619-655) on the resident arrays: four sequentially summed energies through
    the exact scan (the potential sum takes two additions per star when there is a central mass), singles, binaries in use and
    destroyed binaries, against the oracle (pinned against the compiled reference in test_oracle_vs_ref.py::test_compute_energy_totals)"""
    nb = N // 20
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=35, masses="imf")
    rng = np.random.default_rng(10)
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    bins = {k: np.concatenate([[0.0], rng.uniform(lo, hi, nb)]) for k, lo, hi in (("m1", 0.1, 5.0), ("m2", 0.1, 5.0), ("a", 1e-6, 1e-3),
                                                                                  ("Eint1", 0.0, 1e-7), ("Eint2", 0.0, 1e-7))}
    bins["inuse"] = np.concatenate([[0], (rng.uniform(size=nb) < 0.9)]).astype(np.int32)
    Eint = rng.uniform(0.0, 1e-7, N)
    c = Cluster(N, n_star_dim=N + 64, n_bin_dim=nb + 8)
    c.load(r, vr, vt, sm, binind=binind)
    c.field("Eint")[1:N + 1] = Eint
    for k, v in bins.items():
        c.bfield(k)[:nb + 1] = v
    c.cenma.m = cenma
    c.calc_potential_new()
    c.upload()
    c.Eescaped, c.Ebescaped, c.Eintescaped = -0.001, -0.002, 0.0005
    et = c.dev_compute_energy_totals()
    t = ol.tables(c.star_r[:N + 2].copy(), c.star_m[:N + 2].copy(), c.star_phi[:N + 2].copy(), N, N, c.cenma.m)
    K, P, Ei, Eb = ol.compute_energy_totals(t, 1, vr, vt, Eint, binind, bins)
    assert (et.K, et.P, et.Eint, et.Eb) == (K, 0.5 * P, Ei, Eb), ((et.K, et.P, et.Eint, et.Eb), (K, 0.5 * P, Ei, Eb))
    assert et.tot == (K + 0.5 * P + Ei + Eb) + (0.0 + -0.001 + -0.002 + 0.0005)
    assert K > 0 and P < 0 and Ei > 0 and Eb < 0
    c.close()


@pytest.mark.parametrize("N", [30000, 1200000])
def test_comp_multi_mass_percent_bit_exact(N):
    """comp_multi_mass_percent (reference cmc_utils.c:769-861) on the resident tables: one exact sequential-sum chain per mass bin
    (zero terms for the stars of other bins), two binary searches per threshold, the linear interpolation restated -- against the oracle,
    which is pinned against the compiled reference (test_oracle_vs_ref.py::test_comp_multi_mass_percent): a one-star bin, an empty bin (its row
    stays untouched), fractions 0 and 1, no sequential fall-back of the exact scan."""
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=37, masses="imf")
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.calc_potential_new()
    c.upload()
    q = np.quantile(sm, [0.0, 0.3, 0.6, 0.9, 0.99, 0.9999, 1.0])
    bins = np.concatenate([q[:-1] * [0.999, 1, 1, 1, 1, 1], [q[-2] * 1.0000001, q[-1] * 1.5, q[-1] * 2.0]])
    pcs = np.array([0.0, 0.001, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.99, 1.0])
    got = c.dev_comp_multi_mass_percent(bins, 1.0, pcs)
    assert c.potential_stats()[1] == 0, "sequential fall-back in a per-bin chain"
    want = ol.comp_multi_mass_percent(c.star_r[:N + 2].copy(), c.star_m[:N + 2].copy(), N, bins, 1.0, pcs)
    cm.assert_bits_equal(got, want, "multi_mass_r")
    assert (want[-1] == -1.0).all() and (want[:4, 1:] > 0).all()
    # a second call overwrites only the rows of bins that hold stars
    got[:] = -7.0
    got2 = c.dev_comp_multi_mass_percent(bins, 1.0, pcs)
    assert (got2[-1] == -7.0).all() and np.array_equal(got2[:-1].view(np.uint64)[want[:-1] != -1.0], want[:-1].view(np.uint64)[want[:-1] != -1.0])
    c.close()


def test_orbit_step_where_find_root_vr_fires_bit_exact():
    """SURVEY 8a a20 / a21 on the device: calc_orbit_new's Brent fallback find_root_vr and the helpers under it, on a cluster in which it fires
    hundreds of times (common.brent_cluster), host-buffer form and resident form, against the oracle (pinned against the compiled reference
    on the same cluster in test_oracle_vs_ref.py::test_orbit_step_where_find_root_vr_fires)"""
    N = 20000
    r, vr, vt, sm, pot, E, J, planted = cm.brent_cluster(N)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N)
    draws = ol.Taus113(3).draws(200 * N)
    want = ol.get_positions(t, np.arange(1, N + 1), E, J, vr, vt, pot["Rtidal"], draws)
    used = draws[:want["draw_off"][-1]]
    ok = want["status"] == ol.ST_OK
    assert want["n_brent"] > 150 and (ok & planted).sum() > 100
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.calc_potential_new()
    cm.assert_bits_equal(c.star_phi[:N + 2], pot["phi"][:N + 2], "phi")
    c.field("E")[1:N + 1], c.field("J")[1:N + 1] = E, J
    for form in ("host", "resident"):
        if form == "host":
            status = c.new_orbits_calculate(draws=used, draw_off=want["draw_off"])
            kmin, kmax, res = c.last_kmin, c.last_kmax, c.last_orbit
        else:
            c.upload()
            res = c.dev_new_orbits(draws=used, draw_off=want["draw_off"], want_status=True)
            status, kmin, kmax = c.dev_read_status()
            c.download()
        assert int(res.n_brent) == want["n_brent"], (form, int(res.n_brent), want["n_brent"])
        assert np.array_equal(status, want["status"]) and np.array_equal(kmin[ok], want["kmin"][ok]) and np.array_equal(kmax[ok], want["kmax"][ok]), form
        for k in ("rnew", "vrnew", "vtnew"):
            cm.assert_bits_equal(c.field(k)[1:N + 1], want[k], f"{form}: {k}")
        for k in ("r_peri", "r_apo", "X"):
            cm.assert_bits_equal(c.field(k)[1:N + 1][ok], want[k][ok], f"{form}: {k}")
    c.close()


def test_compute_energy_totals_chains_of_zeros_stay_parallel():
    """A cluster without stellar evolution and without binaries: every term of the Eint and Eb chains is exactly 0.  fl(y + 0) = y, so those
    elements are regular -- no event, and above all no fall-back to the single-thread chain (which at N = 1e7 would take seconds)."""
    N = 400000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=36)
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.calc_potential_new()
    c.upload()
    et = c.dev_compute_energy_totals()
    ev, fb = c.potential_stats()
    assert fb == 0, f"{fb} sequential fall-backs"
    t = ol.tables(c.star_r[:N + 2].copy(), c.star_m[:N + 2].copy(), c.star_phi[:N + 2].copy(), N, N, 0.0)
    z = np.zeros(N)
    bins = {k: np.zeros(1) for k in ("m1", "m2", "a", "Eint1", "Eint2")}
    bins["inuse"] = np.zeros(1, np.int32)
    K, P, Ei, Eb = ol.compute_energy_totals(t, 1, vr, vt, z, np.zeros(N, np.int64), bins)
    assert (et.K, et.P, et.Eint, et.Eb) == (K, 0.5 * P, Ei, Eb) and Ei == 0.0 and Eb == 0.0
    c.close()


def test_tidally_strip_stars_uses_the_resident_masses_after_a_sort():
    """On the resident pipeline the host copy of star_m is not refreshed by the sort: the mass of a stripped star must come from the device
    table.  Two resident timesteps with stripping of an IMF cluster; afterwards the escaper sums equal the ones recomputed from the
    downloaded records of the stripped stars (their .m is star_m[g] at the time of the strip, stored by the transfer before the sort)."""
    N = 30000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=91, masses="imf")
    c = Cluster(N, n_star_dim=N + 64, orbit_r=6.0)
    c.load(r, vr, vt, sm)
    c.calc_potential_new()
    c.compute_energy()
    c.upload()
    total = 0
    for step in range(3):
        c.OldTidalMassLoss = c.TidalMassLoss
        c.dev_energy_conservation1()
        c.dev_new_orbits(seed=7)
        c.dev_energy_conservation2()
        e0, d0 = c.Eescaped, c.TidalMassLoss
        stripped = c.dev_tidally_strip_stars(tidal_treatment=1)
        # the records of the stripped stars, straight from the device (host star_m is deliberately stale from step 1 on)
        c.download()
        idx = np.array(stripped, np.int64)
        if len(idx):
            m_dev, E = c.star_m[idx], c.field("E")[idx]
            # one pass at a time is not recoverable from here; the sums over all passes are order-sensitive only at rounding level
            assert np.isclose(c.Eescaped - e0, float(np.sum(E * m_dev / N)), rtol=1e-9, atol=1e-15)
            assert np.isclose(c.TidalMassLoss - d0, float(np.sum(m_dev / N)), rtol=1e-9, atol=1e-15)
            total += len(idx)
        c.dev_compute_intermediate_energy()
        c.dev_qsorts_new()
        c.dev_calc_potential_new()
        c.dev_energy_conservation3()
        c.dev_compute_energy()
        c.tcount += 1
        # make the host table useless on purpose: the next strip must not read it
        c.star_m[:] = -1.0e300
    assert total > 5
    c.close()


def test_lagrange_radii_of_a_run_agree_statistically_with_the_reference_run():
    """north_star: "Lagrangian radii from a full run must statistically agree with the reference run".  Six timesteps of the path
    (orbit step -> r <- rnew -> sort -> potential -> E, J) from the same Plummer state: the CUDA pipeline with its production draws
    (Philox) against the oracle with the reference's taus113 stream.  Different random numbers, same physics: the Lagrange radii of the
    two final states must agree within the sampling noise of N = 1e5 stars, and both stay on the Plummer profile."""
    N, steps = 100000, 6
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=71)
    pcs = np.array([0.05, 0.1, 0.25, 0.5, 0.75, 0.9])
    # reference run (oracle)
    ro, vro, vto = r.copy(), vr.copy(), vt.copy()
    for s in range(steps):
        pot = cm.oracle_potential(ro, sm)
        E, J = cm.energies(pot["phi"][1:N + 1], ro, vro, vto)
        w = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vro, vto, pot["Rtidal"], seed=1000 + s)
        assert np.all(w["status"] <= ol.ST_CIRCULAR)          # nobody leaves: N stays fixed, the comparison is star for star in number
        o = np.argsort(w["rnew"], kind="stable")
        ro, vro, vto = w["rnew"][o], w["vrnew"][o], w["vtnew"][o]
    pot = cm.oracle_potential(ro, sm)
    want = ol.comp_mass_percent(pot["star_r"], pot["star_m"], N, N, vro, vto, pcs, pot["Mtotal"])
    # CUDA run
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.calc_potential_new()
    c.upload()
    for s in range(steps):
        c.dev_step(seed=5)
    c.dev_comp_mass_percent(pcs)
    assert c.clus.N_MAX == N
    a = 3 * np.pi / 16
    plummer = a / np.sqrt(pcs ** (-2.0 / 3.0) - 1.0)
    # sampling noise of a Lagrange radius: dr/r ~ sqrt(f (1 - f) / N) / (f dln M / dln r); a flat 1.5 % covers 4 sigma at every fraction here
    assert np.allclose(c.mass_r, want["mass_r"], rtol=0.015), (c.mass_r, want["mass_r"])
    assert np.allclose(c.mass_r, plummer, rtol=0.03) and np.allclose(want["mass_r"], plummer, rtol=0.03)
    assert np.allclose(c.ke_rad_r + c.ke_tan_r, want["ke_rad_r"] + want["ke_tan_r"], rtol=0.02)   # enclosed kinetic energy at the same fractions
    c.close()


def _cluster_with_new_stars(N, M, seed, pinned=False):
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=seed)
    pot = cm.oracle_potential(r, sm)
    sr, smt, phi, E, J, vra, vta, g_idx, is_new, rnew_in = cm.with_new_stars(r, vr, vt, sm, pot, M, seed=seed + 1)
    c = Cluster(N, n_star_dim=N + M + 64, pinned=pinned)
    c.load(r, vr, vt, sm)
    c.star_r[:N + 2 + M], c.star_m[:N + 2 + M], c.star_phi[:N + 2 + M] = sr, smt, phi
    c.field("E")[1:N + M + 1], c.field("J")[1:N + M + 1] = E, J
    c.field("vr")[1:N + M + 1], c.field("vt")[1:N + M + 1] = vra, vta
    c.field("rnew")[1:N + M + 1] = rnew_in
    c.field("id")[N + 1:N + M + 1] = N + 1 + np.arange(M)
    c.clus.N_MAX_NEW = N + M                  # the new stars sit behind the old shard; N_MAX, Start / End still describe the old stars
    c.Rtidal = pot["Rtidal"]
    return c, (sr, smt, phi, E, J, vra, vta, g_idx, is_new, rnew_in), pot


def test_orbit_step_with_stars_created_this_step_bit_exact():
    """new stars (collision products in CMC): local index beyond End - Start + 1, global index N_MAX + 2 .. (get_global_idx), position given in
    star[].rnew and located by FindZero_r (cmc_orbit.c:39-41) -- host-buffer and resident form against the oracle, incl. the energy lookups"""
    N, M = 60000, 300
    c, (sr, smt, phi, E, J, vra, vta, g_idx, is_new, rnew_in), pot = _cluster_with_new_stars(N, M, seed=21)
    want = cm.oracle_orbits(sr, smt, phi, N, N, E, J, vra, vta, pot["Rtidal"], seed=8, g_idx=g_idx, is_new=is_new, rnew_in=rnew_in)
    assert (want["status"][N:] == ol.ST_OK).sum() > M // 2
    status = c.new_orbits_calculate(draws=want["draws"], draw_off=want["draw_off"])
    assert np.array_equal(status, want["status"])
    assert np.array_equal(c.last_kmin, want["kmin"]) and np.array_equal(c.last_kmax, want["kmax"])
    for f in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(c.field(f)[1:N + M + 1], want[f], f)
    host_form = c.star[:N + M + 1].copy()
    c.field("rnew")[1:N + M + 1] = rnew_in
    c.upload()
    c.dev_new_orbits(draws=want["draws"], draw_off=want["draw_off"])
    c.dev_energy_conservation2()
    c.download()
    for f in ("rnew", "vrnew", "vtnew", "X", "r_peri", "r_apo"):
        cm.assert_bits_equal(c.field(f)[1:N + M + 1], star_field(host_form, f)[1:N + M + 1], f"resident {f}")
    t = ol.tables(sr, smt, phi, N, N, 0.0)
    cm.assert_bits_equal(c.field("Uoldrnew")[1:N + M + 1], ol.energy_conservation2(t, g_idx, want["rnew"]), "Uoldrnew")
    c.close()


def test_step_with_sort_keys_written_by_the_orbit_kernels():
    """cmcb200_dev_step lets the orbit kernels write the radix sort's keys (key_from_double(rnew)) so that the sort skips its pass over the
    records; the three-call form extracts them from the records.  Same arrays after three steps, byte for byte, on a cluster with stars
    created this step (keys of stars placed by FindZero_r) and stars with v_r == 0 from the second step on (queued linear scans); and with the key emission switched off (CMCB200_STEP_KEYS=0, read when the context is created)."""
    N, M = 90000, 500
    outs = []
    for one_wait, env in ((True, None), (False, None), (True, "0")):
        if env is not None:
            os.environ["CMCB200_STEP_KEYS"] = env
        try:
            c, (sr, smt, phi, E, J, vra, vta, g_idx, is_new, rnew_in), pot = _cluster_with_new_stars(N, M, seed=77)
        finally:
            os.environ.pop("CMCB200_STEP_KEYS", None)
        c.upload()
        for _ in range(3):
            res = c.dev_step(seed=11, one_wait=one_wait)
        c.dev_compute_energy()
        c.download()
        n = c.clus.N_MAX
        outs.append((n, c.star[:n + 1].copy(), c.star_r[:n + 2].copy(), c.star_m[:n + 2].copy(), c.star_phi[:n + 2].copy(), int(res.n_deferred)))
        c.close()
    assert outs[0][0] == N + M, "the new stars joined the cluster"
    for other in outs[1:]:
        assert outs[0][0] == other[0]
        assert np.array_equal(outs[0][1], other[1]), "star[]"
        for x, y, name in zip(outs[0][2:5], other[2:5], ("star_r", "star_m", "star_phi")):
            cm.assert_bits_equal(x, y, name)


def test_pipelined_host_buffer_orbit_step_with_new_stars_in_the_last_chunk():
    """three pipeline chunks, the last one holding old AND new stars (its n_old_local is smaller than its length): against the resident form"""
    N, M = 1100000, 4000
    outs = []
    for mode in ("host", "resident"):
        c, (sr, smt, phi, E, J, vra, vta, g_idx, is_new, rnew_in), pot = _cluster_with_new_stars(N, M, seed=31, pinned=(mode == "host"))
        if mode == "host":
            status = c.new_orbits_calculate(seed=3)
        else:
            c.upload()
            c.dev_new_orbits(seed=3, want_status=True)
            status = c.dev_read_status()[0]
            c.download()
        outs.append((status.copy(), c.star[:N + M + 1].copy(), c.last_orbit.n_draws, c.max_r))
        c.close()
    a, b = outs
    assert np.array_equal(a[0], b[0]) and a[2] == b[2] and a[3] == b[3]
    assert np.array_equal(a[1], b[1]), "star[] differs between the pipelined host-buffer form and the resident form"


def test_empty_and_degenerate_inputs():
    """nothing to do is a valid request for every entry point: no local stars, no thresholds, a one-star cluster"""
    c = Cluster(8, n_star_dim=64)
    r = np.array([0.1, 0.2, 0.3, 0.5, 0.8, 1.3, 2.1, 3.4]); v = np.full(8, 0.1)
    c.load(r, v, v, np.ones(8))
    c.calc_potential_new()
    c.compute_energy()
    c.upload()
    assert c.dev_comp_mass_percent(np.array([])) == 0
    # a rank that owns no stars this step
    n_keep, c.clus.N_MAX_NEW = c.clus.N_MAX_NEW, 0
    c.End[0] = 0
    status = c.new_orbits_calculate(seed=1)
    assert status.size == 0 and c.last_orbit.n_draws == 0 and c.max_r == 0.0
    res = c.dev_new_orbits(seed=1)
    assert res.n_draws == 0
    c.dev_energy_conservation2()
    c.bin_slots_in_use = 0
    c.qsorts_new()
    assert c.clus.N_MAX == 0
    c.close()
    # one star: every table has its sentinels only around a single entry (FindZero_r's kmin == kmax - 1 shortcut, cmc_utils.c:136-138)
    c = Cluster(1, n_star_dim=16)
    c.load(np.array([0.7]), np.array([0.05]), np.array([0.3]), np.array([1.0]))
    c.calc_potential_new()
    want = cm.oracle_potential(np.array([0.7]), np.array([1.0]))
    cm.assert_bits_equal(c.star_phi[:3], want["phi"][:3], "phi of a one-star cluster")
    assert c.dev_clusdyn_calculate() == c.star_r[2]           # the loop adds the only star and stops one index later (cmc_utils.c:1619-1623)
    c.compute_energy()
    E, J = c.field("E")[1:2].copy(), c.field("J")[1:2].copy()
    w = cm.oracle_orbits(want["star_r"], want["star_m"], want["phi"], 1, 1, E, J, np.array([0.05]), np.array([0.3]), want["Rtidal"], seed=2)
    assert w["status"][0] == ol.ST_UNBOUND                    # without its own potential (the self term, :806) a lone star is not bound
    c.upload()
    c.dev_step(seed=2)                                        # ... so the step removes it and the sort trims it
    c.download()
    assert c.clus.N_MAX == 0 and c.star_r[1] == cm.SF_INFINITY
    c.close()


DYN_TOL = 1e-12   # sin / cos / log / pow on the device are CUDA's (<= 2 ulp from glibc's): relative to the pair's velocity / energy scale


def _dyn_cluster(d):
    """cm.dynamics_case on a Cluster: tables of the whole cluster, the records of the local stars with noise in every other byte"""
    N, nl, first, f = d["N"], d["nl"], d["first"], d["f"]
    c = Cluster(N, n_star_dim=N + 64, n_bin_dim=d["nb"] + 8)
    c.load(d["r"], np.zeros(N), np.zeros(N), d["sm"])
    c.star_phi[:N + 2] = d["pot"]["phi"][:N + 2]
    rng = np.random.default_rng(3)
    c.star[1:nl + 1] = rng.integers(0, 256, (nl, c.star.shape[1]), dtype=np.uint8)
    for k, v in f.items():
        c.field({"threebb": "threebb_interacted"}.get(k, k))[1:nl + 1] = v
    if "threebb" not in f:
        c.field("threebb_interacted")[1:nl + 1] = 0
    c.bfield("a")[:d["nb"] + 1] = d["bin_a"]
    c.clus.N_MAX_NEW, c.Start[0], c.End[0] = nl, first, first + nl - 1
    c.bin_slots_in_use = d["nb"]
    return c


@pytest.mark.parametrize("case,first,nl", [("relax_only", 1, None), ("encounters", 1, None), ("no_relaxation", 1, None), ("threebb_skips", 1, None),
                                           ("encounters", 5001, 7001), ("large", 1, None)])
def test_dynamics_apply_pairs_match_oracle(case, first, nl):
    """SURVEY 8f N3, dynamics_apply's pair loop on the resident arrays (reference cmc_dynamics.c:206-542) against the restatement that is pinned
    bit for bit against the compiled reference, same three draws per pair: the list of encounter pairs (indices, type) must be the
    oracle's exactly, Y bit for bit; rperi / W / rcm and the relaxed pairs' new vr, vt, E, J to DYN_TOL (the device's sin / cos / log / pow
    are not glibc's); the counts of the scattering-angle statistics exactly, their sums to rounding; nothing else in the records changes."""
    d = cm.dynamics_case(case, first=first, nl=nl)
    N, nl, cfg = d["N"], d["nl"], d["cfg"]
    t = ol.tables(d["pot"]["star_r"], d["pot"]["star_m"], d["pot"]["phi"], N, N)
    dp = ol.dyn_params(**{k: v for k, v in cfg.items() if k in dict(ol.DynParams._fields_)})
    f0 = {k: v.copy() for k, v in d["f"].items()}
    want = ol.dynamics_apply_pairs(t, first, d["dt"], dp, d["draws"], f0, d["bin_a"])
    c = _dyn_cluster(d)
    c.upload()
    before = c.star.copy()
    sw = dict(RELAXATION=cfg["relaxation"], SS_COLLISION=cfg["ss_collision"], BINSINGLE=cfg["binsingle"], BINBIN=cfg["binbin"],
              TIDAL_CAPTURE=cfg["tidal_capture"], BH_CAPTURE=cfg["bh_capture"], AVEKERNEL=cfg["p"], GAMMA=cfg["gamma"], COLL_FACTOR=cfg["coll_factor"],
              TC_FACTOR=cfg["tc_factor"], WD_TC=cfg["wd_tc"], units_mstar=cfg["units_mstar"], units_l=cfg["units_l"], units_t=cfg["units_t"],
              THREEBODYBINARIES=int("threebb" in d["f"]))
    res = c.dev_dynamics_apply(d["dt"], draws=d["draws"], **sw)
    c.download()
    enc = (want["status"] == 1) | (want["status"] == 2)
    e = res.encounters
    assert np.array_equal(e["k"], want["pair_k"][enc]) and np.array_equal(e["kp"], want["pair_kp"][enc])
    assert np.array_equal(e["type"], want["status"][enc])
    for k in ("rperi", "W", "rcm"):
        np.testing.assert_allclose(e[k], want[k][enc], rtol=DYN_TOL, atol=0, err_msg=k)
    cm.assert_bits_equal(c.field("Y")[1:nl + 1], f0["Y"], "Y")
    vscale = np.sqrt(f0["vr"] ** 2 + f0["vt"] ** 2)
    for k, scale in (("vr", vscale), ("vt", vscale), ("E", np.abs(f0["E"]) + vscale ** 2), ("J", np.abs(f0["J"]) + 1e-300)):
        err = np.abs(c.field(k)[1:nl + 1] - f0[k]) / scale
        assert err.max() <= DYN_TOL, (k, err.max())
    assert res.Nrel == want["Nrel"] and list(res.Nrelbeta) == list(want["Nrelbeta"])
    for got, w in ((res.qaverelbeta, want["qave"]), (res.maverelbeta, want["mave"]), (res.raverelbeta, want["rave"])):
        np.testing.assert_allclose(np.array(got[:]), w, rtol=1e-11)
    if case in ("encounters", "large"):
        assert (e["type"] == 1).sum() > 10 and (e["type"] == 2).sum() > 10
    if cfg["relaxation"]:
        assert res.Nrel > 0 and np.any(c.field("vr")[1:nl + 1] != d["f"]["vr"])
    _only_these_fields_changed(before, c, nl, ("vr", "vt", "E", "J", "Y"))
    c.close()


def test_dynamics_apply_philox_mode_is_statistically_the_reference():
    """production mode (no draw array): Philox draws keyed by (seed, tcount, star).  The relaxation step conserves each pair's momentum and
    kinetic energy whatever the angles are, and two runs with the same seed are identical while another seed differs."""
    d = cm.dynamics_case("relax_only")
    N, nl = d["N"], d["nl"]
    outs = []
    for seed in (5, 5, 6):
        c = _dyn_cluster(d)
        c.upload()
        res = c.dev_dynamics_apply(d["dt"], seed=seed)
        c.download()
        outs.append((c.field("vr")[1:nl + 1].copy(), c.field("vt")[1:nl + 1].copy(), c.field("Y")[1:nl + 1].copy()))
        assert res.Nrel == nl // 2 and res.n_encounters == 0
        c.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert not np.array_equal(outs[0][0], outs[2][0])
    vr, vt, Y = outs[0]
    m = d["sm"][:nl]
    npair = nl // 2
    ke0 = (m * (d["f"]["vr"] ** 2 + d["f"]["vt"] ** 2))[:2 * npair].reshape(-1, 2).sum(1)
    ke1 = (m * (vr ** 2 + vt ** 2))[:2 * npair].reshape(-1, 2).sum(1)
    np.testing.assert_allclose(ke1, ke0, rtol=1e-12)                       # elastic scattering in the pair's centre-of-mass frame
    pr0 = (m * d["f"]["vr"])[:2 * npair].reshape(-1, 2).sum(1)
    pr1 = (m * vr)[:2 * npair].reshape(-1, 2).sum(1)
    np.testing.assert_allclose(pr1, pr0, rtol=0, atol=1e-12 * np.abs(m * np.hypot(d["f"]["vr"], d["f"]["vt"])).max())
    assert 0.45 < Y[::2][:npair].mean() < 0.55 and Y.min() >= 0.0 and Y.max() < 1.0


SIGMA_CASES = [
    # kind, N, masses, p, first, n_local
    ("plummer", 30000, "imf", 20, 1, None),
    ("king", 20000, "equal", 20, 1, None),
    ("plummer", 5000, "imf", 3, 1, None),           # BH_AVEKERNEL
    ("plummer", 200, "imf", 20, 1, None),           # less than one tile
    ("plummer", 41, "equal", 20, 1, None),          # hardly more than one window
    ("plummer", 6155, "imf", 20, 1, None),          # the last tile (10 stars) lies where the window has stopped sliding: its terms reach back beyond si - p
    ("plummer", 60000, "imf", 20, 30001, 2048 * 5 + 8),  # the same for a middle shard (the window keeps sliding into the ghost stars)
    ("plummer", 30, "equal", 20, 1, None),          # the window does not fit: the loop as one thread runs it
    ("plummer", 60000, "imf", 20, 20001, 20000),    # a middle shard: an empty window p stars before its first star, ghost stars on both sides
    ("plummer", 60000, "imf", 20, 40001, 20000),    # the last shard
    ("plummer", 60000, "imf", 20, 1, 20000),        # the first shard
    ("king", 3000000, "imf", 20, 1, None),
    ("plummer", 2000000, "equal", 20, 1, None),
]


@pytest.mark.parametrize("kind,N,masses,p,first,nl", SIGMA_CASES)
def test_calc_sigma_r_bit_exact(kind, N, masses, p, first, nl):
    """SURVEY 8f N3, calc_sigma_r on the resident arrays (reference cmc_dynamics_helper.c:2821-2975) against the restatement that is pinned
    against the compiled reference on 1-4 ranks: sigma_array.r / .sigma and the mean-mass variant, bit for bit -- the running window sums
    carry every rounding of the rank's subtract-then-add updates (csrc/sigma_scan.h), whole clusters and shards with ghost stars; the
    sequential replay of the same loop as a cross-check; nothing in the records changes."""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + p, masses=masses)
    sr, smt = cm.make_tables(r, sm, extra=2 * p)    # a window wider than the cluster reads table entries behind the sentinel: zeros, as after start-up
    nl = N - first + 1 if nl is None else nl
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    lo = first - 1
    rng = np.random.default_rng(1)
    if first > 1 or nl < N:                         # a shard: its records at 1 .. nl
        c.star[1:nl + 1] = rng.integers(0, 256, (nl, c.star.shape[1]), dtype=np.uint8)
        c.field("vr")[1:nl + 1], c.field("vt")[1:nl + 1] = vr[lo:lo + nl], vt[lo:lo + nl]
    c.clus.N_MAX_NEW, c.Start[0], c.End[0] = nl, first, first + nl - 1
    c.bin_slots_in_use = 0
    ghosts = {}
    if first > 1:
        ghosts.update(prev_vr=vr[lo - p:lo], prev_vt=vt[lo - p:lo])
    if lo + nl < N:
        ghosts.update(next_vr=vr[lo + nl:lo + nl + p], next_vt=vt[lo + nl:lo + nl + p])
    c.upload()
    before = c.star.copy()
    for mave in (False, True):
        want_a, want_s = ol.calc_sigma_r(p, sr, smt, N, N, vr, vt, first_global=first, n_local=nl, mave=mave)
        a, sg = c.dev_calc_sigma_r(p, mave=mave, ghosts=ghosts)
        cm.assert_bits_equal(a[1:], want_a[1:], "r / mean mass")
        cm.assert_bits_equal(sg[1:], want_s[1:], "sigma")
        st = c.sigma_stats
        assert st["rerun"] == 0 and st["broken"] == 0 and st["hard_segments"] <= 4, st
        assert st["hard_tiles"] <= max(2, st["tiles"] // 50), st
    if N <= 100000:
        a2, sg2 = c.dev_calc_sigma_r(p, mave=True, ghosts=ghosts, force_sequential=True)
        cm.assert_bits_equal(a2[1:], want_a[1:], "mean mass, one thread")
        cm.assert_bits_equal(sg2[1:], want_s[1:], "sigma, one thread")
    c.download()
    assert np.array_equal(before, c.star)
    c.close()


def test_whole_timestep_from_dynamics_apply_resident():
    """src/cmc/cmc.c:286-353 with nothing on the host: calc_sigma_r -> dynamics_apply's pair loop -> energy_conservation1 -> new_orbits_calculate
    -> ... -> ComputeEnergy on the resident records.  Checked in two legs, because a relaxed velocity that differs in the last bits (the device's
    sin / cos) can flip a rejection try of the orbit step: (1) sigma_array bit for bit and the relaxed state to DYN_TOL against the oracle;
    (2) from the DEVICE's own relaxed state -- the E, J the pair loop stored are what the orbit step reads -- the rest of the timestep bit for
    bit against the oracle chain."""
    kind, N, masses, dt = "plummer", 80000, "imf", 2e-3
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 9, masses=masses)
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
    g = np.arange(1, N + 1)
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.star_phi[:N + 2] = pot["phi"][:N + 2]
    c.field("E")[1:N + 1], c.field("J")[1:N + 1] = E, J
    c.Rtidal = pot["Rtidal"]
    c.upload()
    # leg 1
    a, sg = c.dev_calc_sigma_r()
    want_a, want_s = ol.calc_sigma_r(20, pot["star_r"], pot["star_m"], N, N, vr, vt)
    cm.assert_bits_equal(a[1:], want_a[1:], "sigma_array.r")
    cm.assert_bits_equal(sg[1:], want_s[1:], "sigma_array.sigma")
    draws = np.random.default_rng(4).random(3 * (N // 2))
    res = c.dev_dynamics_apply(dt, draws=draws)
    assert res.n_encounters == 0 and res.Nrel == N // 2
    f0 = dict(vr=vr.copy(), vt=vt.copy(), E=E.copy(), J=J.copy(), Y=np.zeros(N), binind=np.zeros(N, np.int64), se_k=np.zeros(N, np.int32), rad=np.zeros(N))
    ol.dynamics_apply_pairs(t, 1, dt, ol.dyn_params(), draws, f0)
    c.download()
    vr1, vt1, E1, J1 = (c.field(k)[1:N + 1].copy() for k in ("vr", "vt", "E", "J"))
    vs = np.sqrt(vr ** 2 + vt ** 2)
    for k, got, scale in (("vr", vr1, vs), ("vt", vt1, vs), ("E", E1, np.abs(E) + vs ** 2), ("J", J1, np.abs(J) + 1e-300)):
        assert (np.abs(got - f0[k]) / scale).max() <= DYN_TOL, k
    # leg 2: the oracle chain from the device's relaxed state
    o1 = ol.energy_conservation1(t, g, vr1, vt1)
    want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E1, J1, vr1, vt1, pot["Rtidal"], seed=3)
    Uoldrnew = ol.energy_conservation2(t, g, want["rnew"])
    order = np.argsort(want["rnew"], kind="stable")
    n2 = int((want["rnew"] < cm.SF_INFINITY).sum())
    order = order[:n2]
    pot2 = cm.oracle_potential(want["rnew"][order], sm[order], N_STAR=N)
    t2 = ol.tables(pot2["star_r"], pot2["star_m"], pot2["phi"], n2, N, 0.0)
    f = dict(interacted=np.zeros(n2, np.int64), rOld=r[order], vrold=o1["vrold"][order], vtold=o1["vtold"][order], Uoldrold=o1["Uoldrold"][order],
             Uoldrnew=Uoldrnew[order], vr=want["vrnew"][order], vt=want["vtnew"][order], vrnew=want["vrnew"][order], vtnew=want["vtnew"][order])
    io, flags, left0 = ol.set_velocities3_first(t2, np.arange(1, n2 + 1), f)
    E2, J2 = cm.energies(pot2["phi"][1:n2 + 1], want["rnew"][order], io["vr"], io["vt"])
    c.dev_timestep(draws=want["draws"], draw_off=want["draw_off"])
    c.download()
    assert c.clus.N_MAX == n2
    cm.assert_bits_equal(c.star_phi[:n2 + 2], pot2["phi"][:n2 + 2], "phi")
    assert np.array_equal(c.field("id")[1:n2 + 1], order + 1)
    for k in ol.SV3_IO:
        cm.assert_bits_equal(c.field(k)[1:n2 + 1], io[k], k)
    cm.assert_bits_equal(c.field("E")[1:n2 + 1], E2, "E")
    cm.assert_bits_equal(c.field("J")[1:n2 + 1], J2, "J")
    # and the production form: one call, Philox draws, nothing read back in between
    c.dev_timestep(seed=7, tidally_strip=1, dynamics_dt=dt)
    c.download()
    n3 = c.clus.N_MAX
    assert np.all(np.diff(c.field("r")[1:n3 + 1]) >= 0) and len(np.unique(c.field("id")[1:n3 + 1])) == n3
    c.close()


def test_simul_relax_bit_exact():
    """simul_relax_new (reference cmc_dynamics_helper.c:2526-2580) on the resident arrays against the restatement pinned against the
    compiled reference: the same double, whole cluster and a shard"""
    for kind, N, masses, first, nl in (("plummer", 300000, "imf", 1, None), ("king", 50011, "equal", 1, None), ("plummer", 120000, "imf", 30001, 50001),
                                       ("plummer", 50, "equal", 1, None)):
        r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 2, masses=masses)
        pot = cm.oracle_potential(r, sm)
        nl = N - first + 1 if nl is None else nl
        c = Cluster(N, n_star_dim=N + 64)
        c.load(r, vr, vt, sm)
        c.star_phi[:N + 2] = pot["phi"][:N + 2]
        if first > 1 or nl < N:
            c.field("vr")[1:nl + 1], c.field("vt")[1:nl + 1] = vr[first - 1:first - 1 + nl], vt[first - 1:first - 1 + nl]
        c.clus.N_MAX_NEW, c.Start[0], c.End[0] = nl, first, first + nl - 1
        c.upload()
        t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N)
        for theta in (1.0, 1.4142135623730951):
            want = ol.simul_relax_new(t, first, 20, theta, vr[first - 1:first - 1 + nl], vt[first - 1:first - 1 + nl])
            got = c.dev_simul_relax(theta)
            assert np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64), (kind, N, got, want)
        c.close()


def test_lagrange_radii_of_a_full_run_agree_with_the_reference_run():
    """"Lagrangian radii from a full run must statistically agree with the reference run" (north star): a Plummer sphere of 10^5 equal-mass
    stars evolved by two-body relaxation towards core collapse with the configuration of examples/PlummerSphere.ini (CMC_GAMMA 0.1, THETASEMAX 1,
    AVEKERNEL 20, radial tidal criterion) -- EVERY N-sized operation of the main loop on the resident arrays (Cluster.evolve: calc_timestep,
    dynamics_apply, the energy scheme, orbits, sort, potential, Lagrange radii), more than a thousand timesteps without the records visiting
    the host -- against the Lagrange radii of the reference's own example run (docs/.../plummer.lagrad.dat -> tests/golden/plummer_lagrad_run.npz):
    the inner radii contract and the outer ones expand on the reference's clock.  Tolerances: the sampling noise of the mass fraction in both
    runs plus the run-to-run scatter of the collapse."""
    import os
    ref = np.load(os.path.join(cm.GOLDEN, "plummer_lagrad_run.npz"))
    fr, t_ref, r_ref = ref["mass_pc"], ref["t"], ref["r"]
    N = 100000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=77)
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, sm)
    c.calc_potential_new()
    c.compute_energy()
    c.upload()
    before = c.transfer_bytes(reset=True)
    hist = c.evolve(t_end=1.4, max_steps=20000, seed=11, THETASEMAX=1.0, tidally_strip=0, GAMMA=0.1, mass_pc=fr, record_every=5)
    moved = c.transfer_bytes()
    steps = len(hist) * 5
    assert 300 < steps < 20000, steps
    assert moved[0] + moved[1] < 64 * steps * 25 * 8 + (1 << 22), f"the records must stay on the device: {moved} bytes in {steps} steps"
    th = np.array([h[0] for h in hist])
    lag = np.array([h[3] for h in hist])
    sel = [int(np.argmin(np.abs(fr - f))) for f in (0.01, 0.05, 0.1, 0.3, 0.5, 0.7, 0.9)]
    worst = 0.0
    for tq in (0.25, 0.5, 0.75, 1.0, 1.25, 1.38):      # 80 % of the way to core collapse (t = 1.72 in the reference run; the last stretch is a runaway whose onset scatters by a few per cent from run to run)
        w = np.abs(th - tq) <= 0.015
        assert w.sum() >= 1, tq
        mine = lag[w].mean(axis=0)
        want = r_ref[int(np.argmin(np.abs(t_ref - tq)))]
        for j in sel:
            tol = 0.10 if fr[j] <= 0.05 else 0.06
            dev = abs(mine[j] / want[j] - 1)
            worst = max(worst, dev)
            assert dev < tol, f"t = {tq}: r({fr[j]}) = {mine[j]:.4f}, the reference run has {want[j]:.4f}"
    # the evolution itself, not just the initial model: the 10 % radius has contracted, the 90 % radius expanded, like in the reference run
    j10, j90 = sel[2], sel[6]
    assert lag[-1][j10] < 0.75 * lag[0][j10] and lag[-1][j90] > 1.3 * lag[0][j90]
    print(f"full run: {steps} steps to t = {th[-1]:.3f}, N_MAX {N} -> {hist[-1][2]}, worst Lagrange-radius deviation from the reference run {worst:.3f}")
    c.close()


@pytest.mark.parametrize("kind", ["hover_at_power_of_two", "zeros", "wide_range", "steep_trend", "rising_trend", "spikes"])
def test_calc_sigma_r_bit_exact_on_adversarial_inputs(kind):
    """calc_sigma_r on the device where the binade bookkeeping is stressed (tests/common.py: sums hovering on a power of two, chains that reach
    zero, ten decades of terms, trends over dozens of binades, spikes): the serial loop's bits, through whichever path -- tables, replayed
    segments, tiles walked group by group, the one-thread replay"""
    rng = np.random.default_rng(len(kind) * 1000 + 7)
    N = 60000
    m, vr, vt = cm.adversarial_sigma_inputs(kind, N, rng)
    r = np.sort(rng.random(N)) + 0.01
    sr, smt = cm.make_tables(r, m)
    c = Cluster(N, n_star_dim=N + 64)
    c.load(r, vr, vt, m)
    c.upload()
    want_a, want_s = ol.calc_sigma_r(20, sr, smt, N, N, vr, vt, mave=True)
    a, sg = c.dev_calc_sigma_r(20, mave=True)
    cm.assert_bits_equal(a[1:], want_a[1:], f"{kind}: mean mass")
    both_nan = np.isnan(sg[1:]) & np.isnan(want_s[1:])
    cm.assert_bits_equal(np.where(both_nan, 0.0, sg[1:]), np.where(both_nan, 0.0, want_s[1:]), f"{kind}: sigma")
    assert c.sigma_stats["rerun"] == 0, c.sigma_stats
    c.close()


def test_one_wait_step_equals_three_calls():
    """cmcb200_dev_step (orbit step, r <- rnew + sort, potential rebuild queued together, one host wait) against the three calls with a wait each:
    identical records, binaries, tables and scalars over three steps with binaries and escaping stars"""
    N, nb = 120000, 4000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=19, masses="imf")
    rng = np.random.default_rng(6)
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    vr = vr.copy()
    vr[rng.choice(N, 300, replace=False)] *= 4.0            # a few hundred unbound stars: the sort trims them
    out = []
    for one_wait in (True, False):
        c = Cluster(N, n_star_dim=N + 64, n_bin_dim=nb + 8)
        c.load(r, vr, vt, sm, binind=binind)
        c.binary[1:nb + 1] = np.random.default_rng(7).integers(0, 256, (nb, c.binary.shape[1]), dtype=np.uint8)
        c.calc_potential_new()
        c.compute_energy()
        c.upload()
        for step in range(3):
            c.dev_step(seed=5, one_wait=one_wait)
        c.download()
        out.append((c.clus.N_MAX, c.N_b_local, c.Mtotal, c.Rtidal, c.max_r, c.star[:c.clus.N_MAX + 1].copy(), c.binary[:c.N_b_local + 1].copy(),
                    c.star_r[:c.clus.N_MAX + 2].copy(), c.star_m[:c.clus.N_MAX + 1].copy(), c.star_phi[:c.clus.N_MAX + 2].copy()))
        c.close()
    a, b = out
    assert a[:5] == b[:5] and a[0] < N, (a[:5], b[:5])
    for x, y in zip(a[5:], b[5:]):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
