"""The oracle restatement (oracle/cmc_oracle.c) against the UNMODIFIED reference compiled into oracle/_ref/libcmc_ref.so, on seeded
inputs larger and more varied than the committed fixtures.  Skipped where the compiled reference is absent (it is built from
/root/reference by oracle/Makefile and travels to the GPU box prebuilt)."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import oraclelib as ol
import refharness as rh
from cmc_cosmic_b200 import layout

pytestmark = pytest.mark.skipif(not rh.available(), reason="oracle/_ref/libcmc_ref.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def R():
    return rh.Ref()


def test_layout_constants_match_reference_headers(R):
    """sizeof/offsetof of star_t / binary_t in the compiled reference (include/cmc/cmc.h:409-601, :231-390) == the frozen layout the
    kernels address (csrc/star_layout.h, reported by cmcb200_layout) == the numpy views (layout.py)."""
    from cmc_cosmic_b200 import lib
    out = (C.c_long * 16)()
    lib.load().cmcb200_layout(out)
    names = ["r", "vr", "vt", "m", "E", "J", "rnew", "vrnew", "vtnew", "rOld", "X", "r_peri", "r_apo", "binind"]
    assert out[0] == R.sizeof_star == layout.STAR_DTYPE_SIZE == 424
    assert out[1] == R.sizeof_binary == layout.BINARY_DTYPE_SIZE == 464
    for i, f in enumerate(names):
        assert out[2 + i] == R.star_offset(f) == layout.STAR_OFFSETS[f], f
    for f in layout.STAR_OFFSETS:
        assert R.star_offset(f) == layout.STAR_OFFSETS[f], f
    # every offset the kernels use (csrc/star_layout.h)
    import os, re
    hdr = open(os.path.join(os.path.dirname(lib.__file__), "csrc", "star_layout.h")).read()
    by_upper = {k.upper(): k for k in layout.STAR_OFFSETS}
    defs = re.findall(r"#define CMCB_OFF_(\w+) (\d+)", hdr)
    assert len(defs) >= 24
    for name, off in defs:
        assert R.star_offset(by_upper[name]) == int(off), name
    for f, off in layout.BINARY_OFFSETS.items():
        assert R.binary_offset(f) == off, f


@pytest.mark.parametrize("kind,N,masses,cenma", [("plummer", 30000, "equal", 0.0), ("king", 20000, "equal", 0.0), ("plummer", 20000, "imf", 0.0),
                                                 ("plummer", 5000, "imf", 37.5)])
def test_full_step(R, kind, N, masses, cenma):
    """potential -> E,J -> orbit step (taus113 draws recorded) -> transfer -> sort -> potential, reference vs oracle, bit for bit"""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 1, masses=masses)
    R.init(N, n_star_dim=N + 64, idum=4321)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N, cenma_m=cenma)
    R.L.ref_calc_potential_new()
    R.compute_energy()
    sc = R.scalars()
    want = cm.oracle_potential(r, sm, N, cenma_m=cenma)
    cm.assert_bits_equal(want["phi"][:N + 2], R.star_phi()[:N + 2], "phi")
    assert want["Mtotal"] == sc["Mtotal"] and want["Rtidal"] == sc["Rtidal"]
    E, J = R.star_field("E")[1:N + 1].copy(), R.star_field("J")[1:N + 1].copy()
    E2, J2 = cm.energies(want["phi"][1:N + 1], r, vr, vt)
    cm.assert_bits_equal(E2, E, "E")
    cm.assert_bits_equal(J2, J, "J")
    R.L.ref_draw_mode(1)
    R.L.ref_new_orbits_calculate()
    R.L.ref_draw_mode(0)
    draws = R.draws()
    t = ol.tables(want["star_r"], want["star_m"], want["phi"], N, N, cenma)
    out = ol.get_positions(t, np.arange(1, N + 1), E, J, vr, vt, sc["Rtidal"], draws)
    assert int(out["draw_off"][-1]) == len(draws)
    for f in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(out[f], R.star_field(f)[1:N + 1], f)
    ok = out["status"] == ol.ST_OK
    for f in ("X", "r_peri", "r_apo"):
        cm.assert_bits_equal(out[f][ok], R.star_field(f)[1:N + 1][ok], f)
    assert out["max_rad"] == R.scalars()["max_r"]
    # the energy scheme's potential lookups at the new radii: energy_conservation2, ComputeIntermediateEnergy (EI + the transfer)
    R.star_field("EI")[1:N + 1] = -7.0
    R.L.ref_energy_conservation2()
    cm.assert_bits_equal(ol.energy_conservation2(t, np.arange(1, N + 1), out["rnew"]), R.star_field("Uoldrnew")[1:N + 1], "Uoldrnew")
    R.L.ref_compute_intermediate_energy()
    cm.assert_bits_equal(ol.intermediate_energy(t, np.arange(1, N + 1), out["rnew"], vr, vt, np.full(N, -7.0)), R.star_field("EI")[1:N + 1], "EI")
    assert np.array_equal(R.star_field("r")[1:N + 1], out["rnew"]) and np.array_equal(R.star_field("rOld")[1:N + 1], want["star_r"][1:N + 1])
    # sort + second potential (the transfer was ComputeIntermediateEnergy's)
    R.L.ref_qsorts_new()
    R.L.ref_post_sort_comm()
    R.L.ref_calc_potential_new()
    n2 = int(R.scalars()["N_MAX"])
    order = np.argsort(out["rnew"], kind="stable")
    assert n2 == int((out["rnew"] < cm.SF_INFINITY).sum())
    assert np.array_equal(R.star_field("id", np.int64)[1:n2 + 1], order[:n2] + 1)
    want2 = cm.oracle_potential(out["rnew"][order][:n2], sm[order][:n2], N, cenma_m=cenma)
    cm.assert_bits_equal(want2["phi"][:n2 + 2], R.star_phi()[:n2 + 2], "phi after")
    R.free()


@pytest.mark.parametrize("kind,N,masses", [("plummer", 40000, "imf"), ("king", 25000, "equal")])
def test_energy_scheme_velocity_correction(R, kind, N, masses):
    """energy_conservation1 and energy_conservation3 -> set_velocities3 (reference cmc_utils.c:2153-2174, 254-363) of the
    compiled reference against the restatement, bit for bit, on fields that make every branch fire (cm.sv3_fields)"""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 7, masses=masses)
    R.init(N, n_star_dim=N + 64)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N)
    R.L.ref_calc_potential_new()
    sr, smt, phi = R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy()
    t = ol.tables(sr, smt, phi, N, N, 0.0)
    g = np.arange(1, N + 1)
    R.L.ref_energy_conservation1()
    for k, v in ol.energy_conservation1(t, g, vr, vt).items():
        cm.assert_bits_equal(v, R.star_field(k)[1:N + 1], k)
    f = cm.sv3_fields(r, vr, vt, phi[1:N + 1], seed=N)
    for k, v in f.items():
        R.star_field(k, np.int64 if k == "interacted" else np.float64)[1:N + 1] = v
    R.L.ref_set_Eoops(0.125)
    R.L.ref_energy_conservation3()
    io, flags, left = ol.set_velocities3_first(t, g, f)
    assert left > 0 and 0 < flags.sum() < N and (f["interacted"] == 0).sum() - flags.sum() > 10     # sources, sinks and a leftover
    touched = sum(int((io[k] != f[k]).sum()) for k in ("vr", "vt"))
    assert touched > 0.9 * N
    for k in ol.SV3_IO:
        cm.assert_bits_equal(io[k], R.star_field(k)[1:N + 1], k)
    assert R.L.ref_get_Eoops() == 0.125                  # single rank: Eexcess <- Eexcess_prev = 0 before the second loop (:322)
    io2, eoops, passes = ol.set_velocities3(t, [(1, N)], [f])
    assert passes == 1 and eoops[0] == 0.0
    for k in ol.SV3_IO:
        cm.assert_bits_equal(io2[0][k], io[k], k + " (driver)")
    R.free()


@pytest.mark.parametrize("kind,N,masses,p", [("plummer", 30000, "imf", 20), ("king", 20000, "equal", 20), ("plummer", 5000, "imf", 3), ("plummer", 200, "imf", 20)])
def test_calc_sigma_r(R, kind, N, masses, p):
    """calc_sigma_r (reference cmc_dynamics_helper.c:2821-2975) of the compiled reference against the restatement, both
    output modes, bit for bit: the running window sums carry every rounding of the reference's subtract-then-add updates"""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + p, masses=masses)
    R.init(N, n_star_dim=N + 64)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N)
    R.L.ref_calc_potential_new()
    sr, smt = R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy()
    for mave in (False, True):
        a, sg = R.calc_sigma_r(p, N, mave)
        a0, sg0 = ol.calc_sigma_r(p, sr, smt, N, N, vr, vt, mave=mave)
        cm.assert_bits_equal(a[1:], a0[1:], "r" if not mave else "mean mass")
        cm.assert_bits_equal(sg[1:], sg0[1:], "sigma")
    assert np.all(np.isfinite(sg0[1:])) and np.all(sg0[1:] > 0)
    # the drift the reference's running sums accumulate against a freshly summed window: why only an exact replay can match it
    k = N // 2
    w = slice(k - p - 1, k + p - 1)
    fresh = np.sqrt(np.sum(smt[1:][w] / N * (vr[w] ** 2 + vt[w] ** 2)) / np.sum(smt[1:][w] / N))
    assert abs(sg0[k] / fresh - 1) < 1e-9
    R.free()


DYN_CASES = {
    # the Plummer example's configuration (examples/PlummerSphere.ini): relaxation only, odd N (the last star stays unpaired)
    "relax_only": dict(kind="plummer", N=20001, masses="imf", nb=0, dt=2e-3, cfg=dict(relaxation=1)),
    # the King example's switches (examples/KingProfile.ini) with radii and a time step large enough for encounters of every type
    "encounters": dict(kind="king", N=20000, masses="imf", nb=1500, dt=4e-3,
                       cfg=dict(relaxation=1, ss_collision=1, binsingle=1, binbin=1, bh_capture=1, tidal_capture=1, tc_factor=3.0, wd_tc=2.5, co_tde=1)),
    "no_relaxation": dict(kind="plummer", N=6000, masses="equal", nb=0, dt=1e-3, cfg=dict(relaxation=0, ss_collision=1)),
    "threebb_skips": dict(kind="plummer", N=8000, masses="imf", nb=0, dt=2e-3, cfg=dict(relaxation=1), threebb=True),
}


def dyn_case_setup(R, case, seed=11):
    """a cluster with stellar types, radii and binaries on the harness + the same state as plain arrays for the restatement"""
    c = DYN_CASES[case]
    N, nb = c["N"], c["nb"]
    r, vr, vt, sm = cm.cluster_state(c["kind"], N, seed=N + seed, masses=c["masses"])
    rng = np.random.default_rng(seed)
    binind = np.zeros(N, np.int64)
    if nb:
        binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    R.init(N, n_star_dim=N + 64, n_bin_dim=nb + 8, idum=77)
    R.load_cluster(r, vr, vt, sm, binind=binind)
    R.set_scalars(N, N, N_BINARY=nb)
    se_k = rng.choice(np.array([0, 1, 1, 1, 3, 7, 10, 11, 13, 14, 14], np.int32), N).astype(np.int32)
    rad = rng.uniform(2e-4, 3e-3, N)
    R.star_field("se_k", np.int32)[1:N + 1] = se_k
    R.star_field("rad")[1:N + 1] = rad
    bin_a = np.zeros(nb + 1)
    if nb:
        bin_a[1:] = rng.uniform(1e-4, 4e-3, nb)
        layout.binary_field(R.binary_bytes(), "a")[1:nb + 1] = bin_a[1:]
        layout.binary_field(R.binary_bytes(), "inuse")[1:nb + 1] = 1
    threebb = None
    if c.get("threebb"):
        threebb = np.zeros(N, np.int64)
        tri = rng.choice(N // 3 - 2, 40, replace=False) * 3
        for k in range(3):
            threebb[tri + k] = 1
        threebb[N - 30:] = 0
        R.star_field("threebb_interacted", np.int64)[1:N + 1] = threebb
    R.L.ref_calc_potential_new()
    R.compute_energy()
    cfg = dict(ol.DYN_DEFAULTS)
    cfg.update(c["cfg"])
    f = dict(vr=vr.copy(), vt=vt.copy(), E=R.star_field("E")[1:N + 1].copy(), J=R.star_field("J")[1:N + 1].copy(), Y=np.zeros(N), binind=binind,
             se_k=se_k, rad=rad)
    if threebb is not None:
        f["threebb"] = threebb
    tabs = (R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy())
    return c, cfg, f, bin_a, tabs


def per_pair_draws(stream, n_pairs, enc_first_stars, pair_k, relaxation):
    """the reference's sequential stream cut into three draws per pair: an encounter pair (or any pair without relaxation) consumed two"""
    out, pos = np.zeros(3 * n_pairs), 0
    for i in range(n_pairs):
        take = 2 if (pair_k[i] in enc_first_stars or not relaxation) else 3
        out[3 * i:3 * i + take] = stream[pos:pos + take]
        pos += take
    assert pos == len(stream)
    return out


@pytest.mark.parametrize("case", list(DYN_CASES))
def test_dynamics_apply_pairs(R, case):
    """dynamics_apply (reference cmc_dynamics.c:17-542) of the compiled reference, its strong-encounter routines re-bound to
    recorders, against the restated pair loop: kinematics, local density, every cross section, the encounter decision, the relaxation branch
    (new vr, vt, E, J, Y of every star, bit for bit) and the scattering-angle statistics as the reference prints them"""
    c, cfg, f, bin_a, (sr, smt, phi) = dyn_case_setup(R, case)
    N = c["N"]
    R.L.ref_draw_mode(1)
    lk, ld, txt = R.dynamics_apply(c["dt"], cfg)
    stream = R.draws()
    R.L.ref_draw_mode(0)
    # pairing is independent of the draws: a dry run of the restatement gives each pair's first star
    t = ol.tables(sr, smt, phi, N, N)
    dp = ol.dyn_params(**{k: v for k, v in cfg.items() if k in dict(ol.DynParams._fields_)})
    dry = ol.dynamics_apply_pairs(t, 1, c["dt"], dp, np.ones(3 * (N // 2)), {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in f.items()}, bin_a)
    draws = per_pair_draws(stream, len(dry["pair_k"]), set(lk[:, 1].tolist()), dry["pair_k"], cfg["relaxation"])
    got = ol.dynamics_apply_pairs(t, 1, c["dt"], dp, draws, f, bin_a)
    for name in ("vr", "vt", "E", "J", "Y"):
        cm.assert_bits_equal(R.star_field(name)[1:N + 1], f[name], name)
    enc = got["status"] > 0
    enc &= got["status"] < 3
    assert np.array_equal(got["pair_k"][enc], lk[:, 1]) and np.array_equal(got["pair_kp"][enc], lk[:, 2])
    assert np.array_equal(got["status"][enc] - 1, lk[:, 0])
    cm.assert_bits_equal(got["rperi"][enc], ld[:, 0], "rperi")
    cm.assert_bits_equal(got["W"][enc], ld[:, 1], "W")
    cm.assert_bits_equal(got["rcm"][enc], ld[:, 2], "rcm")
    if case == "encounters":
        assert (lk[:, 0] == 0).sum() > 20 and (lk[:, 0] == 1).sum() > 20, lk[:, 0]
    if case == "threebb_skips":
        assert not np.any(np.isin(got["pair_k"], np.nonzero(f["threebb"])[0] + 1))
    if cfg["relaxation"]:
        assert got["Nrel"] == len(got["pair_k"]) - enc.sum() and got["Nrelbeta"][3] > 0
        # the reference's own line in the relaxation file (cmc_dynamics.c:523-531)
        want = "%g" % 0.0
        with np.errstate(invalid="ignore", divide="ignore"):
            for i in range(4):
                nb_ = float(got["Nrelbeta"][i])
                want += " %s %s %s %s" % tuple(_g(x) for x in (nb_ / got["Nrel"], got["qave"][i] / nb_, got["mave"][i] / nb_, got["rave"][i] / nb_))
        assert txt.splitlines()[-1] == want
    R.free()


def _g(x):
    """printf's %g of a double (nan prints as -nan / nan depending on the sign bit, like glibc)"""
    if np.isnan(x):
        return "-nan" if np.signbit(x) else "nan"
    return "%g" % x


@pytest.mark.parametrize("kind,N,masses,p", [("plummer", 30000, "imf", 20), ("king", 20011, "equal", 20), ("plummer", 5000, "imf", 3), ("plummer", 50, "equal", 20)])
def test_simul_relax_new(R, kind, N, masses, p):
    """simul_relax_new, the relaxation time step of GetTimeStep (reference cmc_dynamics_helper.c:2526-2580), compiled reference against the
    restatement: the same double, bit for bit (fresh block sums, sqrt and divisions only)"""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + p, masses=masses)
    R.init(N, n_star_dim=N + 64)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N)
    R.L.ref_calc_potential_new()
    R.L.ref_simul_relax_new.restype = C.c_double
    R.L.ref_simul_relax_new.argtypes = [C.c_long, C.c_double]
    t = ol.tables(R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy(), N, N)
    for theta in (1.0, 1.4142135623730951):
        want = R.L.ref_simul_relax_new(p, theta)
        got = ol.simul_relax_new(t, 1, p, theta, vr, vt)
        assert np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64), (got, want)
    assert (np.isinf(want) and N < 3 * p) or (0 < want < 1)
    R.free()


@pytest.mark.parametrize("treatment,orbit_r", [(1, 6.0), (0, 6.0), (1, 1.0e6)])
def test_tidally_strip_stars(R, treatment, orbit_r):
    """tidally_strip_stars (reference cmc_evolution_thr.c:443-645), the step between energy_conservation2 and ComputeIntermediateEnergy when
    TIDALLY_STRIP_STARS is on (both example .ini files): after an orbit step of the compiled reference, its own loop against the restated
    criterion + the host loop (Rtidal shrinking with the mass lost until a pass strips nothing), both tidal treatments, singles and binaries"""
    N, nb = 20000, 900
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=31, masses="imf")
    rng = np.random.default_rng(5)
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    R.init(N, n_star_dim=N + 64, n_bin_dim=nb + 8, idum=99, orbit_r=orbit_r)
    R.load_cluster(r, vr, vt, sm, binind=binind)
    bb = R.binary_bytes()
    bf = {k: layout.binary_field(bb, k) for k in ("m1", "m2", "a", "Eint1", "Eint2", "inuse")}
    for k, lo, hi in (("m1", 0.1, 5.0), ("m2", 0.1, 5.0), ("a", 1e-6, 1e-3), ("Eint1", 0.0, 1e-7), ("Eint2", 0.0, 1e-7)):
        bf[k][1:nb + 1] = rng.uniform(lo, hi, nb)
    bf["inuse"][1:nb + 1] = 1
    R.star_field("Eint")[1:N + 1] = rng.uniform(0.0, 1e-7, N)
    R.set_scalars(N, N, N_BINARY=nb)
    R.L.ref_calc_potential_new()
    R.compute_energy()
    R.L.ref_new_orbits_calculate()
    sc0 = R.scalars()
    f = {k: R.star_field(k)[1:N + 1].copy() for k in ("r_apo", "E", "J", "Eint", "rnew", "vrnew", "vtnew")}
    bi = R.star_field("binind", np.int64)[1:N + 1].copy()
    bcopy = {k: v[:nb + 1].copy() for k, v in bf.items()}
    t = ol.tables(R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy(), N, N, 0.0)
    smt = R.star_m()[:N + 2].copy()
    Rt, TML, DT = R.tidally_strip_stars(treatment)
    s = dict(Mtotal=sc0["Mtotal"], TidalMassLoss=sc0["TidalMassLoss"], OldTidalMassLoss=0.0, Eescaped=sc0["Eescaped"], Jescaped=sc0["Jescaped"],
             Eintescaped=sc0["Eintescaped"], Ebescaped=sc0["Ebescaped"])
    ol.tidally_strip_loop(lambda thr: ol.tidally_strip_select(treatment, thr, f), lambda x: ol.lib().orc_potential(C.byref(t), x),
                          lambda i: smt[i], lambda i: dict(E=f["E"][i - 1], J=f["J"][i - 1], Eint=f["Eint"][i - 1], binind=int(bi[i - 1])),
                          lambda k: {q: bcopy[q][k] for q in ("m1", "m2", "a", "Eint1", "Eint2")}, s, N, orbit_r, treatment, 0.11)
    sc = R.scalars()
    assert (Rt, TML, DT) == (s["Rtidal"], s["TidalMassLoss"], 0.0) and sc["Rtidal"] == s["Rtidal"]
    for k in ("Eescaped", "Jescaped", "Eintescaped", "Ebescaped", "Etidal"):
        assert sc[k] == s[k], k
    for k in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(R.star_field(k)[1:N + 1], f[k], k)
    if orbit_r < 100:
        assert s["passes"] >= 2 and len(s["stripped"]) > 5
        assert treatment == 0 or (bi[np.array(s["stripped"]) - 1] > 0).sum() > 2       # the Giersz criterion strips hundreds, binaries among them
    else:
        assert s["passes"] == 1 and not s["stripped"]
    R.free()


@pytest.mark.parametrize("cenma", [0.0, 25.0])
def test_compute_energy_totals(R, cenma):
    """ComputeEnergy's energy sums (reference cmc_utils.c:619-655) against the restatement: kinetic, potential (two additions per star),
    internal and binary binding energy, singles, binaries in use and destroyed binaries"""
    N, nb = 20000, 900
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=33, masses="imf")
    rng = np.random.default_rng(9)
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    R.init(N, n_star_dim=N + 64, n_bin_dim=nb + 8)
    R.load_cluster(r, vr, vt, sm, binind=binind)
    bb = R.binary_bytes()
    bf = {k: layout.binary_field(bb, k) for k in ("m1", "m2", "a", "Eint1", "Eint2", "inuse")}
    for k, lo, hi in (("m1", 0.1, 5.0), ("m2", 0.1, 5.0), ("a", 1e-6, 1e-3), ("Eint1", 0.0, 1e-7), ("Eint2", 0.0, 1e-7)):
        bf[k][1:nb + 1] = rng.uniform(lo, hi, nb)
    bf["inuse"][1:nb + 1] = (rng.uniform(size=nb) < 0.9).astype(np.int32)
    Eint = rng.uniform(0.0, 1e-7, N)
    R.star_field("Eint")[1:N + 1] = Eint
    R.set_scalars(N, N, N_BINARY=nb, cenma_m=cenma)
    R.L.ref_calc_potential_new()
    R.L.ref_compute_energy()
    et = np.zeros(5)
    R.L.ref_get_etotal.argtypes = [C.c_void_p]
    R.L.ref_get_etotal(et.ctypes.data)
    t = ol.tables(R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy(), N, N, cenma)
    K, P, Ei, Eb = ol.compute_energy_totals(t, 1, vr, vt, Eint, binind, {k: v[:nb + 1].copy() for k, v in bf.items()})
    sc = R.scalars()
    assert (K, 0.5 * P, Ei, Eb) == (et[1], et[2], et[3], et[4])
    assert et[0] == (K + 0.5 * P + Ei + Eb) + (0.0 + sc["Eescaped"] + sc["Ebescaped"] + sc["Eintescaped"])      # cenma.E = 0 here
    assert K > 0 and P < 0 and Ei > 0 and Eb < 0
    R.free()


def test_comp_multi_mass_percent(R):
    """comp_multi_mass_percent (reference cmc_utils.c:769-861): Lagrange radii per mass bin, restatement against the compiled
    reference.  The interpolation is GSL's gsl_interp_linear, un-vendored: both sides run a restatement of its published algorithm
    (oracle/ref_shim/gsl_shim.c), so the pin covers the binning, the sequential per-bin sums and the search, not GSL itself."""
    N = 30000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=37, masses="imf")
    R.init(N, n_star_dim=N + 64)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N)
    R.L.ref_calc_potential_new()
    sr, smt = R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy()
    q = np.quantile(smt[1:N + 1], [0.0, 0.3, 0.6, 0.9, 0.99, 0.9999, 1.0])
    bins = np.concatenate([q[:-1] * [0.999, 1, 1, 1, 1, 1], [q[-2] * 1.0000001, q[-1] * 1.5, q[-1] * 2.0]])   # a one-star bin and an empty one at the top
    pcs = np.array([0.0, 0.001, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.99, 1.0])
    got = R.comp_multi_mass_percent(bins, 1.0, pcs)
    want = ol.comp_multi_mass_percent(sr, smt, N, bins, 1.0, pcs)
    cm.assert_bits_equal(got, want, "multi_mass_r")
    assert (want[-1] == -1.0).all() and (want[:4, 1:] > 0).all() and (want[:4, 0] == 0).all() and np.all(np.diff(want[0][1:]) > 0)
    counts = [(want[b] != -1).all() for b in range(len(bins))]
    assert sum(counts) >= 6
    R.free()


def test_sort_with_binaries_and_stripped_stars(R):
    n, nb, n_strip = 20000, 1500, 23
    rng = np.random.default_rng(8)
    R.init(n, n_star_dim=n + 64, n_bin_dim=nb + 8)
    sb, bb = R.star_bytes(), R.binary_bytes()
    sb[1:n + 1] = rng.integers(0, 256, (n, R.sizeof_star), dtype=np.uint8)
    bb[1:nb + 1] = rng.integers(0, 256, (nb, R.sizeof_binary), dtype=np.uint8)
    r = rng.uniform(1e-3, 1e2, n)
    r[rng.choice(n, n_strip, replace=False)] = cm.SF_INFINITY
    R.star_field("r")[1:n + 1] = r
    R.star_field("m")[1:n + 1] = rng.uniform(0.1, 2.0, n)
    binind = np.zeros(n, np.int64)
    live = np.nonzero(r < cm.SF_INFINITY)[0]
    who = rng.choice(live, nb, replace=False)
    binind[who] = rng.permutation(np.arange(1, nb + 1))
    R.star_field("binind", np.int64)[1:n + 1] = binind
    R.set_scalars(n, n, N_BINARY=nb)
    mine = sb[1:n + 1].copy()
    mine_b = bb[1:].copy()
    n_new, nb_out, perm = ol.sort_stars(mine, n, R.sizeof_star, 0, R.star_offset("binind"), mine_b, nb + 8, R.sizeof_binary)
    R.L.ref_qsorts_new()
    sc = R.scalars()
    assert int(sc["N_MAX"]) == n_new == n - n_strip and int(sc["N_b_local"]) == nb_out == nb
    assert np.array_equal(sb[1:n_new + 1], mine[:n_new])
    assert np.array_equal(bb[1:], mine_b)
    R.free()


def test_potential_lookup_and_orbit_probes(R):
    """potential(r), find_zero_Q, function_q, calc_orbit_new as standalone functions on random arguments"""
    N = 5000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=3, masses="imf")
    R.init(N, n_star_dim=N + 64)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N)
    R.L.ref_calc_potential_new()
    R.compute_energy()
    sr, smt, phi = R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy()
    t = ol.tables(sr, smt, phi, N, N)
    rng = np.random.default_rng(0)
    L = ol.lib()
    for rad in np.concatenate([rng.uniform(0.0, r[-1], 300), r[:50], [0.0, r[0] * 0.5, r[-1]]]):
        a, b = R.L.ref_potential(float(rad)), L.orc_potential(C.byref(t), float(rad))
        assert a == b or (np.isnan(a) and np.isnan(b))
    E, J = R.star_field("E")[1:N + 1].copy(), R.star_field("J")[1:N + 1].copy()
    o = ol.Orbit()
    buf = np.zeros(7)
    for j in rng.integers(1, N + 1, 400):
        Ej = E[j - 1] + (smt[j] / sr[j]) * (1.0 / N)
        R.L.ref_calc_orbit_new(int(j), Ej, J[j - 1], buf.ctypes.data)
        L.orc_calc_orbit_new(C.byref(t), int(j), 0, 0.0, Ej, J[j - 1], C.byref(o))
        assert (buf[0], buf[1], buf[2], buf[3], buf[4], buf[5], buf[6]) == (o.rp, o.ra, o.dQdrp, o.dQdra, o.kmax, o.kmin, o.circular_flag)
    R.free()


def test_find_root_vr_brent_shim_agrees(R):
    """find_root_vr (cmc_nr.c:500-603) iterates GSL's Brent solver, an un-vendored dependency: the compiled reference runs on
    oracle/ref_shim/gsl_shim.c, the oracle on its own restatement -- two independent restatements of the published algorithm must
    agree bit for bit (this branch stays 'parity unpinned' against a real GSL)."""
    N = 4000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=5)
    R.init(N, n_star_dim=N + 64)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N)
    R.L.ref_calc_potential_new()
    R.compute_energy()
    sr, smt, phi = R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy()
    t = ol.tables(sr, smt, phi, N, N)
    E, J = R.star_field("E")[1:N + 1].copy(), R.star_field("J")[1:N + 1].copy()
    buf = np.zeros(7)
    L = ol.lib()
    rng = np.random.default_rng(1)
    n_checked = 0
    for j in rng.integers(1, N + 1, 300):
        Ej = E[j - 1] + (smt[j] / sr[j]) * (1.0 / N)
        R.L.ref_calc_orbit_new(int(j), Ej, J[j - 1], buf.ctypes.data)
        if buf[6]:
            continue
        for k in (int(buf[5]), int(buf[4])):
            if k < 1:
                continue
            a, b = R.L.ref_find_root_vr(int(j), k, Ej, J[j - 1]), L.orc_find_root_vr(C.byref(t), int(j), k, Ej, J[j - 1])
            assert a == b
            n_checked += 1
    assert n_checked > 100
    R.free()


def test_orbit_step_where_find_root_vr_fires(R):
    """The whole orbit step of the compiled reference on common.brent_cluster -- hundreds of stars take calc_orbit_new's Brent fallback
    (cmc_orbit.c:459-474, 485-500) -- against the oracle, bit for bit.  (GSL itself stays un-vendored: both sides iterate a restatement of
    gsl_root_fsolver_brent, so this pins everything around the solver: when the fallback is taken, the bracket, the function it
    evaluates -- calc_vr_within_interval / calc_pot_within_interval / function_q --, the post-selection of the end with v_r^2 >= 0.)"""
    N = 20000
    r, vr, vt, sm, pot, E, J, planted = cm.brent_cluster(N)
    R.init(N, n_star_dim=N + 64, idum=77)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N)
    R.L.ref_calc_potential_new()
    cm.assert_bits_equal(R.star_phi()[:N + 2], pot["phi"][:N + 2], "phi")
    R.star_field("E")[1:N + 1], R.star_field("J")[1:N + 1] = E, J
    before = {k: R.star_field(k)[1:N + 1].copy() for k in ("X", "r_peri", "r_apo")}
    R.L.ref_draw_mode(1)
    R.L.ref_new_orbits_calculate()
    R.L.ref_draw_mode(0)
    draws = R.draws()
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N)
    want = ol.get_positions(t, np.arange(1, N + 1), E, J, vr, vt, R.scalars()["Rtidal"], draws)
    assert want["draw_off"][-1] == len(draws) and want["n_brent"] > 150
    ok = want["status"] == ol.ST_OK
    for k in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(R.star_field(k)[1:N + 1], want[k], k)
    for k in ("X", "r_peri", "r_apo"):
        cm.assert_bits_equal(R.star_field(k)[1:N + 1][ok], want[k][ok], k)
        cm.assert_bits_equal(R.star_field(k)[1:N + 1][~ok], before[k][~ok], k + " of the other stars is untouched")
    R.free()


def test_taus113_known_answers(R):
    """seeding, stepping and the 2^80 jump of the reference's generator (src/libs/taus113-v2.c) vs the oracle's restatement; includes the
    reference's own self-check: 2^20 steps == one JPoly_2_20 jump (testjumping(), taus113-v2.c:272-292)"""
    for seed in (1, 1234, 1729217, 0):
        z = np.zeros(4, np.uint32)
        R.L.ref_rng_reset(seed, z.ctypes.data)
        mine = ol.Taus113(seed)
        assert np.array_equal(z, mine.z)
        a = np.array([R.L.ref_rng_next(z.ctypes.data) for _ in range(100)])
        assert np.array_equal(a, mine.draws(100))
        R.L.ref_rng_jump80(z.ctypes.data)
        mine.jump(ol.Taus113.JPOLY_2_80)
        assert np.array_equal(z, mine.z)
    g1, g2 = ol.Taus113(99), ol.Taus113(99)
    g1.draws(1 << 20)
    g2.jump(ol.Taus113.JPOLY_2_20)
    assert np.array_equal(g1.z, g2.z)


def _snapshot_state(R, N, nb, seed):
    """a sorted cluster in the compiled reference with random content in every field write_snapshot reads: stellar types (some black holes),
    luminosities, spins, binaries with BSE blocks"""
    rng = np.random.default_rng(seed)
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=seed, masses="imf")
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    R.init(N, n_star_dim=N + 64, n_bin_dim=nb + 8, idum=7)
    sb, bb = R.star_bytes(), R.binary_bytes()
    # finite random doubles everywhere (the se_* / bse_* blocks are copied, not computed with), then the fields with a meaning
    sb[1:N + 1] = rng.uniform(-50.0, 50.0, (N, R.sizeof_star // 8)).view(np.uint8).reshape(N, R.sizeof_star)
    bb[1:nb + 1] = rng.uniform(-50.0, 50.0, (nb, R.sizeof_binary // 8)).view(np.uint8).reshape(nb, R.sizeof_binary)
    bb[0] = 0
    R.load_cluster(r, vr, vt, sm, binind=binind, ids=rng.permutation(np.arange(10, N + 10)))
    R.star_field("E")[1:N + 1], R.star_field("J")[1:N + 1] = rng.uniform(-2, 0, N), rng.uniform(0, 1, N)
    R.star_field("rad")[1:N + 1] = rng.uniform(1e-9, 1e-7, N)
    layout.star_field(sb, "se_k")[1:N + 1] = rng.integers(0, 16, N)
    for k, lo, hi in (("m1", 0.1, 5.0), ("m2", 0.1, 5.0), ("a", 1e-6, 1e-3), ("e", 0.0, 0.99), ("rad1", 1e-9, 1e-7), ("rad2", 1e-9, 1e-7)):
        layout.binary_field(bb, k)[1:nb + 1] = rng.uniform(lo, hi, nb)
    layout.binary_field(bb, "id1", np.int64)[1:nb + 1] = rng.integers(1, 10 ** 7, nb)
    layout.binary_field(bb, "id2", np.int64)[1:nb + 1] = rng.integers(1, 10 ** 7, nb)
    kw = np.lib.stride_tricks.as_strided(bb.reshape(-1)[84:92].view(np.int32), shape=(bb.shape[0], 2), strides=(R.sizeof_binary, 4))   # bse_kw[2]
    kw[1:nb + 1] = rng.integers(0, 16, (nb, 2))
    R.set_scalars(N, N, N_BINARY=nb)
    R.L.ref_calc_potential_new()
    sigma = np.concatenate(([0.0], rng.uniform(0.2, 0.9, N)))
    return sigma


@pytest.mark.parametrize("bh_only", [False, True])
def test_write_snapshot_rows(R, bh_only):
    """SURVEY 8f N4, the HDF5 snapshot table: write_snapshot (cmc_io.c:2960-3345) compiled unmodified with its HDF5 calls re-bound to
    recorders, against the host build of csrc/snapshot_core.h (what the device kernel runs) on the same arrays -- the table's geometry
    (62 fields, offsets, H5 types, 488-byte rows) and every field of every row bit for bit: unit conversions, -100 markers, the NaN block of
    single stars, binaries' BSE blocks, bin_Eb and eta (calc_average_mass_sqr windows at both ends of the cluster); all stars and bh_only"""
    import hostemul as he
    N, nb = 6000, 700
    sigma = _snapshot_state(R, N, nb, seed=123)
    units_m, units_l = 1.989e33 * 4.0e5, 3.0857e18 * 1.7
    rows, geom = R.write_snapshot(units_m, units_l, sigma=sigma, avekernel=20, bh_only=bh_only)
    dt = layout.SNAPSHOT_DTYPE
    assert [g[0] for g in geom] == list(layout.SNAPSHOT_FIELDS)
    assert [g[1] for g in geom] == [dt.fields[n][1] for n in layout.SNAPSHOT_FIELDS]
    assert [g[2] for g in geom] == [{np.dtype(np.float64): 1, np.dtype(np.int64): 2, np.dtype(np.int32): 3}[dt.fields[n][0]] for n in layout.SNAPSHOT_FIELDS]
    sb, bb = R.star_bytes().copy(), R.binary_bytes().copy()
    bi = layout.star_field(sb, "binind")[1:N + 1]
    assert len(rows) == (N if not bh_only else len(rows)) and 0 < len(rows) <= N
    single = np.flatnonzero(rows["binflag"] == -100)
    nan_bits = int(rows["rad0"][single[0]].view(np.uint64))
    assert np.isnan(rows["rad0"][single[0]])
    got = he.snapshot_rows(sb, bb, R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy(), sigma, N, N, 1, N, N,
                           units_m, units_l, 20, nan_bits, bh_only=bh_only)
    assert len(got) == len(rows)
    if not bh_only:
        assert (rows["binflag"] == 1).sum() == nb and np.array_equal(rows["binflag"] == 1, bi > 0)
    else:
        assert len(rows) < N // 2
    for n in layout.SNAPSHOT_FIELDS:
        a, b = rows[n], got[n]
        if a.dtype == np.float64:
            cm.assert_bits_equal(a, b, n)
        else:
            assert np.array_equal(a, b), n
    assert np.isfinite(rows["eta"][rows["binflag"] == 1]).all() and (rows["bin_Eb"][rows["binflag"] == 1] < 0).all()


def test_write_snapshot_rows_with_stars_created_this_step(R):
    """the same with M stars behind the old shard (clus.N_MAX_NEW > End - Start + 1: collision products before the next sort): their table
    entries sit at N_MAX + 2 .. (get_global_idx, cmc_utils.c:2384-2407), one of them is a binary whose mass window reaches back into the
    old stars"""
    import hostemul as he
    N, nb, M = 3000, 200, 37
    sigma = _snapshot_state(R, N, nb, seed=321)
    rng = np.random.default_rng(9)
    sb, bb = R.star_bytes(), R.binary_bytes()
    sb[N + 1:N + M + 1] = rng.uniform(-50.0, 50.0, (M, R.sizeof_star // 8)).view(np.uint8).reshape(M, R.sizeof_star)
    bi = layout.star_field(sb, "binind")
    bi[N + 1:N + M + 1] = 0
    donor = int(np.flatnonzero(bi[1:N + 1] > 0)[0]) + 1
    bi[N + 5], bi[donor] = bi[donor], 0      # a binary among the new stars (it takes over the slot of an old star, which becomes single)
    layout.star_field(sb, "se_k")[N + 1:N + M + 1] = rng.integers(0, 16, M)
    layout.star_field(sb, "id")[N + 1:N + M + 1] = 10 ** 6 + np.arange(M)
    R.star_field("rad")[N + 1:N + M + 1] = rng.uniform(1e-9, 1e-7, M)
    R.star_r()[N + 2:N + 2 + M], R.star_m()[N + 2:N + 2 + M], R.star_phi()[N + 2:N + 2 + M] = rng.uniform(0.1, 3.0, M), rng.uniform(0.1, 5.0, M), -rng.uniform(0.1, 1.0, M)
    R.set_scalars(N, N + M, N_BINARY=nb)
    sigma = np.concatenate((sigma, rng.uniform(0.2, 0.9, M)))
    rows, _ = R.write_snapshot(7.0e38, 2.0e18, sigma=sigma, avekernel=20)
    nt = N + 2 + M + 64
    got = he.snapshot_rows(sb.copy(), bb.copy(), R.star_r()[:nt].copy(), R.star_m()[:nt].copy(), R.star_phi()[:nt].copy(), sigma, N + M, N, 1, N, N,
                           7.0e38, 2.0e18, 20, 0xFFF8000000000000)
    assert len(rows) == len(got) == N + M and rows["binflag"][N + 4] == 1 and rows["id"][N + M - 1] == 10 ** 6 + M - 1
    cm.assert_bits_equal(rows["r"][N:], R.star_r()[N + 2:N + 2 + M].copy(), "r of the new stars")
    for n in layout.SNAPSHOT_FIELDS:
        if rows[n].dtype == np.float64:
            cm.assert_bits_equal(rows[n], got[n], n)
        else:
            assert np.array_equal(rows[n], got[n]), n
