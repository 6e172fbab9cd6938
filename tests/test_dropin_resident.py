"""The RESIDENT mode of the drop-in shim, end to end: the reference's own compiled objects and harness with the whole block
energy_conservation1 .. ComputeEnergy of the main loop (reference src/cmc/cmc.c:302-353) re-bound at link time (ld --wrap) to
cmc_cosmic_b200/shim/cmcb200_shim.c built with -DCMCB200_SHIM_RESIDENT -- the records go up once when energy_conservation1 opens the window
and come down once when ComputeEnergy closes it -- against the untouched reference build, on identical inputs and identical taus113 draws,
for two consecutive timesteps.  Everything the reference leaves behind must be byte-identical: star[], binary[], the replicated tables, the
escaper accumulators and the esc log length, Eoops, Etotal, the Lagrange radii."""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
import oraclelib as ol
import refharness as rh

REF_B200RES_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "libcmc_ref_b200res.so")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (rh.available() and os.path.exists(REF_B200RES_SO)), reason="oracle/_ref not built")]
PCS = np.array([0.01, 0.1, 0.5, 0.9])


def _bits(x):
    """scalars compared by their bit patterns (a NaN equals itself)"""
    if isinstance(x, dict):
        return {k: _bits(v) for k, v in x.items()}
    if isinstance(x, (tuple, list)):
        return tuple(_bits(v) for v in x)
    return np.float64(x).view(np.uint64)


def _timestep(R, snap, step, treatment, draws=None, draw_off=None, is_shim=False):
    """cmc.c:302-353 in the reference's order, through the harness's entry points"""
    N = int(R.scalars()["N_MAX"])
    R.L.ref_energy_conservation1()
    if is_shim:
        d = np.ascontiguousarray(draws); o = np.ascontiguousarray(draw_off, np.int64)
        R.L.cmcb200_shim_set_draws.argtypes = [C.c_void_p, C.c_void_p]
        R.L.cmcb200_shim_set_draws(d.ctypes.data, o.ctypes.data)
        snap[f"_keep{step}"] = (d, o)
    else:
        snap[f"E{step}"] = R.star_field("E")[1:N + 1].copy()
        snap[f"J{step}"] = R.star_field("J")[1:N + 1].copy()
        snap[f"vr{step}"] = R.star_field("vr")[1:N + 1].copy()
        snap[f"vt{step}"] = R.star_field("vt")[1:N + 1].copy()
        snap[f"tables{step}"] = (R.star_r()[:N + 2].copy(), R.star_m()[:N + 2].copy(), R.star_phi()[:N + 2].copy())
        snap[f"Rtidal{step}"] = R.scalars()["Rtidal"]
        R.L.ref_draw_mode(1)
    R.L.ref_new_orbits_calculate()
    if not is_shim:
        R.L.ref_draw_mode(0)
        snap[f"draws{step}"] = R.draws()
    sc = R.scalars()
    snap[f"after_orbits{step}"] = {k: sc[k] for k in ("max_r", "TidalMassLoss", "Eescaped", "Jescaped", "Eintescaped", "Ebescaped", "Etidal", "esc_bytes")}
    R.L.ref_energy_conservation2()
    snap[f"strip{step}"] = R.tidally_strip_stars(treatment, 0.11)
    sc = R.scalars()
    snap[f"after_strip{step}"] = {k: sc[k] for k in ("Rtidal", "TidalMassLoss", "Eescaped", "Jescaped", "Eintescaped", "Ebescaped", "Etidal", "esc_bytes")}
    R.L.ref_compute_intermediate_energy()
    R.L.ref_qsorts_new()
    R.L.ref_post_sort_comm()
    R.L.ref_calc_potential_new()
    R.L.ref_energy_conservation3()
    snap[f"lagrange{step}"] = R.comp_mass_percent(PCS)
    R.L.ref_compute_energy()
    et = np.zeros(5)
    R.L.ref_get_etotal.argtypes = [C.c_void_p]
    R.L.ref_get_etotal(et.ctypes.data)
    n2 = int(R.scalars()["N_MAX"])
    snap[f"n{step}"] = n2
    snap[f"Etotal{step}"] = et
    snap[f"Eoops{step}"] = R.L.ref_get_Eoops()
    snap[f"star{step}"] = R.star_bytes()[:n2 + 1].copy()
    snap[f"binary{step}"] = R.binary_bytes().copy()
    snap[f"tables_final{step}"] = (R.star_r()[:n2 + 2].copy(), R.star_m()[:n2 + 2].copy(), R.star_phi()[:n2 + 2].copy())
    snap[f"scalars{step}"] = R.scalars()


def _run(R, N, r, vr, vt, sm, binind, nb, orbit_r, treatment, ref_snap=None, dirty2=None):
    is_shim = ref_snap is not None
    R.init(N, n_star_dim=N + 64, n_bin_dim=nb + 8, idum=99, orbit_r=orbit_r)
    R.load_cluster(r, vr, vt, sm, binind=binind)
    rng = np.random.default_rng(N + 1)
    if nb:
        words = (np.arange(1, nb + 1, dtype=np.int64)[:, None] * 64 + np.arange(R.sizeof_binary // 8, dtype=np.int64)[None, :])
        R.binary_bytes()[1:nb + 1] = words.view(np.uint8).reshape(nb, R.sizeof_binary)
        bb = R.binary_bytes()
        for name, lo, hi in (("m1", 0.1, 5.0), ("m2", 0.1, 5.0), ("a", 1e-6, 1e-3), ("Eint1", 0.0, 1e-7), ("Eint2", 0.0, 1e-7), ("e", 0.0, 0.9)):
            off = R.binary_offset(name)
            bb[1:nb + 1, off:off + 8] = rng.uniform(lo, hi, nb).view(np.uint8).reshape(nb, 8)
        off = R.binary_offset("inuse")
        bb[1:nb + 1, off:off + 4] = np.ones(nb, np.int32).view(np.uint8).reshape(nb, 4)
    R.star_field("Eint")[1:N + 1] = rng.uniform(0.0, 1e-7, N)
    R.star_field("interacted", np.int64)[1:N + 1] = (rng.uniform(size=N) < 0.02)
    # the cold part of the records travels too
    R.star_bytes()[1:N + 1, 216:] = rng.integers(0, 256, (N, R.sizeof_star - 216), dtype=np.uint8)
    R.star_field("se_k", np.int32)[1:N + 1] = 1
    R.set_scalars(N, N, N_BINARY=nb)
    R.L.ref_calc_potential_new()
    R.compute_energy()
    snap = {}
    for step in (1, 2):
        if is_shim and step == 2 and dirty2 is not None:
            R.L.cmcb200_shim_set_dirty_words.argtypes = [C.c_ulonglong]
            R.L.cmcb200_shim_set_dirty_words(dirty2)
        if is_shim:
            srt, smt, phi = ref_snap[f"tables{step}"]
            n = len(ref_snap[f"E{step}"])
            t = ol.tables(srt.copy(), smt.copy(), phi.copy(), n, N)
            off = ol.get_positions(t, np.arange(1, n + 1), ref_snap[f"E{step}"], ref_snap[f"J{step}"], ref_snap[f"vr{step}"], ref_snap[f"vt{step}"],
                                   ref_snap[f"Rtidal{step}"], ref_snap[f"draws{step}"])["draw_off"]
            assert off[-1] == len(ref_snap[f"draws{step}"])
            _timestep(R, snap, step, treatment, ref_snap[f"draws{step}"], off, True)
        else:
            _timestep(R, snap, step, treatment)
    if is_shim:
        R.L.cmcb200_shim_shutdown()
    R.free()
    return snap


@pytest.mark.parametrize("kind,N,masses,nb,orbit_r,treatment,dirty2", [
    ("plummer", 30000, "equal", 0, 1.0e6, 1, None),
    ("plummer", 24000, "imf", 2000, 5.0, 0, None),
    ("king", 26000, "imf", 1500, 4.0, 1, 0),          # second window: the host wrote nothing in between and says so
])
def test_reference_with_resident_shim_equals_reference(kind, N, masses, nb, orbit_r, treatment, dirty2):
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 3, masses=masses)
    rng = np.random.default_rng(N)
    binind = np.zeros(N, np.int64)
    if nb:
        binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    a = _run(rh.Ref(), N, r, vr, vt, sm, binind, nb, orbit_r, treatment)
    b = _run(rh.Ref(REF_B200RES_SO), N, r, vr, vt, sm, binind, nb, orbit_r, treatment, ref_snap=a, dirty2=dirty2)
    for step in (1, 2):
        assert _bits(b[f"after_orbits{step}"]) == _bits(a[f"after_orbits{step}"]), (step, a[f"after_orbits{step}"], b[f"after_orbits{step}"])
        assert _bits(b[f"strip{step}"]) == _bits(a[f"strip{step}"]) and _bits(b[f"after_strip{step}"]) == _bits(a[f"after_strip{step}"]), (step, a[f"after_strip{step}"], b[f"after_strip{step}"])
        assert b[f"n{step}"] == a[f"n{step}"]
        assert np.array_equal(b[f"star{step}"], a[f"star{step}"]), f"step {step}: star[] after the timestep"
        assert np.array_equal(b[f"binary{step}"], a[f"binary{step}"]), f"step {step}: binary[]"
        for x, y, name in zip(b[f"tables_final{step}"], a[f"tables_final{step}"], ("star_r", "star_m", "star_phi")):
            cm.assert_bits_equal(x, y, f"step {step}: {name}")
        assert _bits(b[f"scalars{step}"]) == _bits(a[f"scalars{step}"]), (step, a[f"scalars{step}"], b[f"scalars{step}"])
        assert np.array_equal(b[f"Etotal{step}"], a[f"Etotal{step}"]), (step, a[f"Etotal{step}"], b[f"Etotal{step}"])
        assert b[f"Eoops{step}"] == a[f"Eoops{step}"]
        for k in a[f"lagrange{step}"]:
            cm.assert_bits_equal(b[f"lagrange{step}"][k], a[f"lagrange{step}"][k], f"step {step}: {k}")
    if orbit_r < 100:
        assert a["n1"] < N and (a["after_orbits1"]["esc_bytes"] > 0 or a["after_strip1"]["esc_bytes"] > 0), "the case must remove or strip stars"


def _dyn_setup(R, N, r, vr, vt, sm, seed):
    R.init(N, n_star_dim=N + 64, n_bin_dim=8, idum=77)
    R.load_cluster(r, vr, vt, sm)
    R.set_scalars(N, N, N_BINARY=0)
    rng = np.random.default_rng(seed)
    R.star_field("se_k", np.int32)[1:N + 1] = rng.choice(np.array([0, 1, 1, 1, 3, 7, 10, 11, 13, 14, 14], np.int32), N)
    R.star_field("rad")[1:N + 1] = rng.uniform(1e-4, 1e-3, N)
    R.star_bytes()[1:N + 1, 232:] = rng.integers(0, 256, (N, R.sizeof_star - 232), dtype=np.uint8)     # the cold part of the records travels too
    R.L.ref_calc_potential_new()
    R.compute_energy()


@pytest.mark.parametrize("kind,N,masses,cfg", [
    ("plummer", 30001, "imf", dict(relaxation=1)),                                               # examples/PlummerSphere.ini
    ("king", 24000, "imf", dict(relaxation=1, ss_collision=1, bh_capture=1, tidal_capture=1, tc_factor=3.0, wd_tc=2.5)),
])
def test_dynamics_apply_through_the_resident_shim(kind, N, masses, cfg):
    """dynamics_apply of the compiled reference re-bound to the shim (it opens the window: calc_sigma_r and the pair loop run on the device,
    the encounter pairs come back for the reference's own sscollision_do -- here the harness's recorder) against the untouched reference on
    the same state and the same draws: sigma_array bit for bit, the encounter list identical, vr / vt / E / J of the relaxed stars to 1e-12,
    every other byte of star[] identical, the relaxation file's line equal."""
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 5, masses=masses)
    full = dict(ol.DYN_DEFAULTS)
    full.update(cfg)
    dt = 3e-3
    A = rh.Ref()
    _dyn_setup(A, N, r, vr, vt, sm, 5)
    A.L.ref_draw_mode(1)
    lk, ld, txt = A.dynamics_apply(dt, full)
    stream = A.draws()
    A.L.ref_draw_mode(0)
    A.L.ref_sigma_array.restype = C.c_void_p
    A.L.ref_sigma_array.argtypes = [C.c_int]
    sig_a = [A._view(A.L.ref_sigma_array(w), (N + 1) * 8, np.float64).copy() for w in (0, 1)]
    star_a = A.star_bytes()[:N + 1].copy()
    fa = {k: A.star_field(k)[1:N + 1].copy() for k in ("vr", "vt", "E", "J", "Y")}
    A.free()
    # three draws per pair from the reference's sequential stream (an encounter pair consumed two)
    enc_first = set(lk[:, 1].tolist())
    draws, pos = np.zeros(3 * (N // 2)), 0
    for i in range(N // 2):
        take = 2 if (2 * i + 1) in enc_first else 3
        draws[3 * i:3 * i + take] = stream[pos:pos + take]
        pos += take
    assert pos == len(stream)
    B = rh.Ref(REF_B200RES_SO)
    _dyn_setup(B, N, r, vr, vt, sm, 5)
    B.L.cmcb200_shim_set_dynamics_draws.argtypes = [C.c_void_p]
    B.L.cmcb200_shim_set_dynamics_draws(draws.ctypes.data)
    lk2, ld2, txt2 = B.dynamics_apply(dt, full)
    B.L.cmcb200_shim_window_open.restype = C.c_int
    assert B.L.cmcb200_shim_window_open() == 1, "the binding did not take the device path"
    B.L.ref_sigma_array.restype = C.c_void_p
    B.L.ref_sigma_array.argtypes = [C.c_int]
    for w, name in ((0, "sigma_array.r"), (1, "sigma_array.sigma")):
        cm.assert_bits_equal(B._view(B.L.ref_sigma_array(w), (N + 1) * 8, np.float64)[1:], sig_a[w][1:], name)
    B.L.cmcb200_shim_close_window()
    assert np.array_equal(lk2, lk)
    np.testing.assert_allclose(ld2, ld, rtol=1e-12)
    cm.assert_bits_equal(B.star_field("Y")[1:N + 1], fa["Y"], "Y")
    vs = np.sqrt(fa["vr"] ** 2 + fa["vt"] ** 2)
    for k, scale in (("vr", vs), ("vt", vs), ("E", np.abs(fa["E"]) + vs ** 2), ("J", np.abs(fa["J"]) + 1e-300)):
        assert (np.abs(B.star_field(k)[1:N + 1] - fa[k]) / scale).max() <= 1e-12, k
    mask = np.ones(B.sizeof_star, bool)
    for k in ("vr", "vt", "E", "J"):
        mask[B.star_offset(k):B.star_offset(k) + 8] = False
    assert np.array_equal(B.star_bytes()[:N + 1][:, mask], star_a[:, mask])
    assert txt2.splitlines()[-1].split()[:2] == txt.splitlines()[-1].split()[:2]          # time and the first fraction; the sums differ in the last digit at most
    np.testing.assert_allclose([float(x) for x in txt2.splitlines()[-1].split()], [float(x) for x in txt.splitlines()[-1].split()], rtol=1e-5, equal_nan=True)
    if cfg.get("ss_collision"):
        assert len(lk) > 20
    B.L.cmcb200_shim_shutdown()
    B.free()


def test_whole_timestep_from_dynamics_apply_in_one_window():
    """dynamics_apply opens the window and ComputeEnergy closes it: the whole timestep of src/cmc/cmc.c:286-353 on the resident records,
    through the reference's own function names, with the device's own Philox draws.  Checked through what does not depend on the draws:
    the stars are a permutation of the input, sorted, the tables are the records', the potential obeys its recurrence, energy is conserved
    to the scheme's accuracy."""
    N = 40000
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=8, masses="equal")
    B = rh.Ref(REF_B200RES_SO)
    _dyn_setup(B, N, r, vr, vt, sm, 6)
    et = np.zeros(5)
    B.L.ref_get_etotal.argtypes = [C.c_void_p]
    B.L.ref_compute_energy()
    B.L.ref_get_etotal(et.ctypes.data)
    e0 = et[0]
    full = dict(ol.DYN_DEFAULTS)
    for step in range(2):
        B.dynamics_apply(2e-3, full)
        B.L.cmcb200_shim_window_open.restype = C.c_int
        assert B.L.cmcb200_shim_window_open() == 1
        B.L.ref_energy_conservation1()
        B.L.ref_new_orbits_calculate()
        B.L.ref_energy_conservation2()
        B.tidally_strip_stars(1, 0.11)
        B.L.ref_compute_intermediate_energy()
        B.L.ref_qsorts_new()
        B.L.ref_post_sort_comm()
        B.L.ref_calc_potential_new()
        B.L.ref_energy_conservation3()
        B.L.ref_compute_energy()
        assert B.L.cmcb200_shim_window_open() == 0
    n = int(B.scalars()["N_MAX"])
    assert n == N
    ids = B.star_field("id", np.int64)[1:n + 1]
    assert np.array_equal(np.sort(ids), np.arange(1, N + 1))
    rr = B.star_field("r")[1:n + 1]
    assert np.all(np.diff(rr) >= 0) and np.array_equal(rr, B.star_r()[1:n + 1])
    B.L.ref_get_etotal(et.ctypes.data)
    assert abs(et[0] / e0 - 1) < 2e-3, (et[0], e0)
    B.L.cmcb200_shim_shutdown()
    B.free()
