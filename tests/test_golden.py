"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from runs of the UNMODIFIED reference) against

  * the oracle restatement (oracle/cmc_oracle.c)               -- CPU, pins the oracle where /root/reference does not exist
  * the host build of the kernel's per-star code (orbit_core.h) -- CPU
  * the CUDA path through the C ABI                             -- `-m gpu`

Everything is compared bit for bit (the north_star bars are: indices bit-exact, Phi <= 1e-12, r_min/r_max/r <= 1e-10).
"""
import glob
import os

import numpy as np
import pytest

import common as cm
import oraclelib as ol

ALL = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(cm.GOLDEN, "*.npz")))
CASES = [c for c in ALL if not c.startswith(("lagrange_", "energy_", "snapshot_", "plummer_lagrad_run"))]   # the last: the Lagrange radii of the reference's example run
SNAPSHOT = [c for c in ALL if c.startswith("snapshot_")]
LAGRANGE = [c for c in ALL if c.startswith("lagrange_")]
ENERGY = [c for c in ALL if c.startswith("energy_")]
SF = cm.SF_INFINITY


def load(case):
    g = np.load(os.path.join(cm.GOLDEN, case + ".npz"))
    return {k: (g[k] if g[k].ndim else g[k].item()) for k in g.files}


def test_fixture_inventory():
    assert {"plummer_n4000", "plummer_n4000_tidal", "king_w5_n3000", "king_w7_imf_bin", "ci_input", "cosmic_plummer_n100", "cosmic_king_n100"} <= set(CASES)


def _lagrange_inputs(case):
    g, base = load(case), None
    base = load(g["case"])
    n = base["N_STAR"]
    sr, sm = cm.make_tables(base["r"], base["star_m"])
    return g, base, n, sr, sm


@pytest.mark.parametrize("case", LAGRANGE)
def test_oracle_lagrange_radii_match_reference(case):
    """comp_mass_percent (reference cmc_utils.c:872-985) by the oracle against what the compiled reference produced"""
    g, base, n, sr, sm = _lagrange_inputs(case)
    assert len(LAGRANGE) >= 4
    want = ol.comp_mass_percent(sr, sm, n, n, base["vr"], base["vt"], g["mass_pc"], g["Mtotal"], cenma_m=g["cenma_m"], initial_total_mass=g["initial_total_mass"],
                                bh_include=bool(g["bh_include"]), minimum_r=g["minimum_r"])
    for f in ol.MASS_PC_FIELDS:
        cm.assert_bits_equal(want[f], g[f], f)
    assert want["filled"] == len(g["mass_pc"]) - 1 and g["mass_r"][-1] == -1.0
    assert ol.clusdyn_calculate(sr, sm, n, g["Mtotal"]) == g["clusdyn_rh"]          # clusdyn_calculate, the half-mass radius


@pytest.mark.gpu
@pytest.mark.parametrize("case", LAGRANGE)
def test_cuda_lagrange_radii_match_reference(case):
    from cmc_cosmic_b200.cluster import Cluster
    g, base, n, sr, sm = _lagrange_inputs(case)
    c = Cluster(n, n_star_dim=n + 64)
    c.load(base["r"], base["vr"], base["vt"], base["star_m"])
    c.cenma.m = g["cenma_m"]
    c.calc_potential_new()
    assert c.Mtotal == g["Mtotal"]
    c.upload()
    c.dev_comp_mass_percent(g["mass_pc"], bh_include=bool(g["bh_include"]), minimum_r=g["minimum_r"], initial_total_mass=g["initial_total_mass"])
    for f in ol.MASS_PC_FIELDS:
        cm.assert_bits_equal(getattr(c, f), g[f], f)
    assert c.dev_clusdyn_calculate() == g["clusdyn_rh"]
    c.close()


def _energy_inputs(case):
    g = load(case)
    base = load(g["case"])
    n = base["N_STAR"]
    sr, sm = cm.make_tables(base["r"], base["star_m"])
    t = ol.tables(sr, sm, np.ascontiguousarray(g["phi"]), n, n, 0.0)
    f = {k[3:]: g[k] for k in g if k.startswith("in_")}
    return g, base, n, sr, sm, t, f


@pytest.mark.parametrize("case", ENERGY)
def test_oracle_energy_scheme_and_sigma_match_reference(case):
    """energy_conservation1, set_velocities3 (reference cmc_utils.c:2153-2174, 254-352) and calc_sigma_r (cmc_dynamics_helper.c:2821-2975) by
    the oracle against what the compiled reference produced"""
    g, base, n, sr, sm, t, f = _energy_inputs(case)
    assert len(ENERGY) >= 2
    idx = np.arange(1, n + 1)
    for k, v in ol.energy_conservation1(t, idx, base["vr"], base["vt"]).items():
        cm.assert_bits_equal(v, g["ec1_" + k], k)
    io, flags, left = ol.set_velocities3_first(t, idx, f)
    for k in ol.SV3_IO:
        cm.assert_bits_equal(io[k], g["out_" + k], k)
    assert g["Eoops"] == 0.0 and left > 0
    for p_win, mave in ((20, 0), (20, 1), (3, 1)):
        a, sg = ol.calc_sigma_r(p_win, sr, sm, n, n, base["vr"], base["vt"], mave=bool(mave))
        cm.assert_bits_equal(a[1:], g[f"sigma_p{p_win}_m{mave}_a"], "r / mean mass")
        cm.assert_bits_equal(sg[1:], g[f"sigma_p{p_win}_m{mave}_sigma"], "sigma")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ENERGY)
def test_cuda_energy_scheme_matches_reference(case):
    from cmc_cosmic_b200.cluster import Cluster
    g, base, n, sr, sm, t, f = _energy_inputs(case)
    c = Cluster(n, n_star_dim=n + 64)
    c.load(base["r"], base["vr"], base["vt"], base["star_m"])
    c.calc_potential_new()
    cm.assert_bits_equal(c.star_phi[:n + 2], g["phi"], "phi")
    c.upload()
    c.dev_energy_conservation1()
    c.download()
    for k in ("vrold", "vtold", "Uoldrold"):
        cm.assert_bits_equal(c.field(k)[1:n + 1], g["ec1_" + k], k)
    for k, v in f.items():
        c.field(k)[1:n + 1] = v
    c.upload()
    c.dev_energy_conservation3()
    c.download()
    for k in ol.SV3_IO:
        cm.assert_bits_equal(c.field(k)[1:n + 1], g["out_" + k], k)
    c.close()


def expected_status(g):
    """per-star branch of get_positions_loop recovered from the reference's outputs.  remove_star ends in destroy_obj -> zero_star
    (cmc_evolution_thr.c:712, cmc_dynamics_helper.c:131-153), so a removed star's record holds no trace of WHY it was removed: unbound
    stars are recognised by E + phi_s >= 0 (:806,823) in the arithmetic of the reference, the other removed ones were tidally stripped."""
    n = g["N_STAR"]
    st = np.full(n, ol.ST_OK, np.int32)
    removed = g["rnew"] == SF
    E_tot = g["E"] + (g["star_m"] / g["r"]) * (1.0 / n)
    st[removed & (E_tot >= 0.0)] = ol.ST_UNBOUND
    st[removed & (E_tot < 0.0)] = ol.ST_TIDAL
    st[removed & (g["star_m"] < cm.ZERO)] = ol.ST_MASSLESS
    circ = ~removed & (g["rnew"] == g["r"]) & (g["X"] == 0.0)
    st[circ] = ol.ST_CIRCULAR
    return st


@pytest.mark.parametrize("case", CASES)
def test_oracle_potential_matches_reference(case):
    g = load(case)
    n = g["N_STAR"]
    want = cm.oracle_potential(g["r"], g["star_m"], n, cenma_m=g["cenma_m"])
    cm.assert_bits_equal(want["phi"][:n + 2], g["phi0"], "phi0")
    assert want["Mtotal"] == g["Mtotal0"] and want["Rtidal"] == g["Rtidal0"] and want["N_MAX"] == n


@pytest.mark.parametrize("case", CASES)
def test_oracle_orbit_step_matches_reference(case):
    g = load(case)
    n = g["N_STAR"]
    sr, sm = cm.make_tables(g["r"], g["star_m"])
    t = ol.tables(sr, sm, g["phi0"].copy(), n, n, g["cenma_m"])
    out = ol.get_positions(t, np.arange(1, n + 1), g["E"], g["J"], g["vr"], g["vt"], g["Rtidal_used"], g["draws"])
    assert int(out["draw_off"][-1]) == len(g["draws"]), "the oracle consumed a different number of draws than the reference"
    st = expected_status(g)
    assert np.array_equal(out["status"], st)
    for f in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(out[f], g[f], f)
    ok = st == ol.ST_OK
    for f in ("X", "r_peri", "r_apo"):
        cm.assert_bits_equal(out[f][ok], g[f][ok], f)
    assert out["max_rad"] == g["max_r"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_sort_and_second_potential_match_reference(case):
    g = load(case)
    n, n2 = g["N_STAR"], g["N_MAX_after"]
    star = np.zeros((n, 424), np.uint8)
    from cmc_cosmic_b200.layout import star_field
    star_field(star, "r")[:] = g["rnew"]
    star_field(star, "m")[:] = g["star_m"]
    star_field(star, "id")[:] = np.arange(1, n + 1)
    binary = None
    nbd = 0
    if "binind" in g:
        star_field(star, "binind")[:] = g["binind"]
        nb = len(g["binary_in"])
        nbd = nb + 8
        binary = np.zeros((nbd - 1, 464), np.uint8)
        binary[:nb] = g["binary_in"]
    n_new, nb_out, perm = ol.sort_stars(star, n, 424, 0, 152, binary, nbd, 464)
    assert n_new == n2
    assert np.array_equal(star_field(star, "id")[:n2], g["sorted_id"])
    if binary is not None:
        assert nb_out == g["N_b_local_after"]
        assert np.array_equal(star_field(star, "binind")[:n2], g["sorted_binind"])
        assert np.array_equal(binary[:nb_out], g["binary_after"][:nb_out])
    cm.assert_bits_equal(star_field(star, "r")[:n2].copy(), g["star_r_after"][1:n2 + 1], "sorted r")
    want = cm.oracle_potential(star_field(star, "r")[:n2].copy(), star_field(star, "m")[:n2].copy(), n, cenma_m=g["cenma_m"])
    cm.assert_bits_equal(want["phi"][:n2 + 2], g["phi_after"], "phi after the step")
    assert want["Mtotal"] == g["Mtotal_after"] and want["Rtidal"] == g["Rtidal_after"]


@pytest.mark.parametrize("case", CASES)
def test_kernel_core_on_host_matches_reference(case):
    """orbit_core.h (the code the CUDA kernel runs per star, incl. the FP64+bound / x87 replay of function_q) built for the host"""
    import hostemul
    g = load(case)
    n = g["N_STAR"]
    sr, sm = cm.make_tables(g["r"], g["star_m"])
    out = hostemul.get_positions(sr, sm, g["phi0"].copy(), n, n, g["cenma_m"], np.arange(1, n + 1), g["E"], g["J"], g["vr"], g["vt"],
                                 g["Rtidal_used"], g["draws"])
    assert out["n_draws"] == len(g["draws"])
    st = expected_status(g)
    assert np.array_equal(out["status"], st)
    for f in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(out[f], g[f], f)
    ok = st == ol.ST_OK
    for f in ("X", "r_peri", "r_apo"):
        cm.assert_bits_equal(out[f][ok], g[f][ok], f)


# ------------------------------------------------------------------------------------------------------------ GPU


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_full_step_matches_reference(case):
    """One pass of the path through the C ABI (host-buffer entry points, the ones the reference-side shim binds) on the reference's
    own inputs and its recorded taus113 draws: every output bit-identical to what the reference produced."""
    from cmc_cosmic_b200.cluster import Cluster
    g = load(case)
    n, n2 = g["N_STAR"], g["N_MAX_after"]
    has_bin = "binind" in g
    nb = len(g["binary_in"]) if has_bin else 0
    c = Cluster(n, n_star_dim=n + 64, n_bin_dim=nb + 8, min_chunk=g["min_chunk"])
    c.load(g["r"], g["vr"], g["vt"], g["star_m"], binind=g["binind"] if has_bin else None)
    if has_bin:
        c.binary[1:nb + 1] = g["binary_in"]
    c.cenma.m = g["cenma_m"]
    c.calc_potential_new()
    cm.assert_bits_equal(c.star_phi[:n + 2], g["phi0"], "phi0")
    assert c.Mtotal == g["Mtotal0"] and c.Rtidal == g["Rtidal0"]
    c.compute_energy()
    cm.assert_bits_equal(c.field("E")[1:n + 1], g["E"], "E")
    cm.assert_bits_equal(c.field("J")[1:n + 1], g["J"], "J")
    c.Rtidal = g["Rtidal_used"]
    # CSR offsets of the sequential stream: from the oracle (already pinned against the same fixture above)
    sr, sm = cm.make_tables(g["r"], g["star_m"])
    t = ol.tables(sr, sm, g["phi0"].copy(), n, n, g["cenma_m"])
    off = ol.get_positions(t, np.arange(1, n + 1), g["E"], g["J"], g["vr"], g["vt"], g["Rtidal_used"], g["draws"])["draw_off"]
    status = c.new_orbits_calculate(draws=g["draws"], draw_off=off)
    st = expected_status(g)
    assert np.array_equal(status, st)
    for f in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(c.field(f)[1:n + 1], g[f], f)
    ok = st == ol.ST_OK
    for f in ("X", "r_peri", "r_apo"):
        cm.assert_bits_equal(c.field(f)[1:n + 1][ok], g[f][ok], f)
    assert c.max_r == g["max_r"]
    # the caller's transfer (ComputeIntermediateEnergy), then sort + potential
    for dst, src in (("r", "rnew"), ("vr", "vrnew"), ("vt", "vtnew")):
        c.field(dst)[1:n + 1] = c.field(src)[1:n + 1]
    c.field("m")[1:n + 1] = c.star_m[1:n + 1]
    c.qsorts_new()
    c.post_sort_comm()
    c.calc_potential_new()
    assert c.clus.N_MAX == n2
    assert np.array_equal(c.field("id")[1:n2 + 1], g["sorted_id"])
    cm.assert_bits_equal(c.star_r[:n2 + 2], g["star_r_after"], "star_r after")
    cm.assert_bits_equal(c.star_m[1:n2 + 1], g["star_m_after"][1:n2 + 1], "star_m after")
    cm.assert_bits_equal(c.star_phi[:n2 + 2], g["phi_after"], "phi after")
    assert c.Mtotal == g["Mtotal_after"] and c.Rtidal == g["Rtidal_after"]
    if has_bin:
        assert c.N_b_local == g["N_b_local_after"]
        assert np.array_equal(c.field("binind")[1:n2 + 1], g["sorted_binind"])
        assert np.array_equal(c.binary[1:nb + 1], g["binary_after"])
    c.close()


def _snapshot_rows_equal(got, g, key):
    from cmc_cosmic_b200 import layout
    want = np.ascontiguousarray(g[key]).view(layout.SNAPSHOT_DTYPE).reshape(-1)
    assert list(g["field_names"]) == list(layout.SNAPSHOT_FIELDS)
    assert list(g["field_offsets"]) == [layout.SNAPSHOT_DTYPE.fields[n][1] for n in layout.SNAPSHOT_FIELDS]
    assert len(got) == len(want) > 0
    for n in layout.SNAPSHOT_FIELDS:
        if want[n].dtype == np.float64:
            cm.assert_bits_equal(got[n], want[n], f"{key}: {n}")
        else:
            assert np.array_equal(got[n], want[n]), (key, n)


@pytest.mark.parametrize("case", SNAPSHOT)
def test_kernel_core_snapshot_rows_match_reference(case):
    """the rows of the HDF5 snapshot table by the host build of csrc/snapshot_core.h against the rows the compiled reference's
    write_snapshot handed to H5TBmake_table (fixture: every field of every row, all stars and bh_only), bit for bit"""
    import hostemul as he
    g = load(case)
    assert len(SNAPSHOT) >= 1
    N = g["N"]
    nan_bits = int(np.ascontiguousarray(g["rows"]).view(np.uint64).reshape(len(g["rows"]), -1)[np.flatnonzero(g["star"][1:N + 1, 152:160].view(np.int64).reshape(-1) == 0)[0], 25])
    assert nan_bits == 0xFFF8000000000000
    for key, bh in (("rows", False), ("rows_bh", True)):
        got = he.snapshot_rows(g["star"], g["binary"], g["star_r"], g["star_m"], g["star_phi"], g["sigma"], N, N, 1, N, N, g["units_m"], g["units_l"],
                               g["avekernel"], nan_bits, bh_only=bh)
        _snapshot_rows_equal(got, g, key)


@pytest.mark.gpu
@pytest.mark.parametrize("case", SNAPSHOT)
def test_cuda_snapshot_rows_match_reference(case):
    """the same rows built on the device from the resident arrays (cmcb200_dev_snapshot_rows) against the reference's"""
    from cmc_cosmic_b200.cluster import Cluster
    g = load(case)
    N, nb = g["N"], g["nb"]
    c = Cluster(N, n_star_dim=N + 64, n_bin_dim=nb + 8)
    c.star[:N + 2] = g["star"]
    c.binary[:nb + 1] = g["binary"]
    c.star_r[:N + 2], c.star_m[:N + 2], c.star_phi[:N + 2] = g["star_r"], g["star_m"], g["star_phi"]
    c.N_b_local = c.clus.N_BINARY = c.bin_slots_in_use = nb
    c.upload()
    for key, bh in (("rows", False), ("rows_bh", True)):
        got = c.dev_snapshot_rows(g["units_m"], g["units_l"], sigma=g["sigma"], avekernel=g["avekernel"], bh_only=bh)
        _snapshot_rows_equal(got, g, key)
    c.close()
