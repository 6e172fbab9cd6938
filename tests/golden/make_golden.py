#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the UNMODIFIED reference (compiled by oracle/Makefile into
oracle/_ref/libcmc_ref.so) on inputs produced by the reference's OWN generators / input file.

    python tests/golden/make_golden.py          (needs /root/reference; run in the build container)

The reference has no golden vector for the sort / potential / new-orbit path, so these files are what pins
the oracle restatement (oracle/cmc_oracle.c) and the CUDA path on machines where /root/reference does not exist.
Cases
  plummer_n4000   COSMIC sample/cmc/elson.py draw_r_vr_vt(N, r_max=100, gamma=4), np.random.seed(1234), m = 1/N
                  (examples/PlummerSphere.ini, BASELINE.json configs[1] at small N)
  king_w5_n3000   COSMIC sample/cmc/king.py draw_r_vr_vt(N, w_0=5), np.random.seed(12345), m = 1/N (examples/KingProfile.ini)
  king_w7_imf_bin king.py w_0=7, N=3000, power-law masses, 10 % of the stars carry a binary_t (configs[4] at small N)
  ci_input        ci/input.hdf5 as it is (N = 10 000, Kroupa masses, 1 967 binaries; BASELINE.json configs[0]); the
                  CLUS_OBJ_DATA table is read straight from the file bytes (no HDF5 library in this image)
  cosmic_plummer_n100 / cosmic_king_n100   the N=100 vectors of COSMIC's own unit tests (tests/data/cmc_*_test.npz)
  lagrange_<case> comp_mass_percent (the Lagrange radii) on the initial state of <case>, `make_golden.py lagrange`
  snapshot_imf_bin  write_snapshot's rows (the HDF5 calls recorded), `make_golden.py snapshot`
  energy_<case>   energy_conservation1, energy_conservation3 -> set_velocities3 and calc_sigma_r (8f N3) on the initial state
                  of <case> with the scheme's inputs of tests/common.py:sv3_fields, `make_golden.py energy`
Every case records ONE full pass of the path by the reference:
  calc_potential_new -> ComputeEnergy -> new_orbits_calculate (taus113 draws recorded) -> r <- rnew (ComputeIntermediateEnergy's
  transfer) -> qsorts_new -> post_sort_comm -> calc_potential_new
with all inputs and all outputs.  A second set of arrays holds a tidally truncated run (small Rtidal) to reach the
TIDAL / UNBOUND branches.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refharness as rh  # noqa: E402

REF = os.environ.get("CMC_REFERENCE", "/root/reference")
SAMPLE = os.path.join(REF, "src/bse_wrap/bse/COSMIC/src/cosmic/sample/cmc")
STAR_FIELDS_F = ("r", "vr", "vt", "m", "E", "J", "rnew", "vrnew", "vtnew", "rOld", "X", "r_peri", "r_apo")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(SAMPLE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def read_ci_input():
    """CLUS_OBJ_DATA/block0_values of ci/input.hdf5: contiguous little-endian float64[10002][8] (columns id, k, m, Reff, r, vr, vt,
    binind; src/libs/fitslib.c:381-390), located by its id column pattern."""
    raw = open(os.path.join(REF, "ci/input.hdf5"), "rb").read()
    pat = np.array([1.0, 2.0, 3.0]).tobytes()
    off = None
    for cand in range(0, len(raw) - 200, 8):
        if raw[cand:cand + 8] == pat[:8] and raw[cand + 64:cand + 72] == pat[8:16] and raw[cand + 128:cand + 136] == pat[16:24]:
            off = cand - 64
            break
    assert off is not None
    a = np.frombuffer(raw, dtype="<f8", count=10002 * 8, offset=off).reshape(10002, 8)
    assert a[0, 4] == np.finfo(float).tiny and a[-1, 4] == 1e40 and abs(a[:, 2].sum() - 1.0) < 1e-12 and np.all(np.diff(a[1:-1, 4]) > 0)
    return a


def run_case(R, name, r, vr, vt, obj_m, binind=None, seed=1234, min_chunk=40, rtidal_override=None, cenma_m=0.0, samplesize=1024):
    n = len(r)
    n_bin = int((binind > 0).sum()) if binind is not None else 0
    R.init(n, n_star_dim=n + 64, n_bin_dim=n_bin + 8, min_chunk=min_chunk, idum=seed, samplesize=samplesize)
    sm = obj_m * n                                         # star_m = obj_m * N_STAR ... see cmc_fits.c:123-136 (units.mstar = 1)
    R.load_cluster(r, vr, vt, sm, binind=binind)
    if n_bin:
        bb = R.binary_bytes()
        # recognisable, compressible content: 8-byte word w of binary[i] holds i * 64 + w
        words = (np.arange(1, n_bin + 1, dtype=np.int64)[:, None] * 64 + np.arange(R.sizeof_binary // 8, dtype=np.int64)[None, :])
        bb[1:n_bin + 1] = words.view(np.uint8).reshape(n_bin, R.sizeof_binary)
    R.set_scalars(n, n, N_BINARY=n_bin, cenma_m=cenma_m)
    R.L.ref_calc_potential_new()
    R.compute_energy()
    sc0 = R.scalars()
    out = dict(r=r.copy(), vr=vr.copy(), vt=vt.copy(), star_m=sm.copy(), N_STAR=n, min_chunk=min_chunk, seed=seed, cenma_m=cenma_m,
               phi0=R.star_phi()[:n + 2].copy(), Mtotal0=sc0["Mtotal"], Rtidal0=sc0["Rtidal"],
               E=R.star_field("E")[1:n + 1].copy(), J=R.star_field("J")[1:n + 1].copy())
    if binind is not None:
        out["binind"] = binind.astype(np.int64)
        out["binary_in"] = R.binary_bytes()[1:n_bin + 1].copy()
    if rtidal_override is not None:
        R.set_scalars(n, n, N_BINARY=n_bin, cenma_m=cenma_m, Rtidal=rtidal_override, Mtotal=sc0["Mtotal"])
    out["Rtidal_used"] = R.scalars()["Rtidal"]
    R.L.ref_draw_mode(1)
    R.L.ref_new_orbits_calculate()
    R.L.ref_draw_mode(0)
    out["draws"] = R.draws()
    for f in ("rnew", "vrnew", "vtnew", "X", "r_peri", "r_apo"):
        out[f] = R.star_field(f)[1:n + 1].copy()
    sc1 = R.scalars()
    out["max_r"] = sc1["max_r"]
    out["TidalMassLoss"] = sc1["TidalMassLoss"]
    # the transfer that creates the sort key (ComputeIntermediateEnergy, cmc_utils.c:380-390)
    srr, smm = R.star_r(), R.star_m()
    for dst, src in (("r", "rnew"), ("vr", "vrnew"), ("vt", "vtnew")):
        R.star_field(dst)[1:n + 1] = R.star_field(src)[1:n + 1]
    R.star_field("rOld")[1:n + 1] = srr[1:n + 1]
    R.star_field("m")[1:n + 1] = smm[1:n + 1]
    R.L.ref_qsorts_new()
    R.L.ref_post_sort_comm()
    R.L.ref_calc_potential_new()
    sc2 = R.scalars()
    n2 = int(sc2["N_MAX"])
    out["N_MAX_after"] = n2
    out["sorted_id"] = R.star_field("id", np.int64)[1:n2 + 1].copy()
    out["sorted_binind"] = R.star_field("binind", np.int64)[1:n2 + 1].copy()
    out["star_r_after"] = R.star_r()[:n2 + 2].copy()
    out["star_m_after"] = R.star_m()[:n2 + 2].copy()
    out["phi_after"] = R.star_phi()[:n2 + 2].copy()
    out["Mtotal_after"], out["Rtidal_after"] = sc2["Mtotal"], sc2["Rtidal"]
    out["N_b_local_after"] = int(sc2["N_b_local"])
    if n_bin:
        out["binary_after"] = R.binary_bytes()[1:n_bin + 1].copy()
    R.free()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    st_ok = np.isfinite(out["rnew"]) & (out["rnew"] < 1e10)
    print(f"{name}: N={n} draws={len(out['draws'])} ({(len(out['draws']) / n - 1) / 2:.2f} tries/star) removed={n - int(st_ok.sum())} "
          f"N_MAX_after={n2} n_bin={n_bin}->{out['N_b_local_after']}")


LAGRANGE_PCS = np.array([1e-4, 0.001, 0.005, 0.01, 0.03, 0.05, 0.07, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.99, 1.1])  # examples' MASS_PC + edges


def lagrange_cases(R):
    """lagrange_<case>.npz: comp_mass_percent (reference src/cmc/cmc_utils.c:872-985) run by the reference on the initial state of an existing
    fixture (after calc_potential_new), without and with MASS_PC_BH_INCLUDE + a central mass.  Inputs are those of <case>.npz."""
    for case, cenma, bh in (("plummer_n4000", 0.0, 0), ("king_w7_imf_bin", 0.0, 0), ("ci_input", 0.0, 0), ("king_w5_n3000", 0.003, 1)):
        g = np.load(os.path.join(HERE, case + ".npz"))
        n = int(g["N_STAR"])
        R.init(n, n_star_dim=n + 64, n_bin_dim=8, min_chunk=int(g["min_chunk"]), idum=int(g["seed"]))
        R.load_cluster(g["r"], g["vr"], g["vt"], g["star_m"])
        R.set_scalars(n, n, cenma_m=cenma * n)
        R.L.ref_calc_potential_new()
        out = R.comp_mass_percent(LAGRANGE_PCS, bh_include=bh, minimum_r=1e-6, initial_total_mass=0.98)
        out["clusdyn_rh"] = R.clusdyn_rh()        # clusdyn_calculate (cmc_utils.c:1613-1624)
        out.update(mass_pc=LAGRANGE_PCS, cenma_m=cenma * n, bh_include=bh, minimum_r=1e-6, initial_total_mass=0.98, Mtotal=R.scalars()["Mtotal"], case=case)
        R.free()
        np.savez_compressed(os.path.join(HERE, f"lagrange_{case}.npz"), **out)
        print(f"lagrange_{case}: mass_r = {out['mass_r'][[0, 7, 11, 15]]} no_star_r = {out['no_star_r'][[0, 7, 11, 15]]}")


def energy_cases(R):
    """energy_<case>.npz: the reference's energy_conservation1 (cmc_utils.c:2153-2174), energy_conservation3 -> set_velocities3 (:254-363) and
    calc_sigma_r (cmc_dynamics_helper.c:2821-2975; AVEKERNEL = 20 and BH_AVEKERNEL = 3, both output modes) on the initial state of an
    existing fixture after calc_potential_new.  The scheme's inputs (what the previous timestep left in star[]) come from
    tests/common.py:sv3_fields and are stored with the outputs."""
    import common as cm
    for case in ("plummer_n4000", "king_w7_imf_bin"):
        g = np.load(os.path.join(HERE, case + ".npz"))
        n = int(g["N_STAR"])
        R.init(n, n_star_dim=n + 64, n_bin_dim=8, min_chunk=int(g["min_chunk"]), idum=int(g["seed"]))
        R.load_cluster(g["r"], g["vr"], g["vt"], g["star_m"])
        R.set_scalars(n, n)
        R.L.ref_calc_potential_new()
        phi = R.star_phi()[:n + 2].copy()
        out = dict(case=case, phi=phi)
        R.L.ref_energy_conservation1()
        for k in ("vrold", "vtold", "Uoldrold"):
            out["ec1_" + k] = R.star_field(k)[1:n + 1].copy()
        for p_win, mave in ((20, 0), (20, 1), (3, 1)):
            a, sg = R.calc_sigma_r(p_win, n, bool(mave))
            out[f"sigma_p{p_win}_m{mave}_a"], out[f"sigma_p{p_win}_m{mave}_sigma"] = a[1:].copy(), sg[1:].copy()
        f = cm.sv3_fields(g["r"], g["vr"], g["vt"], phi[1:n + 1], seed=n)
        for k, v in f.items():
            R.star_field(k, np.int64 if k == "interacted" else np.float64)[1:n + 1] = v
            out["in_" + k] = v
        R.L.ref_set_Eoops(0.0)
        R.L.ref_energy_conservation3()
        for k in ("vr", "vt", "vrnew", "vtnew"):
            out["out_" + k] = R.star_field(k)[1:n + 1].copy()
        out["Eoops"] = R.L.ref_get_Eoops()
        R.free()
        np.savez_compressed(os.path.join(HERE, f"energy_{case}.npz"), **out)
        print(f"energy_{case}: {int((out['out_vr'] != f['vr']).sum())} of {n} velocities corrected, sigma[n/2] = {out['sigma_p20_m0_sigma'][n // 2]:.6f}")


def lagrad_run():
    """plummer_lagrad_run.npz: the Lagrange radii of the reference's OWN example run to core collapse (examples/PlummerSphere.ini; its output
    docs/sphinx/source/example_output/plummer.lagrad.dat, 4 373 timesteps, 25 mass fractions), averaged over windows of +-0.01 time units
    around t = 0, 0.05, ..., 1.70 -- what "Lagrangian radii from a full run must statistically agree with the reference run" is checked against."""
    path = os.path.join(REF, "docs/sphinx/source/example_output/plummer.lagrad.dat")
    hdr = open(path).readlines()[1]
    frac = np.array([float(x.split("(")[1].rstrip(")")) for x in hdr.split()[1:]])
    d = np.loadtxt(path)
    times = np.arange(0.0, 1.7001, 0.05)
    rows, count = [], []
    for t in times:
        sel = np.abs(d[:, 0] - t) <= 0.01
        rows.append(d[sel, 1:].mean(axis=0))
        count.append(int(sel.sum()))
    np.savez_compressed(os.path.join(HERE, "plummer_lagrad_run.npz"), mass_pc=frac, t=times, r=np.array(rows), rows_averaged=np.array(count),
                        t_end=d[-1, 0], n_steps=len(d))
    print("plummer_lagrad_run.npz:", len(times), "times x", len(frac), "fractions; run ends at t =", d[-1, 0], "after", len(d), "steps")


def snapshot_case(R):
    """snapshot_imf_bin: the rows write_snapshot() (cmc_io.c:2960-3345, compiled unmodified, HDF5 calls recorded by oracle/ref_harness.c) hands
    to H5TBmake_table for a 1 000-star cluster with 120 binaries and random content in every field it reads -- all stars and bh_only --
    with its inputs: star[] / binary[] bytes, the three tables, sigma_array.sigma, the units (`make_golden.py snapshot`)"""
    sys.path.insert(0, os.path.dirname(HERE))
    import test_oracle_vs_ref as tv
    N, nb = 1000, 120
    sigma = tv._snapshot_state(R, N, nb, seed=2024)
    units_m, units_l = 1.989e33 * 2.5e5, 3.0857e18 * 2.1
    rows, geom = R.write_snapshot(units_m, units_l, sigma=sigma, avekernel=20)
    rows_bh, _ = R.write_snapshot(units_m, units_l, sigma=sigma, avekernel=20, bh_only=True)
    out = dict(N=N, nb=nb, star=R.star_bytes()[:N + 2].copy(), binary=R.binary_bytes()[:nb + 1].copy(), star_r=R.star_r()[:N + 2].copy(),
               star_m=R.star_m()[:N + 2].copy(), star_phi=R.star_phi()[:N + 2].copy(), sigma=sigma, units_m=units_m, units_l=units_l, avekernel=20,
               rows=rows.view(np.uint8).reshape(len(rows), -1), rows_bh=rows_bh.view(np.uint8).reshape(len(rows_bh), -1),
               field_names=np.array([g[0] for g in geom]), field_offsets=np.array([g[1] for g in geom]), field_types=np.array([g[2] for g in geom]))
    np.savez_compressed(os.path.join(HERE, "snapshot_imf_bin.npz"), **out)
    print("snapshot_imf_bin.npz:", len(rows), "rows,", len(rows_bh), "bh_only rows,", (rows["binflag"] == 1).sum(), "binaries")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "lagrad":     # the reference's example run (a data file of the reference, no build needed)
        lagrad_run()
        return
    assert rh.available(), "build oracle/_ref first (make -C oracle)"
    if len(sys.argv) > 1 and sys.argv[1] == "snapshot":   # only the snapshot-table fixture
        snapshot_case(rh.Ref())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "energy":     # only the energy-scheme / sigma fixtures
        energy_cases(rh.Ref())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "lagrange":   # only the comp_mass_percent fixtures (the others are their inputs)
        lagrange_cases(rh.Ref())
        return
    el, kg = _load("elson"), _load("king")
    R = rh.Ref()
    np.random.seed(1234)
    r, vr, vt = el.draw_r_vr_vt(N=4000, r_max=100, gamma=4)
    o = np.argsort(r, kind="stable")
    run_case(R, "plummer_n4000", r[o], vr[o], vt[o], np.full(4000, 1.0 / 4000), seed=1234)
    run_case(R, "plummer_n4000_tidal", r[o], vr[o], vt[o], np.full(4000, 1.0 / 4000), seed=77, rtidal_override=4.0)
    np.random.seed(12345)
    r, vr, vt = kg.draw_r_vr_vt(N=3000, w_0=5)
    o = np.argsort(r, kind="stable")
    run_case(R, "king_w5_n3000", r[o], vr[o], vt[o], np.full(3000, 1.0 / 3000), seed=12345)
    np.random.seed(7)
    r, vr, vt = kg.draw_r_vr_vt(N=3000, w_0=7)
    o = np.argsort(r, kind="stable")
    rng = np.random.default_rng(7)
    m = (0.08 ** -1.3 + rng.uniform(size=3000) * (150.0 ** -1.3 - 0.08 ** -1.3)) ** (-1 / 1.3)
    m /= m.sum()
    binind = np.zeros(3000, np.int64)
    who = rng.choice(3000, 300, replace=False)
    binind[who] = rng.permutation(np.arange(1, 301))
    run_case(R, "king_w7_imf_bin", r[o], vr[o], vt[o], m, binind=binind, seed=7, cenma_m=0.0)
    a = read_ci_input()
    run_case(R, "ci_input", a[1:-1, 4].copy(), a[1:-1, 5].copy(), a[1:-1, 6].copy(), a[1:-1, 2].copy(), binind=a[1:-1, 7].astype(np.int64), seed=1729217,
             min_chunk=40)    # ci/Params.ini:524,536
    for kind in ("plummer", "king"):
        g = np.load(os.path.join(REF, f"src/bse_wrap/bse/COSMIC/src/cosmic/tests/data/cmc_{kind}_test.npz"))
        r, vr, vt = g["arr_0"], g["arr_1"], g["arr_2"]
        o = np.argsort(r, kind="stable")
        run_case(R, f"cosmic_{kind}_n100", r[o], vr[o], vt[o], np.full(100, 0.01), seed=2, min_chunk=20, samplesize=16)


if __name__ == "__main__":
    main()
