"""Shared helpers for the parity tests: build cluster states, run the oracle, compare."""
import os
import numpy as np

import oraclelib as ol
from cmc_cosmic_b200 import models

SF_INFINITY, ZERO = 1.0e10, 1.0e-20
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_tables(r, m_star_units, extra=0):
    """star_r / star_m with sentinels, room for N+2+extra entries."""
    n = len(r)
    sr = np.zeros(n + 2 + extra); sm = np.zeros(n + 2 + extra)
    sr[0] = ZERO
    sr[1:n + 1] = r; sm[1:n + 1] = m_star_units
    sr[n + 1] = SF_INFINITY
    return sr, sm


def cluster_state(kind, N, seed=1234, masses="equal"):
    """(r, vr, vt, star_m) radius-sorted, star_m = obj_m * N_STAR."""
    if kind == "plummer":
        r, vr, vt, m = models.plummer(N, seed=seed)
    elif kind == "king":
        r, vr, vt, m = models.king(N, w0=5.0, seed=seed)
    else:
        raise ValueError(kind)
    if masses == "imf":
        m = models.kroupa_like_masses(N, seed=seed + 1)
    return r, vr, vt, m * N


def oracle_potential(r, star_m, N_STAR=None, cenma_m=0.0, cenma_m_new=0.0, orbit_r=1e6, N_MAX_in=None):
    n = len(r)
    sr, sm = make_tables(r, star_m)
    res = ol.potential_calculate(sr, sm, n if N_MAX_in is None else N_MAX_in, N_STAR or n, cenma_m, cenma_m_new, orbit_r)
    res["star_r"], res["star_m"] = sr, sm
    return res


def energies(phi, r, vr, vt):
    """E, J as ComputeEnergy (reference cmc_utils.c:614-615)."""
    return phi + 0.5 * (vr * vr + vt * vt), r * vt


def oracle_orbits(sr, sm, phi, N_MAX, N_STAR, E, J, vr, vt, Rtidal, seed=1234, cenma_m=0.0, g_idx=None, is_new=None, rnew_in=None):
    """Run the oracle's orbit step on one sequential taus113 stream; returns its outputs + the CSR draw offsets."""
    n = len(E)
    t = ol.tables(sr, sm, phi, N_MAX, N_STAR, cenma_m)
    g = np.arange(1, n + 1, dtype=np.int64) if g_idx is None else g_idx
    rng = ol.Taus113(seed)
    # generous stream; the oracle reports how much it used
    nd = 64 * n + 1024
    draws = rng.draws(nd)
    out = ol.get_positions(t, g, E, J, vr, vt, Rtidal, draws, is_new=is_new, rnew_in=rnew_in)
    used = int(out["draw_off"][-1])
    assert used < nd
    out["draws"] = draws[:used].copy()
    return out


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    av, bv = a.view(np.uint64) if a.dtype == np.float64 else a, b.view(np.uint64) if b.dtype == np.float64 else b
    bad = np.nonzero(av != bv)[0]
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} differ, first at {bad[0]}: {a[bad[0]]!r} vs {b[bad[0]]!r}"
