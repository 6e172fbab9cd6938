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


def with_new_stars(r, vr, vt, sm, pot, n_new, seed=0):
    """A cluster state plus `n_new` stars created this step (collision products / disrupted binaries in CMC): they sit behind the old
    local stars in star[], their r / m are appended to the tables at N_MAX + 2 .. (get_global_idx, reference cmc_utils.c:2384-2407), their
    position is star[].rnew and calc_orbit_new looks their index up with FindZero_r (cmc_orbit.c:39-41).
    Returns tables (sr, sm_t, phi) sized N + 2 + n_new, per-star arrays E, J, vr, vt over all N + n_new stars, g_idx, is_new, rnew_in."""
    N = len(r)
    rng = np.random.default_rng(seed)
    src = rng.choice(N, n_new, replace=False)
    r_new = r[src] * rng.uniform(0.999, 1.001, n_new)
    sr, smt, phi = (np.concatenate([pot[k][:N + 2], np.zeros(n_new)]) for k in ("star_r", "star_m", "phi"))
    sr[N + 2:], smt[N + 2:] = r_new, sm[src] * rng.uniform(0.5, 2.0, n_new)
    t = ol.tables(sr, smt, phi, N, N, 0.0)
    phi_new = np.array([ol.lib().orc_potential(ol.C.byref(t), float(x)) for x in r_new])
    vr_a, vt_a = np.concatenate([vr, vr[src]]), np.concatenate([vt, vt[src]])
    E = np.concatenate([pot["phi"][1:N + 1] + 0.5 * (vr * vr + vt * vt), phi_new + 0.5 * (vr[src] ** 2 + vt[src] ** 2)])
    J = np.concatenate([r * vt, r_new * vt[src]])
    g_idx = np.concatenate([np.arange(1, N + 1), N + 2 + np.arange(n_new)]).astype(np.int64)
    is_new = np.concatenate([np.zeros(N, np.uint8), np.ones(n_new, np.uint8)])
    rnew_in = np.concatenate([np.zeros(N), r_new])
    return sr, smt, phi, E, J, vr_a, vt_a, g_idx, is_new, rnew_in


def sv3_fields(r, vr, vt, phi, seed=0, src_frac=0.004, long_chains=True):
    """Inputs of set_velocities3 (reference cmc_utils.c:254-352) for a sorted cluster with potential phi[1..N]: the fields the energy scheme
    left in star[] (interacted, rOld, vrold, vtold, Uoldrold, Uoldrnew) plus the current (vr, vt, vrnew, vtnew).  Built so that every
    branch fires: interacted stars, stars whose predicted speed is unphysical (vnew2 <= 0: they feed the excess-energy carry) alone, in
    runs, with deficits large enough that the carry crosses thousands of stars, one of them right at the end (the rank keeps a
    leftover), rOld below star_r[1] and beyond star_r[N]."""
    N = len(r)
    rng = np.random.default_rng(seed)
    f = dict(interacted=(rng.uniform(size=N) < 0.03).astype(np.int64))
    f["rOld"] = r * np.exp(rng.normal(0.0, 0.05, N))
    f["rOld"][rng.choice(N, 3, replace=False)] = [r[0] * 0.5, r[-1] * 1.5, r[-1] * 40.0]
    f["vrold"] = vr * rng.uniform(0.97, 1.03, N)
    f["vtold"] = vt * rng.uniform(0.97, 1.03, N)
    f["Uoldrold"] = phi * rng.uniform(0.999, 1.001, N)
    f["Uoldrnew"] = phi * rng.uniform(0.999, 1.001, N)
    src = rng.choice(N, max(3, int(src_frac * N)), replace=False)
    src = np.concatenate([src, np.clip(src[:len(src) // 4] + 1, 0, N - 1), [N - 1]])   # some in runs of two, one at the very end
    f["vrold"][src] *= 0.01
    f["vtold"][src] *= 0.01
    f["Uoldrold"][src] = f["Uoldrnew"][src] + phi[src] * rng.uniform(0.01, 0.5, len(src))  # phi < 0: the predicted v^2 turns negative
    if long_chains:
        big = src[:3]
        f["Uoldrold"][big] = f["Uoldrnew"][big] + phi[big] * 2000.0
    f["vr"], f["vt"] = vr.copy(), vt.copy()
    f["vrnew"], f["vtnew"] = vr * 1.5, vt * 0.5           # overwritten for every corrected star, kept for the others
    return f


def brent_cluster(N=20000, seed=5, stride=20):
    """A cluster state in which find_root_vr (reference cmc_nr.c:500-603; cmc_orbit.c:459-474, 485-500) actually fires.  On an ordinary
    cluster it never does: Henon's closed form needs inside_sqrt = a^2 - 2 J^2 (b - E) >= 0 on the bracketing interval, which holds whenever Q
    changes sign across it -- except by rounding.  So every `stride`-th star j is given the (E, J) of the orbit that is CIRCULAR AT ITS OWN
    RADIUS for the interpolation Phi = a/r + b of a neighbouring interval (J^2 = -a r_j, E = b + a / (2 r_j), nudged by -4 .. +4 ulps): the
    quadratic has a double root at the grid point r_j, Q(r_j) is a few ulps from zero, inside_sqrt comes out negative about as often as not
    and the vertex -J^2/a falls outside the interval by rounding -- the Brent fallback.  Returns r, vr, vt, star_m, pot, E (star[].E, i.e.
    WITHOUT the self term get_positions_loop adds, cmc_evolution_thr.c:806), J, planted (bool mask)."""
    r, vr, vt, sm = cluster_state("plummer", N, seed=seed, masses="imf")
    pot = oracle_potential(r, sm)
    sr, smt, phi = pot["star_r"], pot["star_m"], pot["phi"]
    E, J = energies(phi[1:N + 1], r, vr, vt)
    planted = np.zeros(N, bool)

    def phi_s(rad, j):   # MPI_PHI_S, include/cmc/cmc.h:2175
        return (smt[j] / rad if rad >= sr[j] else smt[j] / sr[j]) * (1.0 / N)

    for n, j in enumerate(range(100, N - 100, stride)):
        k = j - 1 if n % 2 == 0 else j          # the interval below the star (r_min bracket) or above it (r_max bracket)
        rk, rk1 = sr[k], sr[k + 1]
        Uk, Uk1 = phi[k] + phi_s(rk, j), phi[k + 1] + phi_s(rk1, j)
        a = (Uk1 - Uk) / (1. / rk1 - 1. / rk)   # set_a_b, cmc_orbit.c:722-740
        b = (Uk / rk1 - Uk1 / rk) / (1. / rk1 - 1. / rk)
        Jj = np.sqrt(-a * sr[j])
        Ej = b + a / (2 * sr[j])
        d = (n // 2) % 9 - 4
        for _ in range(abs(d)):
            Ej = np.nextafter(Ej, np.inf if d > 0 else -np.inf)
        E[j - 1], J[j - 1] = Ej - phi_s(sr[j], j), Jj
        vr[j - 1], vt[j - 1] = 0.0, Jj / sr[j]
        planted[j - 1] = True
    # such an orbit can also be so thin that the rejection sampling never accepts (N_TRY exceeded: fatal in the reference,
    # cmc_evolution_thr.c:890-893): those stars get their ordinary E, J back
    E0, J0 = energies(phi[1:N + 1], r, vr, vt)
    t = ol.tables(sr, smt, phi, N, N)
    draws = ol.Taus113(17).draws(200 * N)
    for _ in range(4):
        st = ol.get_positions(t, np.arange(1, N + 1), E, J, vr, vt, pot["Rtidal"], draws)["status"]
        bad = np.nonzero(st == 5)[0]
        if not bad.size:
            break
        E[bad], J[bad], planted[bad] = E0[bad], J0[bad], False
    return r, vr, vt, sm, pot, E, J, planted


DYN_CASES = {
    # examples/PlummerSphere.ini: relaxation only; odd N (the last star stays unpaired)
    "relax_only": dict(kind="plummer", N=20001, masses="imf", nb=0, dt=2e-3, cfg=dict(relaxation=1)),
    # the switches of examples/KingProfile.ini with radii and a time step large enough for encounters of every type
    "encounters": dict(kind="king", N=20000, masses="imf", nb=1500, dt=4e-3,
                       cfg=dict(relaxation=1, ss_collision=1, binsingle=1, binbin=1, bh_capture=1, tidal_capture=1, tc_factor=3.0, wd_tc=2.5)),
    "no_relaxation": dict(kind="plummer", N=6000, masses="equal", nb=0, dt=1e-3, cfg=dict(relaxation=0, ss_collision=1)),
    "threebb_skips": dict(kind="plummer", N=8000, masses="imf", nb=0, dt=2e-3, cfg=dict(relaxation=1), threebb=True),
    "large": dict(kind="plummer", N=600000, masses="imf", nb=20000, dt=1e-3, cfg=dict(relaxation=1, ss_collision=1, binsingle=1, binbin=1)),
}


def dynamics_case(case, seed=11, first=1, nl=None):
    """the state dynamics_apply's pair loop reads, for one of DYN_CASES: tables with the oracle's potential, the local stars' fields (a shard
    first .. first + nl - 1 of the cluster when given), stellar types, radii, binaries, the switches, three draws per pair"""
    c = DYN_CASES[case]
    N, nb = c["N"], c["nb"]
    r, vr, vt, sm = cluster_state(c["kind"], N, seed=N + seed, masses=c["masses"])
    rng = np.random.default_rng(seed)
    binind = np.zeros(N, np.int64)
    if nb:
        binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    se_k = rng.choice(np.array([0, 1, 1, 1, 3, 7, 10, 11, 13, 14, 14], np.int32), N).astype(np.int32)
    rad = rng.uniform(2e-4, 3e-3, N)
    bin_a = np.zeros(nb + 1)
    bin_a[1:] = rng.uniform(1e-4, 4e-3, nb)
    threebb = None
    if c.get("threebb"):
        threebb = np.zeros(N, np.int64)
        tri = rng.choice(N // 3 - 2, 40, replace=False) * 3
        for k in range(3):
            threebb[tri + k] = 1
        threebb[N - 30:] = 0
    pot = oracle_potential(r, sm)
    E, J = energies(pot["phi"][1:N + 1], r, vr, vt)
    cfg = dict(ol.DYN_DEFAULTS)
    cfg.update(c["cfg"])
    nl = N - first + 1 if nl is None else nl
    sl = slice(first - 1, first - 1 + nl)
    f = dict(vr=vr[sl].copy(), vt=vt[sl].copy(), E=E[sl].copy(), J=J[sl].copy(), Y=np.zeros(nl), binind=binind[sl].copy(), se_k=se_k[sl].copy(),
             rad=rad[sl].copy())
    if threebb is not None:
        f["threebb"] = threebb[sl].copy()
    draws = np.random.default_rng(seed + 1).random(3 * (nl // 2))
    return dict(N=N, nb=nb, dt=c["dt"], cfg=cfg, f=f, bin_a=bin_a, r=r, sm=sm, pot=pot, draws=draws, first=first, nl=nl)


def adversarial_sigma_inputs(kind, n, rng):
    """(star_m * N units, vr, vt) built to stress the binade bookkeeping of csrc/sigma_scan.h"""
    m = np.ones(n)
    vr, vt = rng.normal(0, 0.5, n), np.abs(rng.normal(0, 0.5, n))
    if kind == "hover_at_power_of_two":      # window sums that sit ON a binade boundary and wobble across it by a few ulps
        p = 20
        m[:] = 1.0
        v2 = np.full(n, 2.0 ** -10 * n / (2 * p))        # m madhoc v^2 = 2^-10 / 2p: the 2p terms of a window sum to 2^-10 ...
        v2 *= 1.0 + rng.integers(-3, 4, n) * 2.0 ** -50   # ... give or take a few ulps
        vr, vt = np.sqrt(v2), np.zeros(n)
    elif kind == "zeros":                    # stretches of stars at rest: the m v^2 chain drops to exactly 0 and climbs back
        z = rng.random(n) < 0.3
        z[1000:1100] = True
        vr[z], vt[z] = 0.0, 0.0
    elif kind == "wide_range":               # masses over six decades, a few stars that outweigh their whole window by 10^4
        m = 10.0 ** rng.uniform(-3, 3, n)
        m[rng.choice(n, 30, replace=False)] *= 1.0e4
    elif kind == "steep_trend":              # v^2 falling by 2^-40 across the shard: the value walks down forty binades
        f = 2.0 ** (-40.0 * np.arange(n) / n)
        vr, vt = vr * np.sqrt(f), vt * np.sqrt(f)
    elif kind == "rising_trend":             # ... and up, steeply: every binade above y_in doubles a table
        f = 2.0 ** (400.0 * np.arange(n) / n)          # 13 binades per tile: more than a tile table holds
        vr, vt = vr * np.sqrt(f) * 1e-60, vt * np.sqrt(f) * 1e-60
    elif kind == "spikes":                   # isolated stars a million times faster than their neighbours
        k = rng.choice(n, 200, replace=False)
        vr[k] *= 1.0e3
    return m, vr, vt
