"""BASELINE.json's full-size configurations, BIT-EXACT against the oracle (not through properties: that is tests/test_gpu_fullsize.py):

  configs[2]  King W0 = 5, N = 1e6
  configs[3]  Plummer sphere N = 1e7  -- the bench workload -- two steps
  configs[4]  King W0 = 7, N = 2e6, IMF masses, 10 % primordial binaries

Per step, with the oracle's own draws fed in as an input array: potential -> E, J -> orbit step (status, k_min, k_max of every star) ->
r <- rnew -> sort (the permutation, through star[].id) -> potential; every record field the path writes, the replicated tables, the
binaries.  What is size-dependent in the CUDA path is asserted not to have saturated: the exact scan never fell back to its sequential
form (event capacity, csrc/seqscan.cuh) and the deferred linear scans stayed inside their queue (ORBIT_DEFER_CAP, csrc/orbit.cuh).
The single-core oracle needs about 25 s per step at N = 1e7."""
import numpy as np
import pytest

import common as cm
import oraclelib as ol
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster

pytestmark = pytest.mark.gpu
ORBIT_DEFER_CAP = 65536          # csrc/orbit.cuh


def _oracle_step(sr, smt, phi, N_MAX, N_STAR, E, J, vr, vt, Rtidal, seed):
    n = len(E)
    t = ol.tables(sr, smt, phi, N_MAX, N_STAR, 0.0)
    nd = 8 * n + (1 << 20)
    draws = ol.Taus113(seed).draws(nd)
    out = ol.get_positions(t, np.arange(1, n + 1, dtype=np.int64), E, J, vr, vt, Rtidal, draws)
    used = int(out["draw_off"][-1])
    assert used < nd
    out["draws"] = draws[:used]
    return out


@pytest.mark.parametrize("name,kind,w0,N,imf,bin_frac,steps", [
    ("C3 King W0=5 N=1e6", "king", 5.0, 1_000_000, False, 0.0, 1),
    ("C5 King W0=7 N=2e6 IMF 10% binaries", "king", 7.0, 2_000_000, True, 0.1, 1),
    ("C4 Plummer N=1e7", "plummer", 0.0, 10_000_000, False, 0.0, 2),
])
def test_full_size_step_bit_exact(name, kind, w0, N, imf, bin_frac, steps):
    if kind == "plummer":
        r, vr, vt, m = models.plummer(N, seed=1234)
    else:
        r, vr, vt, m = models.king(N, w0=w0, seed=12345)
    sm = (models.kroupa_like_masses(N, seed=12346) if imf else m) * N
    nb = int(bin_frac * N)
    binind = None
    if nb:
        rng = np.random.default_rng(9)
        binind = np.zeros(N, np.int64)
        binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    c = Cluster(N, n_star_dim=N + 4096, n_bin_dim=nb + 64 if nb else None)
    c.load(r, vr, vt, sm, binind=binind)
    if nb:      # every binary record carries its own number in each of its 58 words: it must arrive whole behind its star
        c.binary[1:nb + 1] = (np.arange(1, nb + 1, dtype=np.int64)[:, None] * 64 + np.arange(58, dtype=np.int64)[None, :]).view(np.uint8).reshape(nb, 464)
    c.calc_potential_new()
    pot = cm.oracle_potential(r, sm)
    cm.assert_bits_equal(c.star_phi[:N + 2], pot["phi"][:N + 2], f"{name}: initial potential")
    assert c.potential_stats()[1] == 0, "the exact scan fell back to its sequential form"
    c.compute_energy()
    c.upload()
    sr, smt, phi, n = pot["star_r"], pot["star_m"], pot["phi"], N
    ids = np.arange(1, N + 1)
    b_of = binind.copy() if nb else None        # original binary number of each star, carried through the oracle's permutation
    for step in range(steps):
        E, J = cm.energies(phi[1:n + 1], r, vr, vt)
        if step:                                  # ComputeEnergy on the device == the host expression
            c.dev_compute_energy()
        want = _oracle_step(sr, smt, phi, n, N, E, J, vr, vt, pot["Rtidal"], seed=100 + step)
        res = c.dev_new_orbits(draws=want["draws"], draw_off=want["draw_off"], want_status=True)
        status, kmin, kmax = c.dev_read_status()
        assert int(res.n_draws) == len(want["draws"]) and int(res.n_deferred) < ORBIT_DEFER_CAP, (int(res.n_deferred), ORBIT_DEFER_CAP)
        assert np.array_equal(status, want["status"]), f"{name} step {step}: status"
        ok = want["status"] == ol.ST_OK
        assert np.array_equal(kmin[ok], want["kmin"][ok]) and np.array_equal(kmax[ok], want["kmax"][ok]), f"{name} step {step}: k_min / k_max"
        c.dev_transfer_and_sort()
        c.dev_calc_potential_new()
        assert c.potential_stats()[1] == 0, "the exact scan fell back to its sequential form"
        c.download()
        order = np.argsort(want["rnew"], kind="stable")
        n_new = int((want["rnew"] < cm.SF_INFINITY).sum())
        order = order[:n_new]
        assert c.clus.N_MAX == n_new and n_new >= n - n // 200
        assert np.array_equal(c.field("id")[1:n_new + 1], ids[order]), f"{name} step {step}: sort permutation"
        okk = ok[order]
        for f, w in (("r", want["rnew"]), ("vr", want["vrnew"]), ("vt", want["vtnew"]), ("m", smt[1:n + 1]), ("rOld", sr[1:n + 1]), ("E", E), ("J", J)):
            cm.assert_bits_equal(c.field(f)[1:n_new + 1], w[order], f"{name} step {step}: star[].{f}")
        for f in ("r_peri", "r_apo", "X"):
            cm.assert_bits_equal(c.field(f)[1:n_new + 1][okk], want[f][order][okk], f"{name} step {step}: star[].{f}")
        pot = cm.oracle_potential(want["rnew"][order], smt[1:n + 1][order], N_STAR=N)
        for f, w in (("star_r", pot["star_r"]), ("star_m", pot["star_m"]), ("star_phi", pot["phi"])):
            cm.assert_bits_equal(getattr(c, f)[:n_new + 2], w[:n_new + 2], f"{name} step {step}: {f}")
        assert (c.Mtotal, c.Rtidal) == (pot["Mtotal"], pot["Rtidal"])
        if nb:
            b_of = b_of[order]
            has = b_of > 0
            assert c.N_b_local == int(has.sum())
            assert np.array_equal(c.field("binind")[1:n_new + 1][has], np.arange(1, has.sum() + 1)) and not c.field("binind")[1:n_new + 1][~has].any()
            words = c.binary[1:has.sum() + 1].view(np.int64).reshape(-1, 58)
            assert np.array_equal(words, b_of[has][:, None] * 64 + np.arange(58)[None, :]), f"{name}: binaries did not follow their stars"
        # the next step starts from the oracle's state
        r, vr, vt = want["rnew"][order], want["vrnew"][order], want["vtnew"][order]
        sr, smt, phi, n, ids = pot["star_r"], pot["star_m"], pot["phi"], n_new, ids[order]
    c.close()
