"""BASELINE.json's full-size configurations on the GPU, checked through size-independent properties (the oracle needs minutes at these
sizes; bit-exactness is established at the sizes of test_gpu_parity.py / test_golden.py):

  configs[3]  Plummer sphere N = 1e7                         -- the bench workload
  configs[4]  King W0 = 7, N = 2e6, IMF masses, 10 % binaries

after two steps of the resident pipeline: the star array is a permutation of the input (id checksum), sorted by radius, consistent with
the replicated tables; binaries followed their stars and are renumbered 1, 2, ..; every sampled radius lies inside its orbit's
[r_peri, r_apo]; the potential satisfies its recurrence to rounding; Lagrange radii are ordered and (Plummer) at their analytic values.
"""
import numpy as np
import pytest

import common as cm
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster
from cmc_cosmic_b200.layout import binary_field

pytestmark = pytest.mark.gpu


def _check_state(c, N, ids0, steps):
    n = c.clus.N_MAX
    r, m, ids = c.field("r")[1:n + 1], c.field("m")[1:n + 1], c.field("id")[1:n + 1]
    assert n <= N and n >= N - N // 100                              # only a few stars leave the cluster per step
    assert np.all(np.diff(r) >= 0), "star[] is not sorted by radius"
    assert np.array_equal(c.star_r[1:n + 1], r) and np.array_equal(c.star_m[1:n + 1], m), "tables and records disagree"
    assert c.star_r[0] == cm.ZERO and c.star_r[n + 1] == cm.SF_INFINITY and c.star_phi[n + 1] == 0.0
    assert np.unique(ids).size == n and np.isin(ids, ids0).all(), "records lost or duplicated"
    # potential_calculate's recurrence (cmc_utils.c:716-720), evaluated here in one vectorised pass: agreement to the accumulated rounding
    mcum = np.cumsum(c.star_m[1:n + 1][::-1] / N)[::-1]                # mass outside k (inclusive) -> mprev_k = Mtotal - (mass outside k) + m_k/N
    mprev = c.Mtotal - (mcum - c.star_m[1:n + 1] / N)
    inv = 1.0 / c.star_r[1:n + 2]
    dphi = mprev * (inv[:-1] - inv[1:])
    phi = -np.cumsum(dphi[::-1])[::-1]
    assert np.allclose(c.star_phi[1:n + 1], phi, rtol=1e-7, atol=0.0)
    assert np.all(np.diff(c.star_phi[1:n + 2]) >= -1e-9 * abs(c.star_phi[1]))
    # the orbit step's outputs of the last step travelled with the records: rnew became r, and lies inside the orbit
    peri, apo, X = c.field("r_peri")[1:n + 1], c.field("r_apo")[1:n + 1], c.field("X")[1:n + 1]
    sampled = apo > 0
    assert sampled.mean() > 0.99
    assert np.all(r[sampled] >= peri[sampled] * (1 - 1e-12)) and np.all(r[sampled] <= apo[sampled] * (1 + 1e-12))
    assert np.all((X[sampled] >= 0) & (X[sampled] <= 1))
    assert np.array_equal(c.field("rOld")[1:n + 1] > 0, np.ones(n, bool))
    return n


def test_plummer_n1e7_two_resident_steps():
    N = 10_000_000
    r, vr, vt, m = models.plummer(N, seed=1234)
    c = Cluster(N, n_star_dim=N + 4096)
    c.load(r, vr, vt, np.ones(N))
    ids0 = np.arange(1, N + 1)
    c.calc_potential_new()
    c.upload()
    for s in range(2):
        res = c.dev_step(seed=11)
        assert res.n_brent == 0 and abs((res.n_draws / max(res.n_status[0], 1) - 1) / 2 - 1.91) < 0.05   # mean rejection tries per star
    pcs = np.array([0.01, 0.1, 0.25, 0.5, 0.75, 0.9])
    c.dev_comp_mass_percent(pcs)
    rh = c.dev_clusdyn_calculate()
    c.dev_compute_energy()
    c.download()
    n = _check_state(c, N, ids0, 2)
    # Plummer sphere in Henon units (scale 3 pi / 16): r(M) = a / sqrt(M^(-2/3) - 1); the Monte Carlo steps keep the profile stationary
    a = 3 * np.pi / 16
    want = a / np.sqrt(pcs ** (-2.0 / 3.0) - 1.0)
    assert np.all(np.diff(c.mass_r) > 0) and np.allclose(c.mass_r, want, rtol=0.02), (c.mass_r, want)
    assert abs(rh - a / np.sqrt(0.5 ** (-2.0 / 3.0) - 1.0)) < 0.02
    assert np.all(np.abs(c.no_star_r - pcs * N) <= 2)            # equal masses: the count at a mass fraction is that fraction of N (+- the rounding of the running sum)
    E = c.field("E")[1:n + 1]
    assert np.all(E < 0) and abs(E.mean() + 0.75) < 0.01       # bound stars; <Phi + v^2/2> = 2 W + T = -1 + 1/4 for a virialised cluster in Henon units
    c.close()


def test_king_w7_n2e6_binaries_two_resident_steps():
    N, nb = 2_000_000, 200_000
    r, vr, vt, m = models.king(N, w0=7.0, seed=12345)
    sm = models.kroupa_like_masses(N, seed=3) * N
    rng = np.random.default_rng(9)
    binind = np.zeros(N, np.int64)
    who = rng.choice(N, nb, replace=False)
    binind[who] = rng.permutation(np.arange(1, nb + 1))
    c = Cluster(N, n_star_dim=N + 4096, n_bin_dim=nb + 64)
    c.load(r, vr, vt, sm, binind=binind)
    # every binary record carries the id of its star in its first word, the rest is a recognisable pattern
    bid = np.zeros(nb + 1, np.int64)
    bid[binind[who]] = who + 1
    binary_field(c.binary, "id1")[1:nb + 1] = bid[1:]
    binary_field(c.binary, "a")[1:nb + 1] = bid[1:] * 0.5
    c.calc_potential_new()
    c.upload()
    for s in range(2):
        c.dev_step(seed=5)
    c.download()
    n = _check_state(c, N, np.arange(1, N + 1), 2)
    b = c.field("binind")[1:n + 1]
    has = b > 0
    assert c.N_b_local == has.sum() and np.array_equal(b[has], np.arange(1, has.sum() + 1)), "binind not renumbered 1, 2, .. in star order"
    ids = c.field("id")[1:n + 1]
    assert np.array_equal(binary_field(c.binary, "id1")[b[has]], ids[has]), "a binary did not follow its star"
    assert np.array_equal(binary_field(c.binary, "a")[b[has]], ids[has] * 0.5)
    assert has.sum() >= nb - nb // 50
    assert not c.binary[c.N_b_local + 1:].any(), "binary[] beyond the packed records must be zero (cmc_sort.c:541)"
    c.close()
