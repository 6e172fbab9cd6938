"""The resident timestep on HOST arrays (cmcb200_resident_begin / _end): the records cross the host link once each way per timestep.

Checked against the four host-buffer entry points called one after the other -- which tests/test_gpu_parity.py and tests/test_dropin_shim.py
check against the oracle and the compiled reference -- byte for byte over the whole star array, tables and scalars; the orbit step runs
chunk by chunk behind the arriving upload here (N is large enough for several chunks), so this also pins that the chunked step equals the
whole one."""
import numpy as np
import pytest

import common as cm
from cmc_cosmic_b200.cluster import Cluster

pytestmark = pytest.mark.gpu


def _cluster(N, r, vr, vt, sm, pinned, binind=None, nb=0, fill_seed=5):
    c = Cluster(N, n_star_dim=N + 64, n_bin_dim=nb + 8, pinned=pinned)
    c.load(r, vr, vt, sm, binind=binind)
    # every byte of the records travels: fill what the path never looks at with noise
    rng = np.random.default_rng(fill_seed)
    c.star[1:N + 1, 208:] = rng.integers(0, 256, (N, 424 - 208), dtype=np.uint8)
    if nb:
        c.binary[1:nb + 1] = rng.integers(0, 256, (nb, 464), dtype=np.uint8)
    c.calc_potential_new()
    c.compute_energy()
    return c


def _four_calls(c, draws, draw_off, seed):
    n = c.clus.N_MAX_NEW
    c.new_orbits_calculate(draws=draws, draw_off=draw_off, seed=seed)
    # ComputeIntermediateEnergy's transfer (reference cmc_utils.c:380-390), host code of the caller between the calls
    c.field("rOld")[1:n + 1] = c.star_r[1:n + 1]
    c.field("m")[1:n + 1] = c.star_m[1:n + 1]
    for dst, src in (("r", "rnew"), ("vr", "vrnew"), ("vt", "vtnew")):
        c.field(dst)[1:n + 1] = c.field(src)[1:n + 1]
    c.qsorts_new()
    c.post_sort_comm()
    c.calc_potential_new()
    c.tcount += 1


def _same(a, b, what):
    n = a.clus.N_MAX
    assert (a.clus.N_MAX, a.clus.N_MAX_NEW, a.N_b_local) == (b.clus.N_MAX, b.clus.N_MAX_NEW, b.N_b_local), what
    assert np.array_equal(a.star[:n + 1], b.star[:n + 1]), f"{what}: star[]"
    assert np.array_equal(a.binary, b.binary), f"{what}: binary[]"
    for f in ("star_r", "star_m", "star_phi"):
        cm.assert_bits_equal(getattr(a, f)[:n + 2], getattr(b, f)[:n + 2], f"{what}: {f}")
    assert (a.Mtotal, a.Rtidal, a.max_r) == (b.Mtotal, b.Rtidal, b.max_r), what


@pytest.mark.parametrize("N,pinned,nb", [(1_300_000, True, 0), (700_000, False, 20000), (5000, True, 0)])
def test_host_step_equals_the_four_host_buffer_calls(N, pinned, nb):
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=N % 1000, masses="imf" if nb else "equal")
    binind = None
    if nb:
        binind = np.zeros(N, np.int64)
        rng = np.random.default_rng(3)
        binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    a, b = (_cluster(N, r, vr, vt, sm, pinned, binind, nb) for _ in range(2))
    pot = cm.oracle_potential(r, sm)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    want = cm.oracle_orbits(pot["star_r"], pot["star_m"], pot["phi"], N, N, E, J, vr, vt, pot["Rtidal"], seed=3)
    # step 1: recorded draws (the oracle's), everything uploaded
    _four_calls(a, want["draws"], want["draw_off"], 0)
    res = b.host_step(draws=want["draws"], draw_off=want["draw_off"])
    assert int(res.n_draws) == len(want["draws"])
    _same(a, b, "step 1")
    n = a.clus.N_MAX
    o = np.argsort(want["rnew"], kind="stable")[:n]
    cm.assert_bits_equal(b.field("r")[1:n + 1], want["rnew"][o], "sorted new radii vs the oracle")
    h2d, d2h = b.transfer_bytes(reset=True)
    nb_now = b.N_b_local
    assert h2d >= N * 424 + 3 * 8 * (N + 2) and d2h >= n * 424 + 3 * 8 * (n + 2) + nb_now * 464
    # step 2: production draws; the host has rewritten E and J only (ComputeEnergy) and says so: the clean words stay on the device
    for c in (a, b):
        c.compute_energy()
    _four_calls(a, None, None, 77)
    b.host_step(seed=77, dirty_words=Cluster.dirty_mask("E", "J"))
    _same(a, b, "step 2 (dirty words: E, J)")
    h2d, _ = b.transfer_bytes(reset=True)
    assert h2d < n * 16 + 3 * 8 * (n + 2) + nb_now * 464 + 4096, "only the dirty words and the tables were sent"
    # step 3: a dirty mask is ignored (everything sent) once the device copy is not the one resident_end left: here after an upload()
    for c in (a, b):
        c.compute_energy()
    b.upload()
    _four_calls(a, None, None, 78)
    b.host_step(seed=78, dirty_words=Cluster.dirty_mask("E"))
    _same(a, b, "step 3 (stale device copy: full upload)")
    a.close(); b.close()


def test_resident_calls_order_themselves_behind_the_upload():
    """Every dev_* entry point may be the first call after resident_begin: it waits (on the device) for the whole upload.  Here the energy
    scheme's first pass and the sort are called straight after resident_begin, with the host-side upload() form as the reference."""
    N = 1_200_000
    r, vr, vt, sm = cm.cluster_state("king", N, seed=8)
    a, b = (_cluster(N, r, vr, vt, sm, True) for _ in range(2))
    a.upload()
    a.dev_energy_conservation1()
    a.dev_qsorts_new()
    a.dev_calc_potential_new()
    a.download()
    b.resident_begin()
    b.dev_energy_conservation1()
    b.dev_qsorts_new()
    b.dev_calc_potential_new()
    b.resident_end()
    _same(a, b, "ec1 + sort + potential")
    a.close(); b.close()


def test_restart_file_written_from_the_resident_arrays(tmp_path):
    """save_restart_file (reference src/cmc/cmc_io.c:3829-3884): the file is curr_st | restart_struct | clus | star[N_STAR_DIM_OPT] |
    binary[N_BIN_DIM_OPT] | snapshot_window_counters (:3862-3867).  Written while the records are resident and the host copy is stale, it
    must equal, byte for byte, the same layout assembled from the arrays after a download; star[0], the tails and the header blocks are the
    host's."""
    N, nb = 150000, 3000          # 64 MB of records: more than one 32 MB staging chunk
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=12, masses="imf")
    rng = np.random.default_rng(2)
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    c = Cluster(N, n_star_dim=N + 300, n_bin_dim=nb + 50)
    c.load(r, vr, vt, sm, binind=binind)
    c.binary[1:nb + 1] = rng.integers(0, 256, (nb, c.binary.shape[1]), dtype=np.uint8)
    c.star[N + 2:] = rng.integers(0, 256, c.star[N + 2:].shape, dtype=np.uint8)          # the unused tails travel too
    c.binary[nb + 1:] = rng.integers(0, 256, c.binary[nb + 1:].shape, dtype=np.uint8)
    c.star[0] = rng.integers(0, 256, c.star.shape[1], dtype=np.uint8)
    c.calc_potential_new()
    c.compute_energy()
    c.upload()
    c.dev_step(seed=3)                          # the device moves on; the host's records are now stale
    stale = c.star.copy()
    blobs = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (32, 1208, 56)]
    counters = np.arange(7, dtype=np.int32)
    path = tmp_path / "run.restart.1-0.bin"
    c.save_restart_file(path, *blobs, snapshot_window_counters=counters)
    got = np.fromfile(path, np.uint8)
    n_live, nb_live = c.clus.N_MAX_NEW, c.N_b_local
    tail_star, tail_bin, star0, bin0 = c.star[n_live + 1:].copy(), c.binary[nb_live + 1:].copy(), c.star[0].copy(), c.binary[0].copy()
    c.download()
    assert not np.array_equal(stale[1:n_live + 1], c.star[1:n_live + 1])
    want = np.concatenate([np.frombuffer(b, np.uint8) for b in blobs] + [star0, c.star[1:n_live + 1].reshape(-1), tail_star.reshape(-1), bin0,
                          c.binary[1:nb_live + 1].reshape(-1), tail_bin.reshape(-1), counters.view(np.uint8)])
    assert got.size == want.size == 32 + 1208 + 56 + c.N_STAR_DIM * 424 + c.N_BIN_DIM * 464 + 28
    assert np.array_equal(got, want)
    c.close()


@pytest.mark.parametrize("bh_only", [False, True])
def test_snapshot_rows_from_the_resident_arrays(bh_only):
    """SURVEY 8f N4, the HDF5 snapshot table: the rows write_snapshot() builds (reference src/cmc/cmc_io.c:3176-3327, the `Snapshot` struct)
    from the RESIDENT star[], binary[] and tables -- after a step of the resident pipeline, so the host's records are stale -- against the host
    build of the same core on the downloaded arrays (which tests/test_oracle_vs_ref.py::test_write_snapshot_rows pins against the compiled
    reference, bit for bit): every field of every row, all stars and bh_only, sigma given by the caller and taken from the resident
    calc_sigma_r; more rows than one device chunk holds."""
    import hostemul as he
    from cmc_cosmic_b200 import layout
    N, nb = 300000, 20000          # 146 MB of rows: two chunks of 2^18 rows
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=19, masses="imf")
    rng = np.random.default_rng(4)
    binind = np.zeros(N, np.int64)
    binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    c = Cluster(N, n_star_dim=N + 64, n_bin_dim=nb + 8)
    c.star[1:N + 1] = rng.uniform(-50.0, 50.0, (N, c.star.shape[1] // 8)).view(np.uint8).reshape(N, c.star.shape[1])
    c.binary[1:nb + 1] = rng.uniform(-50.0, 50.0, (nb, c.binary.shape[1] // 8)).view(np.uint8).reshape(nb, c.binary.shape[1])
    c.field("rnew")[:] = 0.0
    c.load(r, vr, vt, sm, binind=binind, ids=rng.permutation(np.arange(10, N + 10)))
    c.field("rad")[1:N + 1] = rng.uniform(1e-9, 1e-7, N)
    c.field("se_k")[1:N + 1] = rng.integers(0, 16, N)
    for k, lo, hi in (("m1", 0.1, 5.0), ("m2", 0.1, 5.0), ("a", 1e-6, 1e-3), ("e", 0.0, 0.99), ("rad1", 1e-9, 1e-7), ("rad2", 1e-9, 1e-7)):
        c.bfield(k)[1:nb + 1] = rng.uniform(lo, hi, nb)
    kw = np.lib.stride_tricks.as_strided(c.binary.reshape(-1)[84:92].view(np.int32), shape=(c.binary.shape[0], 2), strides=(c.binary.shape[1], 4))   # bse_kw[2]
    kw[1:nb + 1] = rng.integers(0, 16, (nb, 2))
    c.calc_potential_new()
    c.compute_energy()
    c.upload()
    c.dev_step(seed=9)                                      # the device moves on: the records are sorted by their new radii
    n = int(c.clus.N_MAX_NEW)
    units_m, units_l = 1.989e33 * 4.0e5, 3.0857e18 * 1.7
    sigma = np.concatenate(([0.0], rng.uniform(0.2, 0.9, n)))
    got = c.dev_snapshot_rows(units_m, units_l, sigma=sigma, avekernel=20, bh_only=bh_only)
    _, sg_res = c.dev_calc_sigma_r(20)
    got_res = c.dev_snapshot_rows(units_m, units_l, sigma=None, avekernel=20, bh_only=bh_only)   # eta from the resident sigma array
    c.download()
    nan_bits = 0xFFF8000000000000
    for rows, sg in ((got, sigma), (got_res, sg_res)):
        want = he.snapshot_rows(c.star, c.binary, c.star_r[:n + 2].copy(), c.star_m[:n + 2].copy(), c.star_phi[:n + 2].copy(), sg, n, n, 1, c.clus.N_MAX,
                                c.clus.N_STAR, units_m, units_l, 20, nan_bits, bh_only=bh_only)
        assert len(rows) == len(want) and (len(rows) == n if not bh_only else 0 < len(rows) < n // 2)
        for name in layout.SNAPSHOT_FIELDS:
            if rows[name].dtype == np.float64:
                cm.assert_bits_equal(rows[name], want[name], name)
            else:
                assert np.array_equal(rows[name], want[name]), name
    assert (got["binflag"] == 1).sum() == (c.field("binind")[1:n + 1] > 0).sum() if not bh_only else True
    c.close()
