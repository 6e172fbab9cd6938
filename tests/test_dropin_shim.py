"""The drop-in boundary, end to end: the reference's own compiled objects and harness, with new_orbits_calculate / qsorts_new /
post_sort_comm / calc_potential_new re-bound at link time (ld --wrap) to cmc_cosmic_b200/shim/cmcb200_shim.c and through it to the
CUDA library -- against the untouched reference build, on identical inputs and identical taus113 draws.  Everything around the four
calls (ComputeEnergy, remove_star / destroy_obj side effects, potential(), findLimits) is the reference's code in both runs, so
the whole star_t / binary_t arrays, the replicated tables and the escaper accumulators must come out byte-identical."""
import ctypes as C
import os

import numpy as np
import pytest

import common as cm
import oraclelib as ol
import refharness as rh

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (rh.available() and os.path.exists(rh.REF_B200_SO)), reason="oracle/_ref not built")]


def _run(R, N, r, vr, vt, sm, binind, nb, rtidal, draws=None, draw_off=None, is_shim=False):
    R.init(N, n_star_dim=N + 64, n_bin_dim=nb + 8, idum=99)
    R.load_cluster(r, vr, vt, sm, binind=binind)
    if nb:
        words = (np.arange(1, nb + 1, dtype=np.int64)[:, None] * 64 + np.arange(R.sizeof_binary // 8, dtype=np.int64)[None, :])
        R.binary_bytes()[1:nb + 1] = words.view(np.uint8).reshape(nb, R.sizeof_binary)
    R.set_scalars(N, N, N_BINARY=nb)
    R.L.ref_calc_potential_new()
    R.compute_energy()
    snap = {"phi0": R.star_phi()[:N + 2].copy(), "E": R.star_field("E")[1:N + 1].copy(), "J": R.star_field("J")[1:N + 1].copy()}
    sc = R.scalars()
    if rtidal:
        R.set_scalars(N, N, N_BINARY=nb, Rtidal=rtidal, Mtotal=sc["Mtotal"])
    snap["Rtidal"] = R.scalars()["Rtidal"]
    if is_shim:
        d = np.ascontiguousarray(draws); o = np.ascontiguousarray(draw_off, np.int64)
        R.L.cmcb200_shim_set_draws.argtypes = [C.c_void_p, C.c_void_p]
        R.L.cmcb200_shim_set_draws(d.ctypes.data, o.ctypes.data)
        snap["_keep"] = (d, o)
    else:
        R.L.ref_draw_mode(1)
    R.L.ref_new_orbits_calculate()
    if not is_shim:
        R.L.ref_draw_mode(0)
        snap["draws"] = R.draws()
    snap["star_after_orbits"] = R.star_bytes()[:N + 2].copy()
    snap["scalars_after_orbits"] = R.scalars()
    snap["star_m_after_orbits"] = R.star_m()[:N + 2].copy()
    srr, smm = R.star_r(), R.star_m()
    for dst, src in (("r", "rnew"), ("vr", "vrnew"), ("vt", "vtnew")):
        R.star_field(dst)[1:N + 1] = R.star_field(src)[1:N + 1]
    R.star_field("rOld")[1:N + 1] = srr[1:N + 1]
    R.star_field("m")[1:N + 1] = smm[1:N + 1]
    R.L.ref_qsorts_new()
    R.L.ref_post_sort_comm()
    R.L.ref_calc_potential_new()
    R.compute_energy()
    n2 = int(R.scalars()["N_MAX"])
    snap["n2"] = n2
    snap["star_final"] = R.star_bytes()[:n2 + 1].copy()
    snap["binary_final"] = R.binary_bytes().copy()
    snap["tables_final"] = (R.star_r()[:n2 + 2].copy(), R.star_m()[:n2 + 2].copy(), R.star_phi()[:n2 + 2].copy())
    snap["scalars_final"] = R.scalars()
    S = np.ctypeslib.as_array((C.c_int * 1).from_address(R.L.ref_Start())).copy()
    E = np.ctypeslib.as_array((C.c_int * 1).from_address(R.L.ref_End())).copy()
    snap["limits"] = (int(S[0]), int(E[0]))
    if is_shim:
        R.L.cmcb200_shim_shutdown()
    R.free()
    return snap


@pytest.mark.parametrize("kind,N,masses,nb,rtidal", [("plummer", 30000, "equal", 0, 0.0), ("plummer", 20000, "imf", 2000, 4.0), ("king", 25000, "equal", 500, 0.0)])
def test_reference_with_cuda_shim_equals_reference(kind, N, masses, nb, rtidal):
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=N + 7, masses=masses)
    rng = np.random.default_rng(N)
    binind = np.zeros(N, np.int64)
    if nb:
        binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    a = _run(rh.Ref(), N, r, vr, vt, sm, binind, nb, rtidal)
    # CSR offsets of the reference's sequential draw stream (the oracle replays the same loop; pinned by tests/test_golden.py)
    srt, smt = cm.make_tables(r, sm)
    t = ol.tables(srt, smt, a["phi0"].copy(), N, N)
    off = ol.get_positions(t, np.arange(1, N + 1), a["E"], a["J"], vr, vt, a["Rtidal"], a["draws"])["draw_off"]
    assert off[-1] == len(a["draws"])
    b = _run(rh.Ref(rh.REF_B200_SO), N, r, vr, vt, sm, binind, nb, rtidal, draws=a["draws"], draw_off=off, is_shim=True)
    cm.assert_bits_equal(b["phi0"], a["phi0"], "initial potential")
    cm.assert_bits_equal(b["E"], a["E"], "E")
    assert np.array_equal(b["star_after_orbits"], a["star_after_orbits"]), "star[] after new_orbits_calculate"
    assert np.array_equal(b["star_m_after_orbits"].view(np.uint64), a["star_m_after_orbits"].view(np.uint64))
    assert b["scalars_after_orbits"] == a["scalars_after_orbits"], (a["scalars_after_orbits"], b["scalars_after_orbits"])
    if rtidal:
        assert a["scalars_after_orbits"]["TidalMassLoss"] > 0 and a["scalars_after_orbits"]["esc_bytes"] > 0
    assert b["n2"] == a["n2"] and b["limits"] == a["limits"]
    assert np.array_equal(b["star_final"], a["star_final"]), "star[] after sort + potential + ComputeEnergy"
    assert np.array_equal(b["binary_final"], a["binary_final"]), "binary[]"
    for x, y, name in zip(b["tables_final"], a["tables_final"], ("star_r", "star_m", "star_phi")):
        cm.assert_bits_equal(x, y, name)
    assert b["scalars_final"] == a["scalars_final"]


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("procs,N", [(2, 60000), (2, 200003)])
def test_multirank_reference_with_cuda_shim_equals_multirank_reference(procs, N, tmp_path):
    """P forked MPI ranks of the compiled reference, once unmodified (sample_sort + load_balance + Allgatherv + potential on the CPU) and once
    with qsorts_new / post_sort_comm / calc_potential_new re-bound to the shim (rank i drives GPU i: radius-sharded sort, records stored
    GPU to GPU over NVLink, MPI only for two small collectives and the barriers): every rank must end with the same bytes -- its shard of
    star[], its binaries, the replicated tables, N_MAX, N_MAX_NEW, N_b_local, Start/End"""
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin_multirank_worker.py")
    for mode in ("ref", "shim"):
        p = subprocess.run([sys.executable, worker, mode, str(procs), str(N), str(tmp_path)], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and "WORKER OK" in p.stdout, (mode, p.stdout[-3000:], p.stderr[-3000:])
    for rank in range(procs):
        a, b = np.load(tmp_path / f"ref_rank{rank}.npz"), np.load(tmp_path / f"shim_rank{rank}.npz")
        assert np.array_equal(a["scal"], b["scal"]), (rank, a["scal"], b["scal"])
        assert np.array_equal(a["S"], b["S"]) and np.array_equal(a["E"], b["E"])
        assert np.array_equal(a["star"], b["star"]), f"rank {rank}: star[] shard differs"
        assert np.array_equal(a["binary"], b["binary"]), f"rank {rank}: binary[] differs"
        for f in ("star_r", "star_m", "star_phi"):
            assert np.array_equal(a[f].view(np.uint64), b[f].view(np.uint64)), f"rank {rank}: {f}"


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_multirank_orbit_step_through_the_shim_is_independent_of_the_rank_count(tmp_path):
    """new_orbits_calculate through the shim on P = 1 and P = 2 forked ranks (each with its GPU): with the production draws (Philox keyed by
    the global star index) every star must get the same new position whichever rank owns it; max_r and TidalMassLoss are reduced over the
    ranks like the reference does (cmc_utils.c:2206-2209)"""
    import subprocess
    import sys
    N = 150000
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin_multirank_orbit_worker.py")
    for procs in (1, 2):
        p = subprocess.run([sys.executable, worker, str(procs), str(N), str(tmp_path)], capture_output=True, text=True, timeout=900)
        assert p.returncode == 0 and "WORKER OK" in p.stdout, (procs, p.stdout[-3000:], p.stderr[-3000:])
    one = np.load(tmp_path / "orbit_p1_rank0.npz")
    parts = [np.load(tmp_path / f"orbit_p2_rank{q}.npz") for q in range(2)]
    assert int(parts[0]["lo"]) == 1 and int(parts[1]["hi"]) == N and int(parts[1]["lo"]) == int(parts[0]["hi"]) + 1
    for f in ("E", "J", "rnew", "vrnew", "vtnew", "X", "r_peri", "r_apo"):
        both = np.concatenate([q[f] for q in parts])
        assert np.array_equal(both.view(np.uint64), one[f].view(np.uint64)), f
    assert (one["rnew"] == cm.SF_INFINITY).sum() > 10                          # the tidal cut removed stars
    for q in parts:
        assert q["max_r"] == one["max_r"]
        assert abs(q["tml"] - one["tml"]) <= 1e-12 * abs(one["tml"])             # a sum over ranks: same value up to the order of the additions
