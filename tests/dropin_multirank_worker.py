"""Run by tests/test_dropin_shim.py in a subprocess: P forked ranks (mini-MPI) of the compiled reference do
qsorts_new -> post_sort_comm -> calc_potential_new on a sharded cluster with binaries and stripped stars, either UNMODIFIED (mode "ref",
CPU only) or with the four entry points re-bound to the CUDA shim (mode "shim": rank i drives GPU i, the records travel GPU to GPU),
and every rank dumps what it holds afterwards.  The test compares the two dumps byte for byte.  TEST-ONLY.

    dropin_multirank_worker.py MODE PROCS N OUTDIR
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np

import common as cm
import refharness as rh


def main(mode, procs, N, outdir):
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=33, masses="imf")
    rng = np.random.default_rng(5)
    key = r * np.exp(rng.normal(0.0, 0.6, N))               # shuffled radii: every rank sends stars to every other rank
    key[rng.choice(N, N // 200, replace=False)] = cm.SF_INFINITY   # stripped stars: trimmed by the sort
    R = rh.Ref(rh.REF_B200_SO) if mode == "shim" else rh.Ref()
    assert R.L.ref_spawn(procs, int(3.0 * (N / procs + 4096) * 424) + (1 << 22)) == 0
    rank = R.L.ref_rank()
    if mode == "shim":
        os.environ["CMCB200_DEVICE"] = str(rank)
    nb_dim = N // procs // 4 + 64
    R.init(N, n_star_dim=int(2.5 * (N // procs + 4096)), n_table_dim=N + 4096, n_bin_dim=nb_dim, min_chunk=40)
    S = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start())).copy()
    E = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End())).copy()
    lo, hi = int(S[rank]), int(E[rank])
    nl = hi - lo + 1
    sb, bb = R.star_bytes(), R.binary_bytes()
    sb[1:nl + 1] = np.random.default_rng(100 + rank).integers(0, 256, (nl, R.sizeof_star), dtype=np.uint8)   # every byte of the record must travel
    R.star_field("r")[1:nl + 1] = key[lo - 1:hi]
    R.star_field("m")[1:nl + 1] = sm[lo - 1:hi]
    R.star_field("id", np.int64)[1:nl + 1] = np.arange(lo, hi + 1)
    nbin = nl // 8                                            # every 8th local star carries a binary, in shuffled slots
    binind = np.zeros(nl, np.int64)
    who = np.random.default_rng(200 + rank).choice(nl, nbin, replace=False)
    binind[who] = np.random.default_rng(300 + rank).permutation(np.arange(1, nbin + 1))
    R.star_field("binind", np.int64)[1:nl + 1] = binind
    bb[1:nbin + 1] = np.random.default_rng(400 + rank).integers(0, 256, (nbin, R.sizeof_binary), dtype=np.uint8)
    R.set_scalars(N, nl, N_BINARY=nbin)
    R.L.ref_qsorts_new()
    R.L.ref_post_sort_comm()
    R.L.ref_calc_potential_new()
    sc = R.scalars()
    n2, nl2, nb2 = int(sc["N_MAX"]), int(sc["N_MAX_NEW"]), int(sc["N_b_local"])
    S2 = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start())).copy()
    E2 = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End())).copy()
    np.savez(os.path.join(outdir, f"{mode}_rank{rank}.npz"), star=sb[1:nl2 + 1].copy(), binary=bb[1:].copy(), star_r=R.star_r()[:n2 + 2].copy(),
             star_m=R.star_m()[:n2 + 2].copy(), star_phi=R.star_phi()[:n2 + 2].copy(), scal=np.array([n2, nl2, nb2, sc["Mtotal"], sc["Rtidal"]]),
             S=S2, E=E2)
    if mode == "shim":
        R.L.cmcb200_shim_shutdown()
    rc = R.L.ref_join()
    if rank == 0:
        print("WORKER OK" if rc == 0 else "WORKER FAIL", flush=True)
        sys.exit(0 if rc == 0 else 4)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
