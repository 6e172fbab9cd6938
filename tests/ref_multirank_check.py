"""Run by tests/test_host.py in a subprocess: P forked ranks of the compiled reference (mini-MPI shim) do one sample sort +
post_sort_comm + calc_potential_new; every rank checks its shard and the replicated tables against the oracle.  TEST-ONLY."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np

import common as cm
import oraclelib as ol
import refharness as rh


def main(procs, N):
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=21, masses="imf")
    rng = np.random.default_rng(4)
    rnew = r * np.exp(rng.normal(0.0, 0.5, N))            # shuffled radii: the sort has work to do
    R = rh.Ref()
    assert R.L.ref_spawn(procs, int(3.0 * (N / procs + 4096) * 424) + (1 << 22)) == 0
    rank = R.L.ref_rank()
    R.init(N, n_star_dim=int(2.5 * (N // procs + 4096)), n_table_dim=N + 4096, min_chunk=40)
    S = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start())).copy()
    E = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End())).copy()
    S0, E0 = ol.find_limits(N, 40, procs)
    ok = np.array_equal(S, S0) and np.array_equal(E, E0)
    lo, hi = int(S[rank]), int(E[rank])
    nl = hi - lo + 1
    R.star_field("r")[1:nl + 1] = rnew[lo - 1:hi]
    R.star_field("m")[1:nl + 1] = sm[lo - 1:hi]
    R.star_field("id", np.int64)[1:nl + 1] = np.arange(lo, hi + 1)
    R.set_scalars(N, nl)
    R.L.ref_qsorts_new()
    R.L.ref_post_sort_comm()
    R.L.ref_calc_potential_new()
    order = np.argsort(rnew, kind="stable")
    S = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start())).copy()
    E = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End())).copy()
    lo, hi = int(S[rank]), int(E[rank])
    ok = ok and int(R.scalars()["N_MAX_NEW"]) == hi - lo + 1
    ok = ok and np.array_equal(R.star_field("id", np.int64)[1:hi - lo + 2], order[lo - 1:hi] + 1)
    want = cm.oracle_potential(rnew[order], sm[order], N)
    ok = ok and np.array_equal(R.star_r()[1:N + 1], rnew[order]) and np.array_equal(R.star_phi()[:N + 2].view(np.uint64), want["phi"][:N + 2].view(np.uint64))
    if not ok:
        print(f"rank {rank}: MISMATCH", flush=True)
        sys.exit(3)
    rc = R.L.ref_join()
    if rank == 0:
        print("MULTIRANK OK" if rc == 0 else "MULTIRANK FAIL", flush=True)
        sys.exit(0 if rc == 0 else 4)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
