"""Run by tests/test_dropin_shim.py in a subprocess: P forked ranks of the compiled reference with the CUDA shim do
calc_potential_new -> ComputeEnergy -> new_orbits_calculate (production draws: Philox keyed by the GLOBAL star index, so the result does
not depend on P) and dump the new positions of their shard.  The test compares P = 1 with P = 2.  TEST-ONLY.

    dropin_multirank_orbit_worker.py PROCS N OUTDIR
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np

import common as cm
import refharness as rh


def main(procs, N, outdir):
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=44)
    R = rh.Ref(rh.REF_B200_SO)
    assert R.L.ref_spawn(procs, int(3.0 * (N / procs + 4096) * 424) + (1 << 22)) == 0
    rank = R.L.ref_rank()
    os.environ["CMCB200_DEVICE"] = str(rank)
    R.init(N, n_star_dim=int(2.5 * (N // procs + 4096)), n_table_dim=N + 4096, min_chunk=40, idum=977)
    S = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_Start())).copy()
    E = np.ctypeslib.as_array((C.c_int * procs).from_address(R.L.ref_End())).copy()
    lo, hi = int(S[rank]), int(E[rank])
    nl = hi - lo + 1
    sr, smm = R.star_r(), R.star_m()
    sr[0] = 1e-20; sr[1:N + 1] = r; sr[N + 1] = 1e10; smm[1:N + 1] = sm          # the replicated tables, as after post_sort_comm
    for name, val in (("r", r), ("vr", vr), ("vt", vt), ("m", sm)):
        R.star_field(name)[1:nl + 1] = val[lo - 1:hi]
    R.star_field("id", np.int64)[1:nl + 1] = np.arange(lo, hi + 1)
    R.set_scalars(N, nl)
    R.L.ref_calc_potential_new()
    R.L.ref_compute_energy()                                   # the reference's own ComputeEnergy on the shard (get_global_idx)
    R.set_scalars(N, nl, Rtidal=5.0, Mtotal=R.scalars()["Mtotal"])   # a tidal cut: removed stars exercise the shim's host-side replay
    R.L.ref_new_orbits_calculate()
    out = {f: R.star_field(f)[1:nl + 1].copy() for f in ("rnew", "vrnew", "vtnew", "X", "r_peri", "r_apo", "E", "J")}
    sc = R.scalars()
    np.savez(os.path.join(outdir, f"orbit_p{procs}_rank{rank}.npz"), lo=lo, hi=hi, max_r=sc["max_r"], tml=sc["TidalMassLoss"], **out)
    R.L.cmcb200_shim_shutdown()
    rc = R.L.ref_join()
    if rank == 0:
        print("WORKER OK" if rc == 0 else "WORKER FAIL", flush=True)
        sys.exit(0 if rc == 0 else 4)


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])
