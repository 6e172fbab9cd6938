"""ctypes binding of tests/host_emul/_build/libemul.so: cmc_cosmic_b200/csrc/orbit_core.h (the per-star code the CUDA kernel runs)
compiled for the host with g++ -ffp-contract=off.  TEST-ONLY; the product never computes on the host."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
_SO = os.path.join(_DIR, "_build", "libemul.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, lg, db = C.c_void_p, C.c_long, C.c_double
        L.emul_get_positions.restype = db
        L.emul_get_positions.argtypes = [vp, vp, vp, lg, lg, db, lg, vp, vp, vp, vp, vp, vp, vp, db, vp, lg, vp, C.c_int] + [vp] * 10
        L.emul_find_zero_Q.restype = lg
        L.emul_find_zero_Q.argtypes = [vp, vp, vp, lg, lg, lg, db, db, lg, lg, vp]
        L.emul_find_zero_r_both.argtypes = [vp, lg, vp, lg, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def get_positions(sr, sm, phi, N_MAX, N_STAR, cenma_m, g_idx, E, J, vr, vt, Rtidal, draws, csr_off=None, is_new=None, rnew_in=None):
    n = len(g_idx)
    g_idx = np.ascontiguousarray(g_idx, np.int64)
    arrs = [np.ascontiguousarray(x, np.float64) for x in (E, J, vr, vt)]
    draws = np.ascontiguousarray(draws, np.float64)
    if csr_off is not None:
        csr_off = np.ascontiguousarray(csr_off, np.int64)
    out = dict(status=np.zeros(n, np.int32), rnew=np.zeros(n), vrnew=np.zeros(n), vtnew=np.zeros(n), X=np.zeros(n), r_peri=np.zeros(n),
               r_apo=np.zeros(n), kmin=np.zeros(n, np.int64), kmax=np.zeros(n, np.int64))
    cnt = np.zeros(3, np.int64)
    out["max_rad"] = lib().emul_get_positions(_p(sr), _p(sm), _p(phi), N_MAX, N_STAR, cenma_m, n, _p(g_idx), _p(is_new), _p(arrs[0]), _p(arrs[1]),
                                              _p(arrs[2]), _p(arrs[3]), _p(rnew_in), Rtidal, _p(draws), len(draws), _p(csr_off), 0,
                                              _p(out["status"]), _p(out["rnew"]), _p(out["vrnew"]), _p(out["vtnew"]), _p(out["X"]),
                                              _p(out["r_peri"]), _p(out["r_apo"]), _p(out["kmin"]), _p(out["kmax"]), _p(cnt))
    out["n_draws"], out["n_exact"], out["n_brent"] = int(cnt[0]), int(cnt[1]), int(cnt[2])
    return out


def find_zero_r_both(star_r, N_MAX, radii):
    """(FindZero_r by the reference's bisection, by the 16-ary index) for each radius"""
    radii = np.ascontiguousarray(radii, np.float64)
    a, b = np.zeros(len(radii), np.int64), np.zeros(len(radii), np.int64)
    lib().emul_find_zero_r_both(_p(np.ascontiguousarray(star_r)), N_MAX, _p(radii), len(radii), _p(a), _p(b))
    return a, b
