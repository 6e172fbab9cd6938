"""ctypes binding of tests/host_emul/_build/libemul.so: cmc_cosmic_b200/csrc/orbit_core.h (the per-star code the CUDA kernel runs)
compiled for the host with g++ -ffp-contract=off.  TEST-ONLY; the product never computes on the host."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
_SO = os.path.join(_DIR, "_build", "libemul.so")
_SO_DEEP = os.path.join(_DIR, "_build", "libemul_deep.so")  # tiny top levels: deep ladders on small tables
_lib = None
_deep = False


def use_deep_ladders(on):
    """switch to the build whose ladders have 6-block / 10-sample top levels"""
    global _lib, _deep
    if on != _deep:
        _lib, _deep = None, on


def build():
    subprocess.check_call(["make", "-s", "-C", _DIR])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO_DEEP if _deep else _SO)
        vp, lg, db = C.c_void_p, C.c_long, C.c_double
        L.emul_get_positions.restype = db
        L.emul_get_positions.argtypes = [vp, vp, vp, lg, lg, db, lg, vp, vp, vp, vp, vp, vp, vp, db, vp, lg, vp, C.c_int] + [vp] * 10
        L.emul_find_zero_Q.restype = lg
        L.emul_find_zero_Q.argtypes = [vp, vp, vp, lg, lg, lg, db, db, lg, lg, vp, C.c_int]
        L.emul_ladder_plan.restype = C.c_int
        L.emul_ladder_plan.argtypes = [lg] + [vp] * 6
        L.emul_find_zero_r_both.argtypes = [vp, lg, vp, lg, vp, vp]
        L.emul_set_velocities3.restype = db
        L.emul_set_velocities3.argtypes = [vp, vp, vp, lg, lg, lg, lg, lg] + [vp] * 12
        L.emul_set_velocities3_second.restype = db
        L.emul_set_velocities3_second.argtypes = [vp, lg, lg, lg, vp, db, vp, vp, vp, vp]
        L.emul_calc_sigma_r.restype = C.c_int
        L.emul_calc_sigma_r.argtypes = [lg, lg, lg, lg, vp, vp, db, vp, vp, C.c_int, C.c_int, vp, vp, vp]
        L.emul_sigma_scan.restype = C.c_int
        L.emul_sigma_scan.argtypes = [lg, lg, lg, lg, vp, vp, db, vp, vp, C.c_int, db, vp, vp, vp]
        L.emul_snapshot_rows.restype = lg
        L.emul_snapshot_rows.argtypes = [vp] * 6 + [lg] * 5 + [db, db, db, lg, C.c_uint64, C.c_int, vp]
        L.emul_dynamics_pairs.restype = lg
        L.emul_dynamics_pairs.argtypes = [vp, vp, vp, lg, lg, lg, lg, db] + [vp] * 17
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def get_positions(sr, sm, phi, N_MAX, N_STAR, cenma_m, g_idx, E, J, vr, vt, Rtidal, draws, csr_off=None, is_new=None, rnew_in=None, no_ladders=False, screened=False, sampled=False):
    """no_ladders: every probe on the fine FP64 table; screened: r ladder + FP32 screening copy, no Q ladder; sampled: the device's default configuration (r ladder + sample ladder for the index bisections)"""
    n = len(g_idx)
    g_idx = np.ascontiguousarray(g_idx, np.int64)
    arrs = [np.ascontiguousarray(x, np.float64) for x in (E, J, vr, vt)]
    draws = np.ascontiguousarray(draws, np.float64)
    if csr_off is not None:
        csr_off = np.ascontiguousarray(csr_off, np.int64)
    out = dict(status=np.zeros(n, np.int32), rnew=np.zeros(n), vrnew=np.zeros(n), vtnew=np.zeros(n), X=np.zeros(n), r_peri=np.zeros(n),
               r_apo=np.zeros(n), kmin=np.zeros(n, np.int64), kmax=np.zeros(n, np.int64))
    cnt = np.zeros(5, np.int64)
    out["max_rad"] = lib().emul_get_positions(_p(sr), _p(sm), _p(phi), N_MAX, N_STAR, cenma_m, n, _p(g_idx), _p(is_new), _p(arrs[0]), _p(arrs[1]),
                                              _p(arrs[2]), _p(arrs[3]), _p(rnew_in), Rtidal, _p(draws), len(draws), _p(csr_off), 1 if no_ladders else (2 if screened else (3 if sampled else 0)),
                                              _p(out["status"]), _p(out["rnew"]), _p(out["vrnew"]), _p(out["vtnew"]), _p(out["X"]),
                                              _p(out["r_peri"]), _p(out["r_apo"]), _p(out["kmin"]), _p(out["kmax"]), _p(cnt))
    out["n_draws"], out["n_exact"], out["n_brent"] = int(cnt[0]), int(cnt[1]), int(cnt[2])
    out["q_levels"], out["r_levels"] = int(cnt[3]), int(cnt[4])
    return out


def find_zero_r_both(star_r, N_MAX, radii):
    """(FindZero_r by the reference's bisection, by the r ladder [-1: table too small for a ladder]) for each radius"""
    radii = np.ascontiguousarray(radii, np.float64)
    a, b = np.zeros(len(radii), np.int64), np.zeros(len(radii), np.int64)
    lib().emul_find_zero_r_both(_p(np.ascontiguousarray(star_r)), N_MAX, _p(radii), len(radii), _p(a), _p(b))
    return a, b


def find_zero_Q(sr, sm, phi, N_MAX, N_STAR, g_j, E, J, kmin, kmax, use_ladder=True):
    ne = np.zeros(1, np.int64)
    k = lib().emul_find_zero_Q(_p(sr), _p(sm), _p(phi), N_MAX, N_STAR, g_j, E, J, kmin, kmax, _p(ne), 1 if use_ladder else 0)
    return int(k), int(ne[0])


def ladder_plan(M):
    qs, rs = np.zeros(8, np.int32), np.zeros(8, np.int32)
    qc, rc, ra = np.zeros(8, np.int64), np.zeros(8, np.int64), np.zeros(8, np.int64)
    nr = np.zeros(1, np.int32)
    nq = lib().emul_ladder_plan(M, _p(qs), _p(qc), _p(rs), _p(rc), _p(ra), _p(nr))
    return dict(q_shift=qs[:nq], q_count=qc[:nq], r_shift=rs[:nr[0]], r_count=rc[:nr[0]], r_alloc=ra[:nr[0]])


def set_velocities3(sr, sm, phi, N_MAX, N_STAR, first_global, f, blk=1024):
    """first loop of set_velocities3 as the CUDA path decomposes it (velocities_core.h); f: the fields of cm.sv3_fields for the local stars.
    Returns (vr, vt, vrnew, vtnew after, kind bytes, leftover Eexcess, counters)"""
    n = len(f["vr"])
    inter = np.ascontiguousarray(f["interacted"], np.int64)
    ins = [np.ascontiguousarray(f[k], np.float64) for k in ("rOld", "vrold", "vtold", "Uoldrold", "Uoldrnew")]
    io = {k: np.array(f[k], dtype=np.float64, copy=True) for k in ("vr", "vt", "vrnew", "vtnew")}
    kind, cnt = np.zeros(n, np.uint8), np.zeros(3, np.int64)
    left = lib().emul_set_velocities3(_p(sr), _p(sm), _p(phi), N_MAX, N_STAR, n, first_global, blk, _p(inter), *[_p(a) for a in ins],
                                      *[_p(io[k]) for k in ("vr", "vt", "vrnew", "vtnew")], _p(kind), _p(cnt))
    return io, kind, left, dict(replayed=int(cnt[0]), walked=int(cnt[1]), failed=int(cnt[2]))


def set_velocities3_second(sm, N_MAX, first_global, kind, Eexcess, io):
    return lib().emul_set_velocities3_second(_p(sm), N_MAX, len(kind), first_global, _p(kind), Eexcess, _p(io["vr"]), _p(io["vt"]), _p(io["vrnew"]), _p(io["vtnew"]))


def calc_sigma_r(p, star_r, star_m, N_MAX, N_STAR, vr_all, vt_all, first_global=1, n_local=None, mave=False, exact_scan=False):
    """calc_sigma_r through the operation list of sigma_core.h (exact_scan: through the exact-scan decomposition the device path will use).
    vr_all / vt_all: the whole sorted cluster; the ghost stars are cut from them.  Returns (r or mean mass, sigma, events)"""
    vr_all, vt_all = np.ascontiguousarray(vr_all, np.float64), np.ascontiguousarray(vt_all, np.float64)
    lo = first_global - 1
    n = len(vr_all) - lo if n_local is None else n_local
    def cut(x):
        out = np.zeros(n + 2 * p)
        a, b = lo - p, lo + n + p
        src = x[max(a, 0):min(b, len(x))]
        out[max(a, 0) - a:max(a, 0) - a + len(src)] = src
        return out
    vr, vt = cut(vr_all), cut(vt_all)
    a, sg, cnt = np.zeros(n + 1), np.zeros(n + 1), np.zeros(1, np.int64)
    rc = lib().emul_calc_sigma_r(p, n, first_global, N_MAX, _p(np.ascontiguousarray(star_r)), _p(np.ascontiguousarray(star_m)), 1.0 / N_STAR, _p(vr), _p(vt),
                                 int(mave), int(exact_scan), _p(a), _p(sg), _p(cnt))
    assert rc == 0
    return a, sg, int(cnt[0])


def dynamics_pairs(sr, sm, phi, N_MAX, N_STAR, first_global, dt, cfg, draws, f, bin_a):
    """the per-pair code of dynamics.cu (csrc/dynamics_core.h) in a host loop; f as oraclelib.dynamics_apply_pairs takes it (modified in place)"""
    n = len(f["vr"])
    sw = np.array([cfg[k] for k in ("relaxation", "ss_collision", "binsingle", "binbin", "tidal_capture", "bh_capture", "p")], np.int64)
    par = np.array([cfg[k] for k in ("gamma", "coll_factor", "tc_factor", "wd_tc", "units_mstar", "units_l", "units_t")], np.float64)
    out = dict(status=np.zeros(n // 2, np.int32), rperi=np.zeros(n // 2), W=np.zeros(n // 2), rcm=np.zeros(n // 2), beta=np.zeros(n // 2))
    lib().emul_dynamics_pairs(_p(sr), _p(sm), _p(phi), N_MAX, N_STAR, n, first_global, dt, _p(sw), _p(par), _p(np.ascontiguousarray(draws)), _p(f["vr"]),
                              _p(f["vt"]), _p(f["E"]), _p(f["J"]), _p(f["Y"]), _p(f["binind"]), _p(f["se_k"]), _p(f["rad"]), _p(bin_a),
                              _p(out["status"]), _p(out["rperi"]), _p(out["W"]), _p(out["rcm"]), _p(out["beta"]))
    return out


def sigma_scan(p, star_r, star_m, N_MAX, N_STAR, vr_all, vt_all, first_global=1, n_local=None, mave=False, band_scale=1.0):
    """calc_sigma_r through the segment / tile tables of csrc/sigma_scan.h, the four passes of sigma.cu as host loops.
    Returns (r or mean mass, sigma, counters dict); None when the window does not fit the shard."""
    vr_all, vt_all = np.ascontiguousarray(vr_all, np.float64), np.ascontiguousarray(vt_all, np.float64)
    lo = first_global - 1
    n = len(vr_all) - lo if n_local is None else n_local
    def cut(x):
        out = np.zeros(n + 2 * p)
        a, b = lo - p, lo + n + p
        src = x[max(a, 0):min(b, len(x))]
        out[max(a, 0) - a:max(a, 0) - a + len(src)] = src
        return out
    vr, vt = cut(vr_all), cut(vt_all)
    a, sg, cnt = np.zeros(n + 1), np.zeros(n + 1), np.zeros(6, np.int64)
    rc = lib().emul_sigma_scan(p, n, first_global, N_MAX, _p(np.ascontiguousarray(star_r)), _p(np.ascontiguousarray(star_m)), 1.0 / N_STAR, _p(vr), _p(vt),
                               int(mave), band_scale, _p(a), _p(sg), _p(cnt))
    if rc != 0:
        return None
    return a, sg, dict(zip(("hard_segments", "wide_segments", "hard_tiles", "broken", "inconsistent", "tiles"), cnt.tolist()))


def snapshot_rows(star_bytes, binary_bytes, star_r, star_m, star_phi, sigma, n_local, n_old_local, first_global, N_MAX, N_STAR, units_m, units_l,
                  avekernel, nan_bits, bh_only=False):
    """write_snapshot's rows through the host build of csrc/snapshot_core.h.  star_bytes / binary_bytes: the (n, 424) / (n, 464) uint8 arrays
    with record 0 in front; sigma: [n_local + 1] or None.  Returns a structured array (layout.SNAPSHOT_DTYPE)."""
    from cmc_cosmic_b200 import layout
    sb, bb = np.ascontiguousarray(star_bytes), np.ascontiguousarray(binary_bytes)
    rows = np.zeros(max(n_local, 1), layout.SNAPSHOT_DTYPE)
    sg = None if sigma is None else np.ascontiguousarray(sigma, np.float64)
    k = lib().emul_snapshot_rows(sb.ctypes.data + layout.STAR_DTYPE_SIZE, bb.ctypes.data, _p(np.ascontiguousarray(star_r)), _p(np.ascontiguousarray(star_m)),
                                 _p(np.ascontiguousarray(star_phi)), None if sg is None else _p(sg), n_local, n_old_local, first_global, N_MAX, N_STAR,
                                 units_m, units_l, 1.0 / N_STAR, avekernel, int(nan_bits), int(bh_only), rows.ctypes.data)
    return rows[:k]
