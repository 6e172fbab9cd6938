"""ctypes binding to oracle/libcmc_oracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(_HERE, "..", "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libcmc_oracle.so")

SF_INFINITY = 1.0e10
ZERO = 1.0e-20
ST_OK, ST_MASSLESS, ST_UNBOUND, ST_CIRCULAR, ST_TIDAL, ST_NTRY = range(6)


class Tables(C.Structure):
    _fields_ = [("r", C.c_void_p), ("m", C.c_void_p), ("phi", C.c_void_p), ("N_MAX", C.c_long), ("N_STAR", C.c_long),
                ("cenma_m", C.c_double), ("madhoc", C.c_double)]


class Orbit(C.Structure):
    _fields_ = [("rp", C.c_double), ("ra", C.c_double), ("dQdrp", C.c_double), ("dQdra", C.c_double),
                ("kmax", C.c_long), ("kmin", C.c_long), ("circular_flag", C.c_int), ("used_brent", C.c_int)]


class DynParams(C.Structure):
    _fields_ = [(k, C.c_long) for k in ("relaxation", "ss_collision", "binsingle", "binbin", "tidal_capture", "co_tde", "bh_capture", "p")] + \
               [(k, C.c_double) for k in ("gamma", "coll_factor", "tc_factor", "wd_tc", "units_mstar", "units_l", "units_t")]


class RelaxStats(C.Structure):
    _fields_ = [("Nrel", C.c_long), ("Nrelbeta", C.c_long * 4), ("qaverelbeta", C.c_double * 4), ("maverelbeta", C.c_double * 4),
                ("raverelbeta", C.c_double * 4)]


DYN_DEFAULTS = dict(relaxation=1, ss_collision=0, binsingle=0, binbin=0, tidal_capture=0, co_tde=0, bh_capture=0, p=20, gamma=0.11,
                    coll_factor=1.0, tc_factor=1.0, wd_tc=0.0, units_mstar=1.989e33, units_l=3.0e18, units_t=1.0e13)


def dyn_params(**kw):
    d = dict(DYN_DEFAULTS)
    d.update(kw)
    return DynParams(**d)


def _build():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("cmc_oracle.c", "cmc_oracle.h")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libcmc_oracle.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        _build()
        L = C.CDLL(ORACLE_SO)
        L.orc_find_limits.argtypes = [C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_potential_calculate.restype = C.c_long
        L.orc_potential_calculate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p,
                                              C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_FindZero_r.restype = C.c_long
        L.orc_FindZero_r.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_double]
        L.orc_potential.restype = C.c_double
        L.orc_potential.argtypes = [C.POINTER(Tables), C.c_double]
        L.orc_calc_orbit_new.argtypes = [C.POINTER(Tables), C.c_long, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(Orbit)]
        L.orc_find_root_vr.restype = C.c_double
        L.orc_find_root_vr.argtypes = [C.POINTER(Tables), C.c_long, C.c_long, C.c_double, C.c_double]
        L.orc_get_positions.restype = C.c_double
        L.orc_get_positions.argtypes = [C.POINTER(Tables), C.c_long] + [C.c_void_p] * 7 + [C.c_double, C.c_void_p, C.c_long] + [C.c_void_p] * 12
        L.orc_sort_stars.restype = C.c_long
        L.orc_sort_stars.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p]
        L.orc_energy_conservation2.argtypes = [C.POINTER(Tables), C.c_long] + [C.c_void_p] * 3
        L.orc_intermediate_energy.argtypes = [C.POINTER(Tables), C.c_long] + [C.c_void_p] * 5
        L.orc_energy_conservation1.argtypes = [C.POINTER(Tables), C.c_long] + [C.c_void_p] * 6
        L.orc_set_velocities3_first.restype = C.c_double
        L.orc_set_velocities3_first.argtypes = [C.POINTER(Tables), C.c_long] + [C.c_void_p] * 12
        L.orc_set_velocities3_second.restype = C.c_double
        L.orc_set_velocities3_second.argtypes = [C.POINTER(Tables), C.c_long] + [C.c_void_p] * 3 + [C.c_double] + [C.c_void_p] * 4
        L.orc_calc_sigma_r.argtypes = [C.c_long] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_dynamics_apply_pairs.restype = C.c_long
        L.orc_dynamics_apply_pairs.argtypes = [C.POINTER(Tables), C.c_long, C.c_long, C.c_double, C.POINTER(DynParams)] + [C.c_void_p] * 17 + [C.POINTER(RelaxStats)]
        L.orc_simul_relax_new.restype = C.c_double
        L.orc_simul_relax_new.argtypes = [C.POINTER(Tables), C.c_long, C.c_long, C.c_long, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_tidally_strip_select.restype = C.c_long
        L.orc_tidally_strip_select.argtypes = [C.c_long, C.c_int, C.c_double] + [C.c_void_p] * 6
        L.orc_compute_energy_totals.argtypes = [C.POINTER(Tables), C.c_long, C.c_long] + [C.c_void_p] * 11
        L.orc_comp_multi_mass_percent.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_double, C.c_void_p, C.c_long, C.c_void_p]
        L.orc_clusdyn_calculate.restype = C.c_double
        L.orc_clusdyn_calculate.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_double]
        L.orc_comp_mass_percent.restype = C.c_long
        L.orc_comp_mass_percent.argtypes = ([C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_int]
                                            + [C.c_double] * 4 + [C.c_int, C.c_double, C.c_void_p, C.c_long] + [C.c_void_p] * 10)
        L.orc_t113_seed.argtypes = [C.c_ulong, C.c_void_p]
        L.orc_t113_dbl.restype = C.c_double
        L.orc_t113_dbl.argtypes = [C.c_void_p]
        L.orc_t113_jump.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_t113_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def find_limits(N, blk, procs):
    S = np.zeros(procs, np.int32); E = np.zeros(procs, np.int32)
    lib().orc_find_limits(N, blk, procs, _p(S), _p(E))
    return S, E


def potential_calculate(star_r, star_m, N_MAX_in, N_STAR, cenma_m=0.0, cenma_m_new=0.0, orbit_r=1e6):
    """star_r/star_m: full tables with sentinel at [0]; star_r is modified at [N_MAX+1] like the reference does."""
    phi = np.zeros_like(star_r)
    cm = C.c_double(cenma_m); cmn = C.c_double(cenma_m_new); Mt = C.c_double(); Rt = C.c_double()
    nmax = lib().orc_potential_calculate(_p(star_r), _p(star_m), _p(phi), N_MAX_in, N_STAR, C.addressof(cm), C.addressof(cmn),
                                         1.0 / N_STAR, orbit_r, C.addressof(Mt), C.addressof(Rt))
    return dict(N_MAX=nmax, phi=phi, Mtotal=Mt.value, Rtidal=Rt.value, cenma_m=cm.value)


def tables(star_r, star_m, star_phi, N_MAX, N_STAR, cenma_m=0.0):
    t = Tables(_p(star_r), _p(star_m), _p(star_phi), N_MAX, N_STAR, cenma_m, 1.0 / N_STAR)
    t._keep = (star_r, star_m, star_phi)
    return t


def get_positions(t, g_idx, E, J, vr, vt, Rtidal, draws, csr_off=None, is_new=None, rnew_in=None):
    n = len(g_idx)
    g_idx = np.ascontiguousarray(g_idx, np.int64)
    out = dict(draw_off=np.zeros(n + 1, np.int64), status=np.zeros(n, np.int32), rnew=np.zeros(n), vrnew=np.zeros(n), vtnew=np.zeros(n),
               X=np.zeros(n), r_peri=np.zeros(n), r_apo=np.zeros(n), kmin=np.zeros(n, np.int64), kmax=np.zeros(n, np.int64))
    nb = C.c_long(0)
    draws = np.ascontiguousarray(draws, np.float64)
    if csr_off is not None:
        csr_off = np.ascontiguousarray(csr_off, np.int64)
    out["max_rad"] = lib().orc_get_positions(
        C.byref(t), n, _p(g_idx), _p(is_new), _p(np.ascontiguousarray(E)), _p(np.ascontiguousarray(J)), _p(np.ascontiguousarray(vr)),
        _p(np.ascontiguousarray(vt)), _p(rnew_in), Rtidal, _p(draws), len(draws), _p(csr_off), _p(out["draw_off"]),
        _p(out["status"]), _p(out["rnew"]), _p(out["vrnew"]), _p(out["vtnew"]), _p(out["X"]), _p(out["r_peri"]), _p(out["r_apo"]),
        _p(out["kmin"]), _p(out["kmax"]), C.addressof(nb))
    out["n_brent"] = nb.value
    return out


def sort_stars(star1, n, sizeof_star, key_off, binind_off, binary1=None, n_bin_dim=0, sizeof_binary=464):
    """star1: uint8 array (>= n*sizeof_star) starting at element 1; sorted in place.  Returns (n_new, n_bin, perm)."""
    perm = np.zeros(n, np.int64)
    nb = C.c_long(0)
    n_new = lib().orc_sort_stars(_p(star1), n, sizeof_star, key_off, binind_off, _p(binary1), n_bin_dim, sizeof_binary, C.addressof(nb), _p(perm))
    return n_new, nb.value, perm


class Taus113:
    JPOLY_2_80 = np.array([0x487cf69c, 0x00be6310, 0x04bfe2bb, 0x000824f9], np.uint32)
    JPOLY_2_20 = np.array([0x0c382e31, 0x1b040425, 0x0b49a509, 0x0173f6b0], np.uint32)

    def __init__(self, seed, rank=0):
        self.z = np.zeros(4, np.uint32)
        lib().orc_t113_seed(seed, _p(self.z))
        for _ in range(rank):
            self.jump(self.JPOLY_2_80)

    def jump(self, poly):
        lib().orc_t113_jump(_p(self.z), _p(poly))

    def dbl(self):
        return lib().orc_t113_dbl(_p(self.z))

    def draws(self, n):
        out = np.empty(n)
        lib().orc_t113_fill(_p(self.z), _p(out), n)
        return out


MASS_PC_FIELDS = ("mass_r", "ave_mass_r", "no_star_r", "densities_r", "ke_rad_r", "ke_tan_r", "v2_rad_r", "v2_tan_r")


def comp_mass_percent(star_r, star_m, N_MAX, N_STAR, vr, vt, mass_pc, Mtotal, first_global=1, prefix=None, is_rank0=True, cenma_m=0.0,
                      initial_total_mass=1.0, bh_include=False, minimum_r=0.0):
    """comp_mass_percent as one rank sees it (vr, vt: this rank's local stars).  Returns the eight arrays (entries never reached stay -1),
    `owner`, `totals` (the rank's contribution to the Exscan) and `filled`."""
    vr, vt = np.ascontiguousarray(vr, np.float64), np.ascontiguousarray(vt, np.float64)
    pc = np.ascontiguousarray(mass_pc, np.float64)
    n = len(pc)
    out = {f: np.full(n, -1.0) for f in MASS_PC_FIELDS}
    owner, totals = np.zeros(n, np.int32), np.zeros(4)
    pre = np.zeros(4) if prefix is None else np.ascontiguousarray(prefix, np.float64)
    filled = lib().orc_comp_mass_percent(_p(star_r), _p(star_m), N_MAX, N_STAR, _p(vr), _p(vt), len(vr), first_global, _p(pre), int(is_rank0),
                                         cenma_m, 1.0 / N_STAR, Mtotal, initial_total_mass, int(bh_include), minimum_r, _p(pc), n,
                                         *[_p(out[f]) for f in MASS_PC_FIELDS], _p(owner), _p(totals))
    out.update(owner=owner, totals=totals, filled=int(filled))
    return out


def energy_conservation2(t, g_idx, rnew):
    g_idx, rnew = np.ascontiguousarray(g_idx, np.int64), np.ascontiguousarray(rnew, np.float64)
    out = np.zeros(len(rnew))
    lib().orc_energy_conservation2(C.byref(t), len(rnew), _p(g_idx), _p(rnew), _p(out))
    return out


def intermediate_energy(t, g_idx, rnew, vr, vt, EI_in):
    g_idx = np.ascontiguousarray(g_idx, np.int64)
    rnew, vr, vt = (np.ascontiguousarray(x, np.float64) for x in (rnew, vr, vt))
    out = np.array(EI_in, dtype=np.float64, copy=True)
    lib().orc_intermediate_energy(C.byref(t), len(rnew), _p(g_idx), _p(rnew), _p(vr), _p(vt), _p(out))
    return out


def energy_conservation1(t, g_idx, vr, vt):
    """(vrold, vtold, Uoldrold) of energy_conservation1 (reference cmc_utils.c:2153-2174)"""
    g_idx = np.ascontiguousarray(g_idx, np.int64)
    vr, vt = np.ascontiguousarray(vr, np.float64), np.ascontiguousarray(vt, np.float64)
    out = [np.zeros(len(vr)) for _ in range(3)]
    lib().orc_energy_conservation1(C.byref(t), len(vr), _p(g_idx), _p(vr), _p(vt), *[_p(o) for o in out])
    return dict(vrold=out[0], vtold=out[1], Uoldrold=out[2])


SV3_IN = ("interacted", "rOld", "vrold", "vtold", "Uoldrold", "Uoldrnew")
SV3_IO = ("vr", "vt", "vrnew", "vtnew")


def set_velocities3_first(t, g_idx, f):
    """first loop of set_velocities3 (reference cmc_utils.c:268-311) on the fields of dict f (SV3_IN + SV3_IO, local star order);
    returns (fields after, vnew2_arr, the rank's leftover Eexcess)"""
    g_idx = np.ascontiguousarray(g_idx, np.int64)
    n = len(g_idx)
    inter = np.ascontiguousarray(f["interacted"], np.int64)
    ins = [np.ascontiguousarray(f[k], np.float64) for k in SV3_IN[1:]]
    io = {k: np.array(f[k], dtype=np.float64, copy=True) for k in SV3_IO}
    flags = np.zeros(n, np.int8)
    left = lib().orc_set_velocities3_first(C.byref(t), n, _p(g_idx), _p(inter), *[_p(a) for a in ins], *[_p(io[k]) for k in SV3_IO], _p(flags))
    return io, flags, left


def set_velocities3_second(t, g_idx, interacted, flags, Eexcess, io):
    """one pass of the do-while loop of set_velocities3 (reference cmc_utils.c:323-343) starting from the received Eexcess; io is updated
    in place (vr, vt); returns what is left"""
    g_idx, inter = np.ascontiguousarray(g_idx, np.int64), np.ascontiguousarray(interacted, np.int64)
    return lib().orc_set_velocities3_second(C.byref(t), len(g_idx), _p(g_idx), _p(inter), _p(flags), Eexcess, _p(io["vr"]), _p(io["vt"]),
                                            _p(io["vrnew"]), _p(io["vtnew"]))


def set_velocities3(t, shards, fields):
    """set_velocities3 as P ranks run it (reference cmc_utils.c:254-352).  shards = [(first_global, n_local)] in rank order, fields = one
    dict per rank.  Returns (per-rank fields after, per-rank -Eexcess*madhoc added to Eoops, passes of the do-while loop)."""
    P = len(shards)
    g = [np.arange(s, s + n) for s, n in shards]
    first = [set_velocities3_first(t, g[q], fields[q]) for q in range(P)]
    io, flags, E = [x[0] for x in first], [x[1] for x in first], [x[2] for x in first]
    passes = 0
    while True:
        prev = [0.0] + E[:-1]                                     # rank i sends to i+1 (:317-320); rank 0 keeps Eexcess_prev = 0
        E = [set_velocities3_second(t, g[q], fields[q]["interacted"], flags[q], prev[q], io[q]) for q in range(P)]
        passes += 1
        if not sum(E) > 0:                                        # MPI_Allreduce(SUM), :345-347
            break
    return io, [-e * t.madhoc for e in E], passes


def calc_sigma_r(p, star_r, star_m, N_MAX, N_STAR, vr_all, vt_all, first_global=1, n_local=None, mave=False):
    """calc_sigma_r (reference cmc_dynamics_helper.c:2821-2975) as the rank owning global stars first_global .. first_global + n_local - 1
    sees it.  vr_all, vt_all: the velocities of the WHOLE sorted cluster (star k at [k - 1]); the rank's own stars and the p ghost stars on
    either side (what the reference exchanges with MPI_Send / MPI_Recv) are cut from them.  Returns (r or mean mass, sigma), each
    [n_local + 1] with entry 0 unused."""
    vr_all, vt_all = np.ascontiguousarray(vr_all, np.float64), np.ascontiguousarray(vt_all, np.float64)
    lo = first_global - 1
    n = len(vr_all) - lo if n_local is None else n_local

    def cut(x, start):                                        # p values from x[start:], zero-padded where the cluster ends
        g = np.zeros(p)
        if start >= 0:
            seg = x[start:start + p]
            g[:len(seg)] = seg
        return g
    loc = [np.ascontiguousarray(x[lo:lo + n]) for x in (vr_all, vt_all)]
    prev = [cut(x, lo - p) for x in (vr_all, vt_all)]
    nxt = [cut(x, lo + n) for x in (vr_all, vt_all)]
    a, sg = np.zeros(n + 1), np.zeros(n + 1)
    lib().orc_calc_sigma_r(p, n, first_global, N_MAX, int(first_global == 1), _p(star_r), _p(star_m), 1.0 / N_STAR, _p(loc[0]), _p(loc[1]),
                           _p(prev[0]), _p(prev[1]), _p(nxt[0]), _p(nxt[1]), int(mave), _p(a), _p(sg))
    return a, sg


def dynamics_apply_pairs(t, first_global, dt, dp, draws, f, bin_a=None):
    """dynamics_apply's pair loop without the strong encounters (reference cmc_dynamics.c:206-500).  f: dict of the local stars' fields vr, vt,
    E, J, Y (float64, modified in place), binind (int64), se_k (int32), rad, optionally threebb (int64).  draws: 3 per pair."""
    n = len(f["vr"])
    npmax = n // 2 + 1
    out = dict(pair_k=np.zeros(npmax, np.int64), pair_kp=np.zeros(npmax, np.int64), status=np.zeros(npmax, np.int32), rperi=np.zeros(npmax),
               W=np.zeros(npmax), rcm=np.zeros(npmax))
    st = RelaxStats()
    draws = np.ascontiguousarray(draws, np.float64)
    if len(draws) < 3 * (n // 2):                     # fewer pairs than n / 2 when threebb-flagged stars are skipped
        draws = np.concatenate([draws, np.zeros(3 * (n // 2) - len(draws))])
    for k in ("vr", "vt", "E", "J", "Y", "rad"):
        assert f[k].dtype == np.float64 and f[k].flags.c_contiguous, k
    assert f["binind"].dtype == np.int64 and f["se_k"].dtype == np.int32
    thr = f.get("threebb")
    npairs = lib().orc_dynamics_apply_pairs(C.byref(t), n, first_global, dt, C.byref(dp), _p(draws), _p(f["vr"]), _p(f["vt"]), _p(f["E"]), _p(f["J"]),
                                            _p(f["Y"]), _p(f["binind"]), _p(f["se_k"]), _p(f["rad"]), _p(bin_a), _p(thr), _p(out["pair_k"]),
                                            _p(out["pair_kp"]), _p(out["status"]), _p(out["rperi"]), _p(out["W"]), _p(out["rcm"]), C.byref(st))
    out = {k: v[:npairs] for k, v in out.items()}
    out["Nrel"] = st.Nrel
    out["Nrelbeta"] = np.array(st.Nrelbeta[:])
    out["qave"], out["mave"], out["rave"] = np.array(st.qaverelbeta[:]), np.array(st.maverelbeta[:]), np.array(st.raverelbeta[:])
    return out


def simul_relax_new(t, first_global, p, thetasemax, vr, vt):
    """simul_relax_new (reference cmc_dynamics_helper.c:2526-2580) as the rank owning the local stars vr / vt sees it"""
    vr, vt = np.ascontiguousarray(vr, np.float64), np.ascontiguousarray(vt, np.float64)
    return lib().orc_simul_relax_new(C.byref(t), len(vr), first_global, p, thetasemax, _p(vr), _p(vt))


def tidally_strip_select(tidal_treatment, threshold, f):
    """one pass of tidally_strip_stars' criterion over the local stars (reference cmc_evolution_thr.c:467-609); f["rnew"], ["vrnew"], ["vtnew"]
    are updated in place; returns the 1-based local indices of the stripped stars in star order"""
    n = len(f["rnew"])
    idx = np.zeros(n, np.int64)
    r_apo, E = np.ascontiguousarray(f["r_apo"], np.float64), np.ascontiguousarray(f["E"], np.float64)
    c = lib().orc_tidally_strip_select(n, tidal_treatment, threshold, _p(r_apo), _p(E), _p(f["rnew"]), _p(f["vrnew"]), _p(f["vtnew"]), _p(idx))
    return idx[:c].copy()


def tidally_strip_loop(select, potential, star_m_of, star_fields_of, binary_of, s, N_STAR, orbit_r, tidal_treatment, GAMMA):
    """tidally_strip_stars on one rank: see tidally_strip_loop_ranks"""
    return tidally_strip_loop_ranks([select], potential, [star_m_of], [star_fields_of], [binary_of], [s], N_STAR, orbit_r, tidal_treatment, GAMMA)[0]


def tidally_strip_loop_ranks(selects, potential, star_m_ofs, star_fields_ofs, binary_ofs, states, N_STAR, orbit_r, tidal_treatment, GAMMA):
    """The do-while loop of tidally_strip_stars (reference cmc_evolution_thr.c:443-645) around a per-pass selection, as P ranks run it, with the
    reference's bookkeeping per stripped star in star order (:484-496 / :560-572).  Per rank q: selects[q](threshold) -> 1-based local
    indices; star_m_ofs[q](i) -> star_m[g_i]; star_fields_ofs[q](i) -> dict(E, J, Eint, binind); binary_ofs[q](k) -> dict(m1, m2, a, Eint1,
    Eint2); states[q]: dict with Mtotal, TidalMassLoss, OldTidalMassLoss (the same on every rank) and the rank's Eescaped, Jescaped,
    Eintescaped, Ebescaped -- updated in place, plus Rtidal, Etidal, DTidalMassLoss, passes, stripped.  potential(r) -> Phi(r) (replicated
    tables).  The mass stripped in a pass is summed over the ranks (:612-616, rank order) and the five escaper sums at the end (:626-644)."""
    import math
    P = len(states)
    for s in states:
        s["Etidal"], s["passes"], s["stripped"] = 0.0, 0, []
    N = float(N_STAR)
    while True:
        s0 = states[0]
        Rt = orbit_r * math.pow(s0["Mtotal"] - (s0["TidalMassLoss"] - s0["OldTidalMassLoss"]), 1.0 / 3.0)
        phi_rtidal = potential(Rt)
        thr = Rt if tidal_treatment == 0 else (1.5 - 3.0 * math.pow(math.log(GAMMA * N) / N, 0.25)) * phi_rtidal
        Ds = []
        for q in range(P):
            s = states[q]
            s["Rtidal"] = Rt
            D = 0.0
            for i in selects[q](thr):
                m, f = star_m_ofs[q](i), star_fields_ofs[q](i)
                s["Eescaped"] += f["E"] * m / N
                s["Jescaped"] += f["J"] * m / N
                if f["binind"] == 0:
                    s["Eintescaped"] += f["Eint"]
                else:
                    b = binary_ofs[q](f["binind"])
                    s["Ebescaped"] += -(b["m1"] / N) * (b["m2"] / N) / (2.0 * b["a"])
                    s["Eintescaped"] += b["Eint1"] + b["Eint2"]
                D += m / N
                s["Etidal"] += f["E"] * m / N
                s["stripped"].append(int(i))
            Ds.append(D)
        Dtot = Ds[0]
        for q in range(1, P):
            Dtot = Dtot + Ds[q]
        for s in states:
            s["passes"] += 1
            s["TidalMassLoss"] += Dtot
            s["DTidalMassLoss"] = Dtot
        if not Dtot > 0:
            break
    for k in ("Eescaped", "Jescaped", "Eintescaped", "Ebescaped", "Etidal"):
        tot = states[0][k]
        for q in range(1, P):
            tot = tot + states[q][k]
        for s in states:
            s[k] = tot
    return states


def compute_energy_totals(t, first_global, vr, vt, Eint, binind, bins):
    """(K, P before the factor 0.5, Eint, Eb) of ComputeEnergy's second loop (reference cmc_utils.c:619-636) over one rank's stars;
    bins: dict of arrays indexed by binind (inuse int32, m1, m2, a, Eint1, Eint2)"""
    vr, vt, Eint = (np.ascontiguousarray(x, np.float64) for x in (vr, vt, Eint))
    binind = np.ascontiguousarray(binind, np.int64)
    b = {k: np.ascontiguousarray(bins[k], np.int32 if k == "inuse" else np.float64) for k in ("inuse", "m1", "m2", "a", "Eint1", "Eint2")}
    out = np.zeros(4)
    lib().orc_compute_energy_totals(C.byref(t), len(vr), first_global, _p(vr), _p(vt), _p(Eint), _p(binind), _p(b["inuse"]), _p(b["m1"]), _p(b["m2"]),
                                    _p(b["a"]), _p(b["Eint1"]), _p(b["Eint2"]), _p(out))
    return out


def comp_multi_mass_percent(star_r, star_m, N_MAX, mass_bins, solar_mass_dyn, mass_pc):
    """Lagrange radii per mass bin (reference cmc_utils.c:769-861) -> [n_bins, count]; entries of empty bins stay -1"""
    bins, pc = np.ascontiguousarray(mass_bins, np.float64), np.ascontiguousarray(mass_pc, np.float64)
    out = np.full((len(bins), len(pc)), -1.0)
    lib().orc_comp_multi_mass_percent(_p(star_r), _p(star_m), N_MAX, _p(bins), len(bins), solar_mass_dyn, _p(pc), len(pc), _p(out))
    return out


def clusdyn_calculate(star_r, star_m, N_STAR, Mtotal):
    """clusdyn.rh, the half-mass radius (reference cmc_utils.c:1613-1624)"""
    return lib().orc_clusdyn_calculate(_p(star_r), _p(star_m), N_STAR, Mtotal)
