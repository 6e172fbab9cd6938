"""ctypes binding to oracle/_ref/libcmc_ref.so -- the UNMODIFIED reference hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's reference/cpu_baseline legs and
__graft_entry__.smoke(); never by the product package.  The library is built by
oracle/Makefile from the sources under /root/reference (container) and travels to the GPU box prebuilt.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "..", "oracle", "_ref", "libcmc_ref.so")
# the same reference objects + harness with the four hot-path entry points re-bound to cmc_cosmic_b200/shim/cmcb200_shim.c (ld --wrap)
REF_B200_SO = os.path.join(_HERE, "..", "oracle", "_ref", "libcmc_ref_b200.so")

SF_INFINITY = 1.0e10
ZERO = 1.0e-20


def available():
    return os.path.exists(REF_SO)


class Ref:
    """One in-process rank of the reference (procs == 1)."""

    def __init__(self, so=REF_SO):
        L = C.CDLL(so)
        self.L = L
        L.ref_sizeof_star.restype = C.c_long
        L.ref_sizeof_binary.restype = C.c_long
        L.ref_star_offset.restype = C.c_long
        L.ref_star_offset.argtypes = [C.c_char_p]
        L.ref_binary_offset.restype = C.c_long
        L.ref_binary_offset.argtypes = [C.c_char_p]
        L.ref_init.argtypes = [C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, C.c_int, C.c_long, C.c_double]
        for f in ("ref_star", "ref_binary", "ref_star_r", "ref_star_m", "ref_star_phi", "ref_draw_log", "ref_Start", "ref_End"):
            getattr(L, f).restype = C.c_void_p
        L.ref_set_scalars.argtypes = [C.c_long, C.c_long, C.c_long, C.c_double, C.c_double, C.c_double, C.c_double]
        L.ref_get_scalars.argtypes = [C.c_void_p]
        L.ref_draw_count.restype = C.c_long
        L.ref_set_draws.argtypes = [C.c_void_p, C.c_long]
        L.ref_potential.restype = C.c_double
        L.ref_potential.argtypes = [C.c_double]
        L.ref_find_zero_Q.restype = C.c_long
        L.ref_find_zero_Q.argtypes = [C.c_long, C.c_long, C.c_long, C.c_double, C.c_double]
        L.ref_function_q.restype = C.c_double
        L.ref_function_q.argtypes = [C.c_long, C.c_double, C.c_double, C.c_double, C.c_double]
        L.ref_calc_orbit_new.argtypes = [C.c_long, C.c_double, C.c_double, C.c_void_p]
        L.ref_find_root_vr.restype = C.c_double
        L.ref_find_root_vr.argtypes = [C.c_long, C.c_long, C.c_double, C.c_double]
        L.ref_find_limits.argtypes = [C.c_long, C.c_int]
        L.ref_rng_reset.argtypes = [C.c_ulong, C.c_void_p]
        L.ref_rng_jump80.argtypes = [C.c_void_p]
        L.ref_rng_next.restype = C.c_double
        L.ref_rng_next.argtypes = [C.c_void_p]
        L.ref_time_step.argtypes = [C.c_void_p] * 4
        L.ref_spawn.argtypes = [C.c_int, C.c_size_t]
        L.ref_get_Eoops.restype = C.c_double
        L.ref_set_Eoops.argtypes = [C.c_double]
        self.sizeof_star = L.ref_sizeof_star()
        self.sizeof_binary = L.ref_sizeof_binary()
        self.n_star_dim = 0

    def star_offset(self, name):
        return self.L.ref_star_offset(name.encode())

    def binary_offset(self, name):
        return self.L.ref_binary_offset(name.encode())

    def init(self, n_star, n_star_dim=None, n_bin_dim=None, min_chunk=40, samplesize=1024, idum=1234, orbit_r=1e6, n_table_dim=None):
        n_star_dim = n_star_dim or int(1.5 * (n_star + 2)) + 16
        n_table_dim = n_table_dim or max(n_star_dim, n_star + 16)
        n_bin_dim = n_bin_dim or 16
        rc = self.L.ref_init(n_star, n_star_dim, n_table_dim, n_bin_dim, min_chunk, samplesize, idum, orbit_r)
        assert rc == 0
        self.n_star = n_star
        self.n_star_dim = n_star_dim
        self.n_table_dim = n_table_dim
        self.n_bin_dim = n_bin_dim

    def free(self):
        self.L.ref_free()

    # --- views onto the reference's own arrays (no copies) ---
    def _view(self, ptr, nbytes, dtype):
        buf = (C.c_char * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype)

    def star_bytes(self):
        return self._view(self.L.ref_star(), self.n_star_dim * self.sizeof_star, np.uint8).reshape(self.n_star_dim, self.sizeof_star)

    def binary_bytes(self):
        return self._view(self.L.ref_binary(), self.n_bin_dim * self.sizeof_binary, np.uint8).reshape(self.n_bin_dim, self.sizeof_binary)

    def star_field(self, name, dtype=np.float64):
        off = self.star_offset(name)
        assert off >= 0, name
        raw = self._view(self.L.ref_star(), self.n_star_dim * self.sizeof_star, np.uint8)
        return np.lib.stride_tricks.as_strided(raw[off:off + 8].view(dtype), shape=(self.n_star_dim,), strides=(self.sizeof_star,))

    def star_r(self):
        return self._view(self.L.ref_star_r(), self.n_table_dim * 8, np.float64)

    def star_m(self):
        return self._view(self.L.ref_star_m(), self.n_table_dim * 8, np.float64)

    def star_phi(self):
        return self._view(self.L.ref_star_phi(), self.n_table_dim * 8, np.float64)

    def scalars(self):
        out = np.zeros(16)
        self.L.ref_get_scalars(out.ctypes.data)
        keys = ["N_MAX", "N_MAX_NEW", "Mtotal", "Rtidal", "cenma_m", "max_r", "TidalMassLoss", "Eescaped", "Jescaped",
                "Eintescaped", "Ebescaped", "Etidal", "mpiBegin", "mpiEnd", "N_b_local", "esc_bytes"]
        return dict(zip(keys, out))

    def set_scalars(self, N_MAX, N_MAX_NEW, N_BINARY=0, cenma_m=0.0, cenma_m_new=0.0, Rtidal=1e6, Mtotal=1.0):
        self.L.ref_set_scalars(N_MAX, N_MAX_NEW, N_BINARY, cenma_m, cenma_m_new, Rtidal, Mtotal)

    def draws(self):
        n = self.L.ref_draw_count()
        return self._view(self.L.ref_draw_log(), n * 8, np.float64).copy()

    # --- convenience: load a sorted cluster the way the reference's start-up does ---
    def load_cluster(self, r, vr, vt, m_star_units, binind=None, ids=None):
        """r sorted ascending (N,), m in star_m units (obj_m * N_STAR)."""
        n = len(r)
        assert n == self.n_star
        sr, sm = self.star_r(), self.star_m()
        sr[0] = ZERO; sm[0] = 0.0
        sr[1:n + 1] = r; sm[1:n + 1] = m_star_units
        sr[n + 1] = SF_INFINITY; sm[n + 1] = 0.0
        for name, val in (("r", r), ("vr", vr), ("vt", vt), ("m", m_star_units)):
            self.star_field(name)[1:n + 1] = val
        self.star_field("r")[0] = ZERO
        self.star_field("r")[n + 1] = SF_INFINITY
        self.star_field("id", np.int64)[1:n + 1] = np.arange(1, n + 1) if ids is None else ids
        if binind is not None:
            self.star_field("binind", np.int64)[1:n + 1] = binind

    def clusdyn_rh(self):
        self.L.ref_clusdyn_rh.restype = C.c_double
        return self.L.ref_clusdyn_rh()

    def comp_mass_percent(self, mass_pc, bh_include=False, minimum_r=0.0, initial_total_mass=1.0):
        """the reference's comp_mass_percent on the harness's state; entries it never reaches stay -1"""
        pc = np.ascontiguousarray(mass_pc, np.float64)
        out = np.zeros((8, len(pc)))
        self.L.ref_comp_mass_percent.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_double, C.c_double, C.c_void_p]
        self.L.ref_comp_mass_percent(pc.ctypes.data, len(pc), int(bh_include), minimum_r, initial_total_mass, out.ctypes.data)
        names = ("mass_r", "ave_mass_r", "no_star_r", "densities_r", "ke_rad_r", "ke_tan_r", "v2_rad_r", "v2_tan_r")
        return {k: out[i].copy() for i, k in enumerate(names)}

    def calc_sigma_r(self, p, n_limit, mave=False):
        """the reference's calc_sigma_r (cmc_dynamics_helper.c:2821-2975) on the harness's state -> (r or mean mass, sigma), [n_limit + 1]"""
        a, sg = np.zeros(n_limit + 1), np.zeros(n_limit + 1)
        self.L.ref_calc_sigma_r.argtypes = [C.c_long, C.c_long, C.c_int, C.c_void_p, C.c_void_p]
        self.L.ref_calc_sigma_r(p, n_limit, int(mave), a.ctypes.data, sg.ctypes.data)
        return a, sg

    DYN_FLAGS = ("relaxation", "ss_collision", "binsingle", "binbin", "tidal_capture", "tc_polytrope", "co_tde", "bh_capture", "bh_loss_cone",
                 "threebodybinaries", "p")
    DYN_PARS = ("gamma", "coll_factor", "tc_factor", "wd_tc", "units_mstar", "units_l", "units_t")

    def dynamics_apply(self, dt, cfg, log_cap=1 << 16):
        """the reference's dynamics_apply (cmc_dynamics.c:17-542) on the harness's state with the strong-encounter routines re-bound to
        recorders (oracle/ref_harness.c).  cfg: dict with the keys of DYN_FLAGS / DYN_PARS.  Returns (encounter log [(type, k, kp)],
        [(rperi, W, rcm)], the relaxation file's line)."""
        flags = np.array([int(cfg.get(k, 0)) for k in self.DYN_FLAGS], np.int64)
        par = np.array([float(cfg[k]) for k in self.DYN_PARS], np.float64)
        lk, ld = np.zeros((log_cap, 3), np.int64), np.zeros((log_cap, 3))
        txt = C.create_string_buffer(4096)
        self.L.ref_dynamics_apply.restype = C.c_long
        self.L.ref_dynamics_apply.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        n = self.L.ref_dynamics_apply(dt, flags.ctypes.data, par.ctypes.data, lk.ctypes.data, ld.ctypes.data, log_cap, txt, 4096)
        assert n <= log_cap
        return lk[:n], ld[:n], txt.value.decode()

    def tidally_strip_stars(self, treatment, gamma=0.11):
        """the reference's tidally_strip_stars (cmc_evolution_thr.c:443-645) on the harness's state -> (Rtidal, TidalMassLoss, DTidalMassLoss)"""
        out = np.zeros(3)
        self.L.ref_tidally_strip_stars.argtypes = [C.c_long, C.c_double, C.c_void_p]
        self.L.ref_tidally_strip_stars(treatment, gamma, out.ctypes.data)
        return tuple(out)

    def comp_multi_mass_percent(self, mass_bins, solar_mass_dyn, mass_pc):
        """the reference's comp_multi_mass_percent (cmc_utils.c:769-861) on the harness's tables -> [n_bins, count]"""
        bins, pc = np.ascontiguousarray(mass_bins, np.float64), np.ascontiguousarray(mass_pc, np.float64)
        out = np.zeros((len(bins), len(pc)))
        self.L.ref_comp_multi_mass_percent.argtypes = [C.c_void_p, C.c_long, C.c_double, C.c_void_p, C.c_long, C.c_void_p]
        self.L.ref_comp_multi_mass_percent(bins.ctypes.data, len(bins), solar_mass_dyn, pc.ctypes.data, len(pc), out.ctypes.data)
        return out

    def write_snapshot(self, units_m, units_l, sigma=None, avekernel=20, bh_only=False):
        """the reference's write_snapshot (cmc_io.c:2960-3345) on the harness's state with the HDF5 calls re-bound to recorders
        (oracle/ref_harness.c): the rows it hands to H5TBmake_table as a structured array, plus the table's geometry
        [(name, offset, H5 type code)]"""
        from cmc_cosmic_b200 import layout
        L = self.L
        L.ref_write_snapshot.restype = C.c_long
        L.ref_write_snapshot.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_long, C.c_long]
        L.ref_snapshot_rows.restype = C.c_void_p
        L.ref_snapshot_row_size.restype = C.c_long
        L.ref_snapshot_nfields.restype = C.c_long
        L.ref_snapshot_field.restype = C.c_long
        L.ref_snapshot_field.argtypes = [C.c_long, C.c_char_p, C.c_void_p]
        sg = None if sigma is None else np.ascontiguousarray(sigma[1:], np.float64)
        n = L.ref_write_snapshot(int(bh_only), units_m, units_l, None if sg is None else sg.ctypes.data, 0 if sg is None else len(sg), avekernel)
        assert L.ref_snapshot_row_size() == layout.SNAPSHOT_DTYPE.itemsize
        rows = self._view(L.ref_snapshot_rows(), max(n, 1) * layout.SNAPSHOT_DTYPE.itemsize, layout.SNAPSHOT_DTYPE)[:n].copy()
        geom = []
        for i in range(L.ref_snapshot_nfields()):
            name, typ = C.create_string_buffer(64), C.c_long(0)
            off = L.ref_snapshot_field(i, name, C.byref(typ))
            geom.append((name.value.decode(), off, typ.value))
        return rows, geom

    def compute_energy(self):
        """E, J exactly as ComputeEnergy (reference cmc_utils.c:614-615), single rank."""
        n = int(self.scalars()["N_MAX"])
        phi, r = self.star_phi(), self.star_r()
        vr, vt = self.star_field("vr"), self.star_field("vt")
        self.star_field("E")[1:n + 1] = phi[1:n + 1] + 0.5 * (vr[1:n + 1] * vr[1:n + 1] + vt[1:n + 1] * vt[1:n + 1])
        self.star_field("J")[1:n + 1] = r[1:n + 1] * vt[1:n + 1]
