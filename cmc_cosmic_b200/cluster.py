"""Host-side mirror of the reference's interface for the sort / potential / new-orbit path.

The reference exposes this path as four ``void f(void)`` functions working on process globals
(reference src/cmc/cmc.c:311,335,339,344; globals in include/cmc/cmc_vars.h).  :class:`Cluster` holds those globals as
attributes with the reference's names (``star``, ``binary``, ``star_r``, ``star_m``, ``star_phi``, ``clus``, ``cenma``,
``Mtotal``, ``Rtidal``, ``max_r`` ...) and offers the four functions under their own names.  Each of them forwards to
the C ABI of ``libcmcb200.so`` -- the same entry points the C shim in INTEGRATION.md binds.

Two ways to run:
  * host-buffer form (``new_orbits_calculate`` / ``qsorts_new`` / ``post_sort_comm`` / ``calc_potential_new``): the drop-in
    behaviour, every call copies its inputs to the GPU and its outputs back;
  * resident form (``upload`` once, then ``dev_*`` / ``dev_step``): the arrays stay in HBM across timesteps.
"""
import ctypes as C
from types import SimpleNamespace

import numpy as np

from . import lib as _lib
from .layout import (BINARY_DTYPE_SIZE, SF_INFINITY, STAR_DTYPE_SIZE, ZERO, binary_field, star_field)


def _host_array(shape, dtype, pinned):
    if pinned:
        import torch
        t = torch.zeros(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
        return t.numpy()    # the view keeps the pinned tensor alive
    return np.zeros(shape, dtype)


class Cluster:
    def __init__(self, n_star, n_star_dim=None, n_bin_dim=None, device=0, min_chunk=40, orbit_r=1.0e6, pinned=False, stream=None, n_table_dim=None):
        L = _lib.load()
        self.L = L
        # N_STAR_DIM_OPT = 1.5 * (1 + N/P + 2 N_bin/P), reference src/cmc/cmc_io.c:2092-2099
        self.N_STAR_DIM = int(n_star_dim or int(1.5 * (n_star + 2)) + 16)
        self.N_BIN_DIM = int(n_bin_dim or 16)
        self.N_TABLE_DIM = int(n_table_dim or self.N_STAR_DIM)   # star_r / star_m / star_phi hold the whole cluster (N_STAR_DIM in the reference)
        self.device = device
        self.MIN_CHUNK_SIZE = min_chunk
        self.star = _host_array((self.N_STAR_DIM, STAR_DTYPE_SIZE), np.uint8, pinned)
        self.binary = _host_array((self.N_BIN_DIM, BINARY_DTYPE_SIZE), np.uint8, pinned)
        self.star_r = _host_array((self.N_TABLE_DIM,), np.float64, pinned)
        self.star_m = _host_array((self.N_TABLE_DIM,), np.float64, pinned)
        self.star_phi = _host_array((self.N_TABLE_DIM,), np.float64, pinned)
        self.clus = SimpleNamespace(N_STAR=n_star, N_STAR_NEW=n_star, N_MAX=n_star, N_MAX_NEW=n_star, N_BINARY=0)
        self.cenma = SimpleNamespace(m=0.0, m_new=0.0, E=0.0, E_new=0.0)
        self.madhoc = 1.0 / float(n_star)           # reference src/cmc/cmc_utils.c:1183
        self.orbit_r = orbit_r
        self.Mtotal = 1.0
        self.Rtidal = orbit_r
        self.max_r = 0.0
        self.N_b_local = 0
        self.bin_slots_in_use = None    # None: unknown, derived from star[].binind at the next call that needs it
        self.myid, self.procs = 0, 1
        self.Start, self.End = np.array([1], np.int32), np.array([n_star], np.int32)
        self.mpiBegin, self.mpiEnd = 1, n_star
        self.tcount = 1
        self.Eoops = 0.0                  # energy lost to the scheme (reference include/cmc/cmc_vars.h:293)
        # escaper bookkeeping of remove_star / tidally_strip_stars (include/cmc/cmc_vars.h)
        self.Eescaped = self.Jescaped = self.Eintescaped = self.Ebescaped = self.Etidal = 0.0
        self.TidalMassLoss = self.OldTidalMassLoss = self.DTidalMassLoss = 0.0
        self.last_orbit = None
        self.last_status = None
        ctx = C.c_void_p()
        _lib.check(L.cmcb200_create_sharded(C.byref(ctx), device, self.N_STAR_DIM, self.N_TABLE_DIM, self.N_BIN_DIM), "cmcb200_create")
        self.ctx = ctx
        if stream is not None:
            _lib.check(L.cmcb200_set_stream(ctx, C.c_void_p(stream)), "cmcb200_set_stream")

    def close(self):
        if getattr(self, "ctx", None):
            self.L.cmcb200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- field views -------------------------------------------------------------------------------------------
    def field(self, name):
        return star_field(self.star, name)

    def bfield(self, name):
        return binary_field(self.binary, name)

    def load(self, r, vr, vt, m_star_units, binind=None, ids=None):
        """Fill the arrays from a radius-sorted cluster the way the reference's start-up does
        (sentinels: src/cmc/cmc_fits.c:128-136).  m_star_units = obj_m * N_STAR."""
        n = len(r)
        assert n == self.clus.N_STAR and n + 2 <= self.N_STAR_DIM and n + 2 <= self.N_TABLE_DIM
        self.star_r[0], self.star_m[0] = ZERO, 0.0
        self.star_r[1:n + 1], self.star_m[1:n + 1] = r, m_star_units
        self.star_r[n + 1], self.star_m[n + 1] = SF_INFINITY, 0.0
        for name, val in (("r", r), ("vr", vr), ("vt", vt), ("m", m_star_units)):
            self.field(name)[1:n + 1] = val
        self.field("r")[0] = ZERO
        self.field("r")[n + 1] = SF_INFINITY
        self.field("id")[1:n + 1] = np.arange(1, n + 1) if ids is None else ids
        if binind is not None:
            self.field("binind")[1:n + 1] = binind
            self.N_b_local = self.clus.N_BINARY = int((np.asarray(binind) > 0).sum())
            self.bin_slots_in_use = int(np.max(binind)) if len(binind) else 0

    @classmethod
    def from_input_hdf5(cls, path, **kw):
        """A cluster from one of the reference's initial-condition files (``input.hdf5``, written by COSMIC; reference
        src/libs/fitslib.c:209-448, src/cmc/cmc_fits.c:100-140): objects 1 .. NOBJ in radius order, star_m = obj_m * N_STAR, binaries in
        binary[1 .. NBINARY] with the fields the file carries (ids, masses, radii, a, e), RTID as orbit_r."""
        from . import hdf5_input
        d = hdf5_input.read_input(path)
        n, nb = d["NOBJ"], d["NBINARY"]
        o, b = d["obj"], d["bin"]
        c = cls(n, n_bin_dim=kw.pop("n_bin_dim", nb + 16), orbit_r=kw.pop("orbit_r", d["RTID"]), **kw)
        c.load(o["r"][1:n + 1], o["vr"][1:n + 1], o["vt"][1:n + 1], o["m"][1:n + 1] * n, binind=o["binind"][1:n + 1].astype(np.int64),
               ids=o["id"][1:n + 1].astype(np.int64))
        for dst, src in (("id1", "id1"), ("id2", "id2")):
            c.bfield(dst)[1:nb + 1] = b[src][1:nb + 1].astype(np.int64)
        for dst, src, scale in (("m1", "m1", float(n)), ("m2", "m2", float(n)), ("rad1", "Reff1", 1.0), ("rad2", "Reff2", 1.0), ("a", "a", 1.0), ("e", "e", 1.0)):
            c.bfield(dst)[1:nb + 1] = b[src][1:nb + 1] * scale
        c.bfield("inuse")[1:nb + 1] = 1
        c.input = {k: d[k] for k in ("NOBJ", "NBINARY", "Z", "RVIR", "RTID", "MCLUS")}
        return c

    # ---- the reference's four entry points, host-buffer form ------------------------------------------------------
    def calc_potential_new(self):
        """calc_potential_new -> potential_calculate (reference src/cmc/cmc_utils.c:2006-2012, 675-736)."""
        nmax = C.c_long(self.clus.N_MAX)
        cm, cmn, Mt, Rt = C.c_double(self.cenma.m), C.c_double(self.cenma.m_new), C.c_double(), C.c_double()
        _lib.check(self.L.cmcb200_calc_potential_new(self.ctx, self.star_r.ctypes.data, self.star_m.ctypes.data, self.star_phi.ctypes.data,
                                                     C.byref(nmax), self.clus.N_STAR, C.byref(cm), C.byref(cmn), self.orbit_r,
                                                     C.byref(Mt), C.byref(Rt)), "calc_potential_new")
        self.clus.N_MAX = nmax.value
        self.cenma.m, self.cenma.m_new = cm.value, cmn.value
        self.Mtotal, self.Rtidal = Mt.value, Rt.value
        S, E = np.zeros(1, np.int32), np.zeros(1, np.int32)
        self.L.cmcb200_find_limits(self.clus.N_MAX, self.MIN_CHUNK_SIZE, 1, S.ctypes.data, E.ctypes.data)  # mpiFindIndicesCustom, :2011
        self.mpiBegin, self.mpiEnd = int(S[0]), int(E[0])
        return self.clus.N_MAX

    def qsorts_new(self):
        """qsorts_new -> sample_sort (reference src/cmc/cmc_utils.c:2232-2254) fused with post_sort_comm's r/m extraction."""
        n = C.c_long(self.clus.N_MAX_NEW)
        nb = C.c_long(self._n_bin_in_use())
        _lib.check(self.L.cmcb200_qsorts_new(self.ctx, self.star[1:].ctypes.data, C.byref(n), self.binary[1:].ctypes.data, self.N_BIN_DIM,
                                             C.byref(nb), self.star_r.ctypes.data, self.star_m.ctypes.data), "qsorts_new")
        self.clus.N_MAX = self.clus.N_MAX_NEW = n.value
        self.N_b_local = self.bin_slots_in_use = nb.value      # packed 1 .. nb after the sort
        self.Start[0], self.End[0] = 1, n.value    # findLimits inside sample_sort (cmc_sort.c:421)

    def post_sort_comm(self):
        """post_sort_comm (reference src/cmc/cmc_utils.c:2259-2288): on one rank star_r/star_m were already written by qsorts_new;
        the cenma reductions are identities."""
        return None

    def _n_bin_in_use(self):
        """highest binary[] slot a star points at.  The reference carries this as global state (N_b_local / N_BIN_DIM_OPT loops in
        src/cmc/cmc_sort.c:512-587); here it is known after load(binind=...) and after every sort, and re-derived from star[].binind
        (a strided scan of the host array) only when the caller filled the field by hand."""
        if self.bin_slots_in_use is None:
            b = self.field("binind")[1:self.clus.N_MAX_NEW + 1]
            self.bin_slots_in_use = int(b.max()) if b.size and b.max() > 0 else 0
        return self.bin_slots_in_use

    def orbit_params(self, draws=None, draw_off=None, seed=0, fuse_energy=False):
        n_old = int(self.End[0] - self.Start[0] + 1)
        n_new_stars = max(0, self.clus.N_MAX_NEW - n_old)
        p = _lib.OrbitParams(self.clus.N_MAX_NEW, n_old, int(self.Start[0]), self.clus.N_MAX, self.clus.N_STAR,
                             self.clus.N_MAX + 2 + n_new_stars, self.cenma.m, self.Rtidal, None, None, seed, self.tcount, int(fuse_energy))
        if draws is not None:
            self._draws = np.ascontiguousarray(draws, np.float64)
            self._draw_off = np.ascontiguousarray(draw_off, np.int64)
            p.draws, p.draw_off = self._draws.ctypes.data, self._draw_off.ctypes.data
        return p

    def new_orbits_calculate(self, draws=None, draw_off=None, seed=0, want_status=True):
        """new_orbits_calculate -> get_positions (reference src/cmc/cmc_utils.c:2194-2211).  Returns the per-star status codes;
        the host-side side effects of removed stars (remove_star, destroy_obj) are the caller's, in star order."""
        p = self.orbit_params(draws, draw_off, seed)
        n = self.clus.N_MAX_NEW
        res = _lib.OrbitResult()
        status = np.zeros(n, np.int32) if want_status else None
        kmin = np.zeros(n, np.int64) if want_status else None
        kmax = np.zeros(n, np.int64) if want_status else None
        _lib.check(self.L.cmcb200_new_orbits_calculate(self.ctx, self.star[1:].ctypes.data, self.star_r.ctypes.data, self.star_m.ctypes.data,
                                                       self.star_phi.ctypes.data, C.byref(p),
                                                       status.ctypes.data if want_status else None, kmin.ctypes.data if want_status else None,
                                                       kmax.ctypes.data if want_status else None, C.byref(res)), "new_orbits_calculate")
        self.max_r = res.max_r                       # MPI_Allreduce(max_r, MAX) is the identity on one rank
        self.last_orbit = res
        self.last_status, self.last_kmin, self.last_kmax = status, kmin, kmax
        return status

    def compute_energy(self):
        """E, J as ComputeEnergy does (reference src/cmc/cmc_utils.c:612-616), host arrays."""
        n = self.clus.N_MAX
        vr, vt = self.field("vr")[1:n + 1], self.field("vt")[1:n + 1]
        self.field("E")[1:n + 1] = self.star_phi[1:n + 1] + 0.5 * (vr * vr + vt * vt)
        self.field("J")[1:n + 1] = self.star_r[1:n + 1] * vt

    # ---- resident form ----------------------------------------------------------------------------------------------
    def upload(self):
        n, nb = self.clus.N_MAX_NEW, self._n_bin_in_use()
        # the tables hold the sorted stars [0 .. N_MAX + 1] and, behind them, the stars created this step (reference cmc_utils.c:2392-2395)
        n_new_stars = max(0, n - int(self.End[self.myid] - self.Start[self.myid] + 1))
        _lib.check(self.L.cmcb200_upload(self.ctx, self.star[1:].ctypes.data, n, self.binary[1:].ctypes.data, nb, self.star_r.ctypes.data,
                                         self.star_m.ctypes.data, self.star_phi.ctypes.data,
                                         min(self.N_TABLE_DIM, self.clus.N_MAX + 2 + n_new_stars)), "upload")
        self.N_b_local = nb

    def download(self):
        n = self.clus.N_MAX_NEW
        _lib.check(self.L.cmcb200_download(self.ctx, self.star[1:].ctypes.data, n, self.binary[1:].ctypes.data, self.N_b_local,
                                           self.star_r.ctypes.data, self.star_m.ctypes.data, self.star_phi.ctypes.data,
                                           min(self.N_TABLE_DIM, self.clus.N_MAX + 2)), "download")

    ALL_WORDS = (1 << (STAR_DTYPE_SIZE // 8)) - 1

    @staticmethod
    def dirty_mask(*fields_or_ranges):
        """bit mask over the 8-byte words of star_t for cmcb200_resident_begin: field names (layout.py) or (first byte, last byte + 1) ranges"""
        from .layout import STAR_OFFSETS
        m = 0
        for f in fields_or_ranges:
            lo, hi = (STAR_OFFSETS[f], STAR_OFFSETS[f] + 8) if isinstance(f, str) else f
            for w in range(lo // 8, (hi + 7) // 8):
                m |= 1 << w
        return m

    def resident_begin(self, dirty_words=None):
        """Start the timestep's upload of the host arrays (cmcb200_resident_begin): returns at once, the dev_* calls that follow are ordered
        behind it on the device and the orbit step runs chunk by chunk under it.  dirty_words: see include/cmcb200.h (None = everything)."""
        n, nb = self.clus.N_MAX_NEW, self._n_bin_in_use()
        n_new_stars = max(0, n - int(self.End[self.myid] - self.Start[self.myid] + 1))
        _lib.check(self.L.cmcb200_resident_begin(self.ctx, self.star[1:].ctypes.data, n, self.binary[1:].ctypes.data, nb, self.star_r.ctypes.data,
                                                 self.star_m.ctypes.data, self.star_phi.ctypes.data, min(self.N_TABLE_DIM, self.clus.N_MAX + 2 + n_new_stars),
                                                 self.ALL_WORDS if dirty_words is None else dirty_words), "resident_begin")
        self.N_b_local = nb

    def resident_end(self):
        """Download records, binaries and tables and wait: the host arrays are coherent with the device again (cmcb200_resident_end)."""
        _lib.check(self.L.cmcb200_resident_end(self.ctx, self.star[1:].ctypes.data, self.clus.N_MAX_NEW, self.binary[1:].ctypes.data, self.N_b_local,
                                               self.N_BIN_DIM, self.star_r.ctypes.data, self.star_m.ctypes.data, self.star_phi.ctypes.data,
                                               min(self.N_TABLE_DIM, self.clus.N_MAX + 2)), "resident_end")

    def host_step(self, seed=0, draws=None, draw_off=None, dirty_words=None):
        """One pass of the hot path on HOST arrays with the records crossing the link once each way: upload (asynchronous) -> orbit step
        behind the arriving chunks -> r <- rnew + sort -> potential -> download.  What the resident mode of the shim does between
        new_orbits_calculate and calc_potential_new; E and J are the caller's, as in the reference."""
        self.resident_begin(dirty_words)
        res = self.dev_new_orbits(draws, draw_off, seed)
        self.dev_transfer_and_sort()
        self.dev_calc_potential_new()
        self.resident_end()
        self.tcount += 1
        return res

    def dev_new_orbits(self, draws=None, draw_off=None, seed=0, want_status=False, fuse_energy=False):
        p = self.orbit_params(draws, draw_off, seed, fuse_energy)
        res = _lib.OrbitResult()
        _lib.check(self.L.cmcb200_dev_new_orbits(self.ctx, C.byref(p), int(want_status), C.byref(res)), "dev_new_orbits")
        self.max_r, self.last_orbit = res.max_r, res
        return res

    def dev_read_status(self):
        n = self.clus.N_MAX_NEW
        status, kmin, kmax = np.zeros(n, np.int32), np.zeros(n, np.int64), np.zeros(n, np.int64)
        _lib.check(self.L.cmcb200_dev_read_status(self.ctx, status.ctypes.data, kmin.ctypes.data, kmax.ctypes.data, n), "dev_read_status")
        return status, kmin, kmax

    def dev_transfer_new_positions(self):
        n_old = int(self.End[0] - self.Start[0] + 1)
        _lib.check(self.L.cmcb200_dev_transfer_new_positions(self.ctx, self.clus.N_MAX_NEW, n_old, int(self.Start[0]), self.clus.N_MAX), "dev_transfer")

    def dev_energy_conservation2(self):
        """energy_conservation2 (reference src/cmc/cmc_utils.c:2217-2226) on the resident arrays: Uoldrnew of every local star"""
        p = self.orbit_params()
        _lib.check(self.L.cmcb200_dev_energy_conservation2(self.ctx, C.byref(p)), "dev_energy_conservation2")

    def dev_compute_intermediate_energy(self):
        """ComputeIntermediateEnergy (reference src/cmc/cmc_utils.c:368-391) on the resident arrays: EI, then r <- rnew etc."""
        p = self.orbit_params()
        _lib.check(self.L.cmcb200_dev_compute_intermediate_energy(self.ctx, C.byref(p)), "dev_compute_intermediate_energy")

    def dev_potential(self, r):
        """potential() (reference src/cmc/cmc_utils.c:112-165) at the radii r on the resident tables"""
        r = np.ascontiguousarray(np.atleast_1d(r), np.float64)
        out = np.zeros(len(r))
        p = self.orbit_params()
        _lib.check(self.L.cmcb200_dev_potential(self.ctx, C.byref(p), r.ctypes.data, len(r), out.ctypes.data), "dev_potential")
        return out

    def dev_tidally_strip_stars(self, tidal_treatment=1, GAMMA=0.11, cap=8192):
        """tidally_strip_stars (reference src/cmc/cmc_evolution_thr.c:443-645) on the resident records.  The criterion runs on the
        device (one pass per call of cmcb200_dev_tidally_strip_select); this is the reference's host loop around it: Rtidal from the mass
        lost so far, the per-star sums in star order from the handful of stripped records the device hands back, repeated while a pass
        strips anything.  On several ranks (ShardedCluster) the stripped mass is summed over the ranks every pass and the escaper sums at the
        end, as in the reference.  Returns the 1-based local indices of all stripped stars."""
        import math
        N = float(self.clus.N_STAR)
        def buffers(cap):
            return (np.zeros(cap, np.int64), np.zeros((cap, STAR_DTYPE_SIZE), np.uint8), np.zeros((cap, BINARY_DTYPE_SIZE), np.uint8), np.zeros(cap))
        if getattr(self, "_strip_buf", None) is None or len(self._strip_buf[0]) < cap:      # hand-back buffers, kept between calls
            self._strip_buf = buffers(cap)
        cnt = C.c_long(0)
        self.Etidal, stripped = 0.0, []
        while True:
            self.Rtidal = self.orbit_r * math.pow(self.Mtotal - (self.TidalMassLoss - self.OldTidalMassLoss), 1.0 / 3.0)   # :461-462
            phi_rtidal = float(self.dev_potential(self.Rtidal)[0])
            thr = self.Rtidal if tidal_treatment == 0 else (1.5 - 3.0 * math.pow(math.log(GAMMA * N) / N, 0.25)) * phi_rtidal   # :479 / :553-555
            p = self.orbit_params()
            while True:
                idx, recs, bins, mass = self._strip_buf
                rc = self.L.cmcb200_dev_tidally_strip_select(self.ctx, C.byref(p), tidal_treatment, thr, idx.ctypes.data, len(idx), C.byref(cnt),
                                                             recs.ctypes.data, bins.ctypes.data, mass.ctypes.data)
                if rc == -3 and cnt.value > len(idx):      # more hits than the list holds: nothing was marked, come back with a longer one
                    self._strip_buf = buffers(int(cnt.value) + 1024)
                    continue
                _lib.check(rc, "dev_tidally_strip_select")
                break
            n = cnt.value
            D = 0.0
            if n:
                E, J, Eint, binind = (star_field(recs[:n], k) for k in ("E", "J", "Eint", "binind"))
                bf = {k: binary_field(bins[:n], k) for k in ("m1", "m2", "a", "Eint1", "Eint2")}
                for k in range(n):                                    # :484-496, in star order
                    m = float(mass[k])                                # star_m[g] from the resident table (:471)
                    self.Eescaped += float(E[k]) * m / N
                    self.Jescaped += float(J[k]) * m / N
                    if binind[k] == 0:
                        self.Eintescaped += float(Eint[k])
                    else:
                        self.Ebescaped += -(float(bf["m1"][k]) / N) * (float(bf["m2"][k]) / N) / (2.0 * float(bf["a"][k]))
                        self.Eintescaped += float(bf["Eint1"][k]) + float(bf["Eint2"][k])
                    D += m / N
                    self.Etidal += float(E[k]) * m / N
                stripped.extend(int(i) for i in idx[:n])
            D = self._allreduce_sum([D])[0]                           # :612-616 (one rank: the identity)
            self.TidalMassLoss += D                                   # :617
            self.DTidalMassLoss = D
            if not D > 0:                                             # :624
                break
        # :626-644: the ranks' escaper sums become their totals
        self.Eescaped, self.Jescaped, self.Eintescaped, self.Ebescaped, self.Etidal = self._allreduce_sum(
            [self.Eescaped, self.Jescaped, self.Eintescaped, self.Ebescaped, self.Etidal])
        return stripped

    def _allreduce_sum(self, values):
        """MPI_Allreduce(MPI_SUM) over the ranks of the run; one rank here (ShardedCluster overrides it)"""
        return list(values)

    def dev_energy_new_radii(self):
        """energy_conservation2 and the EI loop of ComputeIntermediateEnergy from one potential(rnew) lookup per star; the transfer is left to
        dev_transfer_and_sort (reference src/cmc/cmc.c:317-335 with TIDALLY_STRIP_STARS == 0)"""
        p = self.orbit_params()
        _lib.check(self.L.cmcb200_dev_energy_new_radii(self.ctx, C.byref(p)), "dev_energy_new_radii")

    def dev_timestep(self, seed=0, draws=None, draw_off=None, tidally_strip=None, GAMMA=0.11, dynamics_dt=None, want_sigma=False):
        """One timestep of the reference's main loop around the hot path (src/cmc/cmc.c:302-353, no interacting stars) on the resident arrays:
        energy_conservation1 -> new_orbits_calculate -> energy_conservation2 -> [tidally_strip_stars] -> ComputeIntermediateEnergy ->
        qsorts_new -> calc_potential_new -> energy_conservation3 -> ComputeEnergy.  tidally_strip: None (TIDALLY_STRIP_STARS = 0) or the
        TIDAL_TREATMENT (0 / 1); TidalMassLoss / OldTidalMassLoss are the caller's to maintain across the orbit step
        (reference cmc_utils.c:2197-2210).  Without stripping the passes that read the same data are merged (one lookup for energy_conservation2 +
        ComputeIntermediateEnergy, the transfer inside the sort's gather); bit-identical to the calls made one by one either way.
        dynamics_dt: a time step Dt puts dynamics_apply (src/cmc/cmc.c:286; calc_sigma_r + the pair loop) in front, so that
        nothing of the timestep needs the records on the host when no pair is a strong encounter; the encounter pairs are returned in
        ``self.last_dynamics`` for the caller's host code (cmcb200_dev_get / _put)."""
        if dynamics_dt is not None:
            if want_sigma:
                self.sigma_array = self.dev_calc_sigma_r()
            else:                                  # sigma_array stays on the device (cmcb200_dev_sigma_arrays): no host reader without encounters
                prm = _lib.SigmaParams(self.DYNAMICS["AVEKERNEL"], int(self.End[self.myid] - self.Start[self.myid] + 1), int(self.Start[self.myid]),
                                       self.clus.N_MAX, self.clus.N_STAR, 0, 0, None, None, None, None)
                _lib.check(self.L.cmcb200_dev_calc_sigma_r(self.ctx, C.byref(prm), None, None, None), "dev_calc_sigma_r")
            self.last_dynamics = self.dev_dynamics_apply(dynamics_dt, seed=seed)
        self.dev_energy_conservation1()
        res = self.dev_new_orbits(draws, draw_off, seed)
        if tidally_strip is None:
            self.dev_energy_new_radii()
            self.dev_transfer_and_sort()
        else:
            self.dev_energy_conservation2()
            self.dev_tidally_strip_stars(tidally_strip, GAMMA)
            self.dev_compute_intermediate_energy()
            self.dev_qsorts_new()
        self.dev_calc_potential_new()
        self.dev_energy_conservation3()
        self.dev_compute_energy()
        self.tcount += 1
        return res

    def dev_energy_conservation1(self):
        """energy_conservation1 (reference src/cmc/cmc_utils.c:2153-2174) on the resident arrays: vrold, vtold, Uoldrold of every local star"""
        p = self.orbit_params()
        _lib.check(self.L.cmcb200_dev_energy_conservation1(self.ctx, C.byref(p)), "dev_energy_conservation1")

    def dev_energy_conservation3(self):
        """energy_conservation3 -> set_velocities3 (reference src/cmc/cmc_utils.c:254-363) on the resident arrays, one rank: the first loop;
        the received Eexcess of the second loop is 0 on rank 0, so it changes nothing and Eoops += -0 (:322, :351).  Returns the leftover
        excess of the first loop (a diagnostic: the reference drops it)."""
        p, left = self.orbit_params(), C.c_double()
        _lib.check(self.L.cmcb200_dev_set_velocities3(self.ctx, C.byref(p), C.byref(left)), "dev_set_velocities3")
        # Eoops += -Eexcess * madhoc with Eexcess = the received 0.0: unchanged
        return left.value

    def dev_set_velocities3_pass(self, Eexcess):
        """one pass of set_velocities3's second loop (:323-343) from the Eexcess received from the previous rank; returns what is left"""
        p, left = self.orbit_params(), C.c_double()
        _lib.check(self.L.cmcb200_dev_set_velocities3_pass(self.ctx, C.byref(p), Eexcess, C.byref(left)), "dev_set_velocities3_pass")
        return left.value

    # the reference's switches dynamics_apply reads (include/cmc/cmc_vars.h), with the defaults of examples/PlummerSphere.ini
    DYNAMICS = dict(RELAXATION=1, SS_COLLISION=0, BINSINGLE=0, BINBIN=0, TIDAL_CAPTURE=0, TC_POLYTROPE=0, BH_CAPTURE=0, BH_LOSS_CONE=0,
                    AVEKERNEL=20, THREEBODYBINARIES=0, GAMMA=0.11, COLL_FACTOR=1.0, TC_FACTOR=1.0, WD_TC=0.0,
                    units_mstar=1.989e33, units_l=3.0e18, units_t=1.0e13)

    def dev_dynamics_apply(self, dt, draws=None, seed=0, **switches):
        """dynamics_apply (reference src/cmc/cmc_dynamics.c:17-542) on the resident arrays: every pair of neighbouring stars is either relaxed in
        place (new vr, vt, E, J) or, when its encounter draw falls below P_enc, listed for the host code that runs the reference's
        binint_do / sscollision_do.  switches: the reference's globals of that name (``DYNAMICS``).  Returns the result block with the
        scattering-angle statistics and ``encounters`` = dict(k, kp, type, rperi, W, rcm) in pair order."""
        cfg = dict(self.DYNAMICS)
        cfg.update(switches)
        n = int(self.End[self.myid] - self.Start[self.myid] + 1)    # mpiEnd - mpiBegin + 1
        p = _lib.DynamicsParams(n, int(self.Start[self.myid]), self.clus.N_MAX, self.clus.N_STAR, dt, cfg["RELAXATION"], cfg["SS_COLLISION"],
                                cfg["BINSINGLE"], cfg["BINBIN"], cfg["TIDAL_CAPTURE"], cfg["TC_POLYTROPE"], cfg["BH_CAPTURE"], cfg["BH_LOSS_CONE"],
                                cfg["AVEKERNEL"], cfg["THREEBODYBINARIES"], self.cenma.m, cfg["GAMMA"], cfg["COLL_FACTOR"], cfg["TC_FACTOR"],
                                cfg["WD_TC"], cfg["units_mstar"], cfg["units_l"], cfg["units_t"], None, seed, self.tcount)
        if draws is not None:
            self._dyn_draws = np.ascontiguousarray(draws, np.float64)
            assert len(self._dyn_draws) >= 3 * (n // 2)
            p.draws = self._dyn_draws.ctypes.data
        res = _lib.DynamicsResult()
        _lib.check(self.L.cmcb200_dev_dynamics_apply(self.ctx, C.byref(p), C.byref(res)), "dev_dynamics_apply")
        ne = res.n_encounters
        enc = dict(k=np.zeros(ne, np.int64), kp=np.zeros(ne, np.int64), type=np.zeros(ne, np.int32), rperi=np.zeros(ne), W=np.zeros(ne), rcm=np.zeros(ne))
        cnt = C.c_long()
        _lib.check(self.L.cmcb200_dev_dynamics_encounters(self.ctx, ne, *[enc[k].ctypes.data for k in ("k", "kp", "type", "rperi", "W", "rcm")],
                                                          C.byref(cnt)), "dev_dynamics_encounters")
        res.encounters = enc
        return res

    def dev_simul_relax(self, THETASEMAX=1.0):
        """simul_relax_new (reference src/cmc/cmc_dynamics_helper.c:2526-2580) on the resident arrays: the relaxation time step, bit for bit"""
        n, out = int(self.End[self.myid] - self.Start[self.myid] + 1), C.c_double()
        _lib.check(self.L.cmcb200_dev_simul_relax(self.ctx, n, int(self.Start[self.myid]), self.clus.N_MAX, self.clus.N_STAR, self.DYNAMICS["AVEKERNEL"],
                                                  THETASEMAX, C.byref(out)), "dev_simul_relax")
        return out.value

    def calc_timestep(self, THETASEMAX=1.0):
        """calc_timestep -> GetTimeStep (reference src/cmc/cmc_utils.c:2132-2148, src/cmc/cmc_evolution_thr.c:317-437) for a run whose only
        physics is two-body relaxation (no collisions, binaries, stellar evolution: every other candidate of GetTimeStep is +inf):
        Dt = simul_relax_new() on the resident arrays, then calc_timestep's tidal-mass-loss rules; advances TotalTime."""
        Dt = self._allreduce_min(self.dev_simul_relax(THETASEMAX))
        if Dt == float("inf"):
            Dt = 0.001                                                        # cmc_evolution_thr.c:428-430
        prev = getattr(self, "Prev_Dt", 0.0)
        lost_old = getattr(self, "TidalMassLoss_old", 0.0) - self.OldTidalMassLoss
        if lost_old > 0.01:                                                   # cmc_utils.c:2138-2141
            Dt = prev * 0.8
        elif Dt > 1.1 * prev and prev > 0 and (self.TidalMassLoss - self.OldTidalMassLoss) > 0.02:
            Dt = prev * 1.1
        self.Dt = Dt
        self.TotalTime = getattr(self, "TotalTime", 0.0) + Dt
        return Dt

    def _allreduce_min(self, x):
        return x

    def evolve(self, t_end, max_steps=10**9, seed=0, THETASEMAX=1.0, tidally_strip=0, GAMMA=0.1, mass_pc=None, record_every=1, min_dt=0.0):
        """The reference's main loop (src/cmc/cmc.c:240-407) for a relaxation-only run (examples/PlummerSphere.ini), every N-sized operation on
        the resident arrays: calc_timestep (simul_relax_new) -> dynamics_apply (calc_sigma_r, two-body relaxation of every pair) ->
        energy_conservation1 -> new_orbits_calculate -> energy_conservation2 -> tidally_strip_stars -> ComputeIntermediateEnergy -> qsorts_new
        -> calc_potential_new -> energy_conservation3 -> comp_mass_percent -> ComputeEnergy.  The 424-byte records never visit the host.
        Stops at t_end, after max_steps, or when the time step falls below min_dt (the runaway of core collapse, where the reference's
        STOPATCORECOLLAPSE ends the run).  Returns the history [(TotalTime, Dt, N_MAX, Lagrange radii)] recorded every `record_every` steps."""
        hist, steps = [], 0
        pc = None if mass_pc is None else np.ascontiguousarray(mass_pc, np.float64)
        while getattr(self, "TotalTime", 0.0) < t_end and steps < max_steps:
            # cmc.c:245-258: the per-step accumulators
            self.TidalMassLoss_old, self.TidalMassLoss = self.TidalMassLoss, 0.0
            Dt = self.calc_timestep(THETASEMAX)
            self.OldTidalMassLoss = self.TidalMassLoss_old                    # new_orbits_calculate, cmc_utils.c:2197
            self.dev_timestep(seed=seed, tidally_strip=tidally_strip, GAMMA=GAMMA, dynamics_dt=Dt)
            self.TidalMassLoss += self.TidalMassLoss_old                      # cmc_utils.c:2210
            self.Prev_Dt = Dt
            steps += 1
            if Dt < min_dt:
                break
            if steps % record_every == 0:
                lag = None
                if pc is not None:
                    self.dev_comp_mass_percent(pc)
                    lag = self.mass_r.copy()
                hist.append((self.TotalTime, Dt, self.clus.N_MAX, lag))
        return hist

    def dev_calc_sigma_r(self, p=None, mave=False, ghosts=None, force_sequential=False):
        """calc_sigma_r (reference src/cmc/cmc_dynamics_helper.c:2821-2975) on the resident arrays, bit for bit.  ghosts (P ranks only):
        dict prev_vr, prev_vt, next_vr, next_vt of p values each (None on the first / last rank).  Returns (sigma_array.r or the mean mass,
        sigma_array.sigma), each [n_local + 1] with entry 0 unused, and leaves ``self.sigma_stats``."""
        p = int(p or self.DYNAMICS["AVEKERNEL"])
        n = int(self.End[self.myid] - self.Start[self.myid] + 1)
        prm = _lib.SigmaParams(p, n, int(self.Start[self.myid]), self.clus.N_MAX, self.clus.N_STAR, int(mave), int(force_sequential), None, None, None, None)
        keep = []
        for k in ("prev_vr", "prev_vt", "next_vr", "next_vt"):
            v = (ghosts or {}).get(k)
            if v is not None:
                v = np.ascontiguousarray(v, np.float64)
                assert len(v) == p
                keep.append(v)
                setattr(prm, "ghost_" + k, v.ctypes.data)
        a, sg, st = np.zeros(n + 1), np.zeros(n + 1), np.zeros(6, np.int64)
        _lib.check(self.L.cmcb200_dev_calc_sigma_r(self.ctx, C.byref(prm), a.ctypes.data, sg.ctypes.data, st.ctypes.data), "dev_calc_sigma_r")
        self.sigma_stats = dict(zip(("hard_segments", "wide_segments", "hard_tiles", "broken", "rerun", "tiles"), st.tolist()))
        return a, sg

    def save_restart_file(self, path, rng_state=b"", restart_struct=b"", clus_struct=b"", snapshot_window_counters=()):
        """save_restart_file (reference src/cmc/cmc_io.c:3829-3884) with the records taken from the device: the three header blocks are the
        caller's bytes (curr_st, restart_struct_t, clus_struct_t), star[] / binary[] are written whole -- their live part from HBM."""
        cnt = np.ascontiguousarray(snapshot_window_counters, np.int32)
        blobs = [np.frombuffer(bytes(b), np.uint8) if len(b) else np.zeros(0, np.uint8) for b in (rng_state, restart_struct, clus_struct)]
        _lib.check(self.L.cmcb200_restart_write(self.ctx, str(path).encode(), blobs[0].ctypes.data, len(blobs[0]), blobs[1].ctypes.data, len(blobs[1]),
                                                blobs[2].ctypes.data, len(blobs[2]), self.star.ctypes.data, self.N_STAR_DIM, self.clus.N_MAX_NEW,
                                                self.binary.ctypes.data, self.N_BIN_DIM, self.N_b_local, cnt.ctypes.data, len(cnt)), "restart_write")

    def dev_snapshot_rows(self, units_m, units_l, sigma=None, avekernel=None, bh_only=False, nan_value=None):
        """write_snapshot's rows (reference src/cmc/cmc_io.c:3176-3327; the `Snapshot` struct, include/cmc/cmc.h:136-200) built on the device
        from the resident arrays: a structured array (layout.SNAPSHOT_DTYPE), in star order -- what the reference hands to H5TBmake_table.
        sigma: sigma_array.sigma ([n_local + 1]) or None for the resident result of dev_calc_sigma_r; nan_value: the caller's `0.0 / 0.0`
        (default: the x86 default NaN, sign bit set, which is what the compiled reference stores)."""
        from .layout import SNAPSHOT_DTYPE
        n = int(self.clus.N_MAX_NEW)
        nan = np.array([0xFFF8000000000000], np.uint64).view(np.float64)[0] if nan_value is None else nan_value
        prm = _lib.SnapshotParams(n, int(self.End[self.myid] - self.Start[self.myid] + 1), int(self.Start[self.myid]), self.clus.N_MAX, self.clus.N_STAR,
                                  units_m, units_l, 1.0 / self.clus.N_STAR, int(avekernel or self.DYNAMICS["AVEKERNEL"]), nan, int(bool(bh_only)))
        rows = np.zeros(max(n, 1), SNAPSHOT_DTYPE)
        sg = None if sigma is None else np.ascontiguousarray(sigma, np.float64)
        assert sg is None or len(sg) >= n + 1
        got = C.c_long(0)
        _lib.check(self.L.cmcb200_dev_snapshot_rows(self.ctx, C.byref(prm), None if sg is None else sg.ctypes.data, rows.ctypes.data, len(rows), C.byref(got)), "dev_snapshot_rows")
        return rows[:got.value]

    def dev_qsorts_new(self):
        n, nb = C.c_long(self.clus.N_MAX_NEW), C.c_long(0)
        _lib.check(self.L.cmcb200_dev_qsorts_new(self.ctx, C.byref(n), max(self.N_b_local, 0), C.byref(nb)), "dev_qsorts_new")
        self.clus.N_MAX = self.clus.N_MAX_NEW = n.value
        self.N_b_local = self.bin_slots_in_use = nb.value
        self.Start[0], self.End[0] = 1, n.value

    def dev_transfer_and_sort(self):
        """dev_transfer_new_positions + dev_qsorts_new in one pass over the records"""
        n, nb = C.c_long(self.clus.N_MAX_NEW), C.c_long(0)
        n_old = int(self.End[0] - self.Start[0] + 1)
        _lib.check(self.L.cmcb200_dev_transfer_and_sort(self.ctx, C.byref(n), n_old, int(self.Start[0]), self.clus.N_MAX, max(self.N_b_local, 0),
                                                        C.byref(nb)), "dev_transfer_and_sort")
        self.clus.N_MAX = self.clus.N_MAX_NEW = n.value
        self.N_b_local = self.bin_slots_in_use = nb.value
        self.Start[0], self.End[0] = 1, n.value

    def dev_calc_potential_new(self):
        nmax = C.c_long(self.clus.N_MAX)
        cm, cmn, Mt, Rt = C.c_double(self.cenma.m), C.c_double(self.cenma.m_new), C.c_double(), C.c_double()
        _lib.check(self.L.cmcb200_dev_calc_potential_new(self.ctx, C.byref(nmax), self.clus.N_STAR, C.byref(cm), C.byref(cmn), self.orbit_r,
                                                         C.byref(Mt), C.byref(Rt)), "dev_calc_potential_new")
        self.clus.N_MAX = nmax.value
        self.cenma.m, self.cenma.m_new = cm.value, cmn.value
        self.Mtotal, self.Rtidal = Mt.value, Rt.value
        self.mpiBegin, self.mpiEnd = 1, self.clus.N_MAX

    def dev_compute_energy(self):
        _lib.check(self.L.cmcb200_dev_compute_energy(self.ctx, self.clus.N_MAX, 1), "dev_compute_energy")

    def dev_compute_energy_totals(self):
        """The second loop of ComputeEnergy (reference src/cmc/cmc_utils.c:619-655) on the resident arrays: Etotal.K / .P / .Eint / .Eb / .tot
        as a SimpleNamespace (one rank; cenma.E is the caller's)."""
        out = np.zeros(4)
        _lib.check(self.L.cmcb200_dev_compute_energy_totals(self.ctx, self.clus.N_MAX_NEW, int(self.Start[self.myid]), self.clus.N_MAX, self.clus.N_STAR,
                                                            self.cenma.m, out.ctypes.data), "dev_compute_energy_totals")
        K, P, Eint, Eb = (float(x) for x in self._allreduce_sum([out[0], 0.5 * out[1], out[2], out[3]]))
        tot = K + P + Eint + Eb
        if self.procs > 1:
            tot = self._allreduce_sum([float(out[0]) + 0.5 * float(out[1]) + float(out[2]) + float(out[3])])[0]   # buf_reduce[0] is reduced on its own (:640)
        tot += getattr(self.cenma, "E", 0.0) + self.Eescaped + self.Ebescaped + self.Eintescaped                    # :655
        self.Etotal = SimpleNamespace(tot=tot, K=K, P=P, Eint=Eint, Eb=Eb)
        return self.Etotal

    def dev_step(self, seed=0, draws=None, draw_off=None, fused=True, one_wait=True):
        """One pass of the hot path on the resident arrays: (E, J from the potential ->) orbit step -> r <- rnew -> sort -> potential.
        fused: the r <- rnew transfer rides on the sort's record gather and E, J are computed inside the orbit step; the unfused form
        runs them as passes of their own (same results, bit for bit).  one_wait: the fused form as ONE call that queues everything and waits
        once (cmcb200_dev_step) instead of three calls with a wait each."""
        if fused and one_wait and type(self).dev_transfer_and_sort is Cluster.dev_transfer_and_sort:
            p = self.orbit_params(draws, draw_off, seed, fuse_energy=True)
            res, out = _lib.OrbitResult(), _lib.StepResult()
            _lib.check(self.L.cmcb200_dev_step(self.ctx, C.byref(p), max(self.N_b_local, 0), self.cenma.m_new, self.orbit_r, C.byref(res), C.byref(out)), "dev_step")
            self.last_orbit, self.max_r = res, res.max_r
            self.clus.N_MAX_NEW = out.n_local
            self.N_b_local = self.bin_slots_in_use = out.n_bin
            self.Start[0], self.End[0] = 1, out.n_local
            self.clus.N_MAX = out.N_MAX
            self.cenma.m, self.cenma.m_new = out.cenma_m, out.cenma_m_new
            self.Mtotal, self.Rtidal = out.Mtotal, out.Rtidal
            self.mpiBegin, self.mpiEnd = 1, self.clus.N_MAX
        elif fused:
            res = self.dev_new_orbits(draws, draw_off, seed, fuse_energy=True)
            self.dev_transfer_and_sort()
            self.dev_calc_potential_new()
        else:
            self.dev_compute_energy()
            res = self.dev_new_orbits(draws, draw_off, seed)
            self.dev_transfer_new_positions()
            self.dev_qsorts_new()
            self.dev_calc_potential_new()
        self.tcount += 1
        return res

    def sync(self):
        _lib.check(self.L.cmcb200_sync(self.ctx), "sync")

    # ---- instrumentation ------------------------------------------------------------------------------------------------
    def profile(self, on=True):
        self.L.cmcb200_profile(self.ctx, int(on))

    def profile_read(self, reset=True):
        ms, n = np.zeros(_lib.N_STAGES), np.zeros(_lib.N_STAGES, np.int64)
        _lib.check(self.L.cmcb200_profile_read(self.ctx, ms.ctypes.data, n.ctypes.data, int(reset)), "profile_read")
        return {k: (float(a), int(b)) for k, a, b in zip(_lib.STAGE_NAMES, ms, n)}

    MASS_PC_FIELDS = ("mass_r", "ave_mass_r", "no_star_r", "densities_r", "ke_rad_r", "ke_tan_r", "v2_rad_r", "v2_tan_r")

    def dev_comp_mass_percent(self, mass_pc, bh_include=False, minimum_r=0.0, initial_total_mass=1.0):
        """comp_mass_percent (reference src/cmc/cmc_utils.c:872-985) on the resident arrays: the Lagrange radii mass_r[], the star counts,
        mean masses, densities and the cumulative kinetic sums at them, under the reference's names.  Entries the loop never reaches
        keep their previous values (-1 on first use)."""
        pc = np.ascontiguousarray(mass_pc, np.float64)
        n = len(pc)
        for f in self.MASS_PC_FIELDS:
            if getattr(self, f, None) is None or len(getattr(self, f)) != n:
                setattr(self, f, np.full(n, -1.0))
        owner = np.zeros(n, np.int32)
        p = _lib.MassPcParams(n, pc.ctypes.data, int(bh_include), minimum_r, initial_total_mass, self.Mtotal, self.cenma.m, self.clus.N_MAX,
                              self.clus.N_STAR, self.clus.N_MAX_NEW, int(self.Start[self.myid]))
        res = _lib.MassPcResult(*[getattr(self, f).ctypes.data for f in self.MASS_PC_FIELDS], owner.ctypes.data)
        _lib.check(self.L.cmcb200_dev_comp_mass_percent(self.ctx, C.byref(p), C.byref(res)), "dev_comp_mass_percent")
        self.mass_pc, self.mass_pc_owner, self.mass_pc_totals = pc, owner, np.array(list(res.totals))
        return int(res.filled)

    def dev_comp_multi_mass_percent(self, mass_bins, solar_mass_dyn, mass_pc):
        """comp_multi_mass_percent (reference src/cmc/cmc_utils.c:769-861) on the resident tables: multi_mass_r[NO_MASS_BINS][MASS_PC_COUNT], the
        Lagrange radii per mass bin (rows of empty bins keep their previous values, -1 on first use)"""
        bins, pc = np.ascontiguousarray(mass_bins, np.float64), np.ascontiguousarray(mass_pc, np.float64)
        if getattr(self, "multi_mass_r", None) is None or self.multi_mass_r.shape != (len(bins), len(pc)):
            self.multi_mass_r = np.full((len(bins), len(pc)), -1.0)
        _lib.check(self.L.cmcb200_dev_comp_multi_mass_percent(self.ctx, self.clus.N_MAX, bins.ctypes.data, len(bins), solar_mass_dyn, pc.ctypes.data, len(pc),
                                                              self.multi_mass_r.ctypes.data), "dev_comp_multi_mass_percent")
        return self.multi_mass_r

    def dev_clusdyn_calculate(self):
        """clusdyn_calculate (reference src/cmc/cmc_utils.c:1613-1624): clusdyn.rh, the half-mass radius, from the resident tables"""
        rh = C.c_double()
        _lib.check(self.L.cmcb200_dev_clusdyn_calculate(self.ctx, self.clus.N_MAX, self.clus.N_STAR, self.Mtotal, C.byref(rh)), "dev_clusdyn_calculate")
        self.clusdyn_rh = rh.value
        return rh.value

    def launch_count(self):
        return int(self.L.cmcb200_launch_count(self.ctx))

    def transfer_bytes(self, reset=False):
        """(host->device, device->host) bytes the host-buffer entry points have moved"""
        a, b = C.c_long(), C.c_long()
        _lib.check(self.L.cmcb200_transfer_bytes(self.ctx, C.byref(a), C.byref(b), int(reset)), "transfer_bytes")
        return a.value, b.value

    def potential_stats(self):
        ev, fb = C.c_long(), C.c_long()
        _lib.check(self.L.cmcb200_potential_stats(self.ctx, C.byref(ev), C.byref(fb)), "potential_stats")
        return ev.value, fb.value
