"""Radius-sharded operation: one process and one GPU per rank, the reference's own MPI partition (findLimits).

The reference runs this path as P MPI ranks, each owning the sorted stars Start[i]..End[i] and a full copy of star_r / star_m /
star_phi (docs/sphinx/source/parallel/index.rst; src/cmc/cmc_sort.c:375-791; src/cmc/cmc_utils.c:2259-2288).  :class:`ShardedCluster`
is that rank: the same attributes as :class:`cmc_cosmic_b200.cluster.Cluster` (``star`` holds THIS rank's records, the tables hold the
whole cluster, ``Start`` / ``End`` / ``myid`` / ``procs`` as in include/cmc/cmc_vars.h) and the same entry points.

Data moves between GPUs inside the library's own kernels (stores into peer memory mapped with CUDA IPC, over NVLink); this host layer
only allgathers a few integers per step and provides the barriers between the phases -- ``torch.distributed`` here, MPI in CMC
(INTEGRATION.md).  The planning arithmetic (:func:`plan_partition`, :func:`host_global_ranks`) is plain integer code, exercised on CPU
under the gloo backend by tests/test_sharded_host.py.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .cluster import Cluster
from .layout import SF_INFINITY, ZERO
from .partition import find_limits


def plan_partition(n_keep_all, blk):
    """New partition after a sort: findLimits over the stars all ranks kept (reference cmc_sort.c:416-421).  Returns Start, End (1-based)."""
    n_keep_all = np.asarray(n_keep_all, np.int64)
    return find_limits(int(n_keep_all.sum()), blk, len(n_keep_all))


def host_global_ranks(runs, g):
    """The rule shard_rank_kernel implements, restated with numpy for the host-side tests: global rank of every element of sorted run g among
    the concatenation of all runs -- ties go to the lower rank, then to the lower local index (a stable sort of the concatenated shards)."""
    mine = np.asarray(runs[g])
    rank = np.arange(len(mine), dtype=np.int64)
    for q, run in enumerate(runs):
        if q == g:
            continue
        rank += np.searchsorted(run, mine, side="right" if q < g else "left")
    return rank


def destinations(rank, Start):
    """destination rank and 0-based local position of a global rank under the partition Start (1-based)"""
    start0 = np.asarray(Start, np.int64) - 1
    d = np.searchsorted(start0, rank, side="right") - 1
    return d, rank - start0[d]


def carry_excess_between_ranks(dist, Eexcess, run_pass, device="cpu"):
    """The do-while loop of set_velocities3 (reference src/cmc/cmc_utils.c:313-347) around `run_pass`, one rank's view.
    Eexcess: this rank's leftover of the first loop.  Every round each rank hands its Eexcess to rank + 1 (MPI_Send / MPI_Recv in the
    reference; an allgather of one double here), rank 0 receives 0, `run_pass(received)` runs the second loop over the local stars and
    returns what is left, and the rounds go on while the sum of the leftovers over the ranks is > 0.  Returns (final Eexcess, rounds);
    the caller books Eoops += -Eexcess * madhoc (:351)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    rounds = 0
    while True:
        parts = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(world)]
        dist.all_gather(parts, torch.tensor([Eexcess], dtype=torch.float64, device=device))
        prev = float(parts[rank - 1]) if rank > 0 else 0.0
        Eexcess = run_pass(prev)
        rounds += 1
        total = torch.tensor([Eexcess], dtype=torch.float64, device=device)
        dist.all_reduce(total)
        if not float(total[0]) > 0:
            return Eexcess, rounds


class ShardedCluster(Cluster):
    """One rank of a radius-sharded cluster.  `dist` is an initialised torch.distributed process group (nccl); rank r drives cuda:`device`."""

    def __init__(self, n_star, dist, device, min_chunk=40, orbit_r=1.0e6, n_bin_dim=None, pinned=False, stream=None, slack=1.5):
        import torch
        self.dist = dist
        rank, world = dist.get_rank(), dist.get_world_size()
        # the library's kernels and torch's collectives must be ordered on ONE stream: the barriers between the phases are collectives
        self._tstream = stream if stream is not None else torch.cuda.Stream(device=device)
        S, E = find_limits(n_star, min_chunk, world)
        share = int((E - S + 1).max())
        # N_STAR_DIM_OPT = 1.5 (1 + N/P + ...) (reference src/cmc/cmc_io.c:2092-2099): room for the shard to breathe between load balances
        n_star_dim = int(slack * (share + 2)) + 4096
        self._n_table_dim = n_star + 4096
        super().__init__(n_star, n_star_dim=n_star_dim, n_bin_dim=n_bin_dim, device=device, min_chunk=min_chunk, orbit_r=orbit_r, pinned=pinned,
                         stream=self._tstream.cuda_stream, n_table_dim=self._n_table_dim)
        self.myid, self.procs = rank, world
        self.Start, self.End = S.copy(), E.copy()
        self.clus.N_MAX_NEW = int(E[rank] - S[rank] + 1)
        self._attach()

    # ---- set-up -------------------------------------------------------------------------------------------------------------
    def _attach(self):
        import torch
        nb = self.L.cmcb200_shard_handle_bytes()
        mine = np.zeros(nb, np.uint8)
        _lib.check(self.L.cmcb200_shard_export(self.ctx, self.myid, self.procs, mine.ctypes.data), "shard_export")
        with torch.cuda.stream(self._tstream):
            t = torch.from_numpy(mine).cuda(self.device)
            allh = [torch.empty_like(t) for _ in range(self.procs)]
            self.dist.all_gather(allh, t)
            blob = np.concatenate([h.cpu().numpy() for h in allh])
        self._handles = blob
        _lib.check(self.L.cmcb200_shard_attach(self.ctx, blob.ctypes.data), "shard_attach")
        self._token = torch.zeros(1, device=f"cuda:{self.device}")
        self.barrier()

    def barrier(self):
        """stream-ordered barrier between the phases: every rank's kernels enqueued so far complete before any rank's later kernels start"""
        import torch
        with torch.cuda.stream(self._tstream):
            self.dist.all_reduce(self._token)

    def close(self):
        if getattr(self, "ctx", None) and getattr(self, "_handles", None) is not None:
            self.sync()
            self.barrier()
            import torch
            torch.cuda.synchronize()
            self.L.cmcb200_shard_detach(self.ctx)
            self._handles = None
        super().close()

    def load_global(self, r, vr, vt, m_star_units, binind_global=None):
        """Every rank is given the whole radius-sorted cluster and keeps its own shard of star[] plus the full tables."""
        n = len(r)
        lo, hi = int(self.Start[self.myid]), int(self.End[self.myid])
        nl = hi - lo + 1
        self.star_r[0], self.star_m[0] = ZERO, 0.0
        self.star_r[1:n + 1], self.star_m[1:n + 1] = r, m_star_units
        self.star_r[n + 1], self.star_m[n + 1] = SF_INFINITY, 0.0
        for name, val in (("r", r), ("vr", vr), ("vt", vt), ("m", m_star_units)):
            self.field(name)[1:nl + 1] = val[lo - 1:hi]
        self.field("r")[0] = ZERO
        self.field("id")[1:nl + 1] = np.arange(lo, hi + 1)
        if binind_global is not None:
            # binaries of this shard, renumbered 1.. in star order (what the reference's load-time distribution leaves)
            b = np.asarray(binind_global[lo - 1:hi])
            loc = np.zeros(nl, np.int64)
            loc[b > 0] = np.arange(1, int((b > 0).sum()) + 1)
            self.field("binind")[1:nl + 1] = loc
            self.N_b_local = int((b > 0).sum())
            self._shard_binary_source = b[b > 0]      # global binary index of each local binary, for the caller that fills binary[]
        self.clus.N_MAX = n
        self.clus.N_MAX_NEW = nl

    def compute_energy(self):
        """E, J of the local stars as ComputeEnergy does (reference src/cmc/cmc_utils.c:612-616), host arrays"""
        nl, lo = self.clus.N_MAX_NEW, int(self.Start[self.myid])
        vr, vt = self.field("vr")[1:nl + 1], self.field("vt")[1:nl + 1]
        self.field("E")[1:nl + 1] = self.star_phi[lo:lo + nl] + 0.5 * (vr * vr + vt * vt)
        self.field("J")[1:nl + 1] = self.star_r[lo:lo + nl] * vt

    def upload(self):
        n, nb = self.clus.N_MAX_NEW, self.N_b_local
        _lib.check(self.L.cmcb200_upload(self.ctx, self.star[1:].ctypes.data, n, self.binary[1:].ctypes.data, nb, self.star_r.ctypes.data,
                                         self.star_m.ctypes.data, self.star_phi.ctypes.data, min(self._n_table_dim, self.clus.N_MAX + 2)), "upload")

    def download(self):
        n = self.clus.N_MAX_NEW
        _lib.check(self.L.cmcb200_download(self.ctx, self.star[1:].ctypes.data, n, self.binary[1:].ctypes.data, self.N_b_local,
                                           self.star_r.ctypes.data, self.star_m.ctypes.data, self.star_phi.ctypes.data,
                                           min(self._n_table_dim, self.clus.N_MAX + 2)), "download")

    # ---- the sharded sort: qsorts_new + post_sort_comm for procs > 1 ---------------------------------------------------------------
    def dev_transfer_and_sort(self):
        """The whole sharded sort as one asynchronous device sequence (cmcb200_shard_step_async): the ranks meet at barriers made of flags in
        each other's device memory, no torch.distributed call and no host round trip before the single wait at the end.
        CMCB200_SHARD_HOST_BARRIERS=1 selects the older form below (host-side barriers through torch.distributed)."""
        import os
        if os.environ.get("CMCB200_SHARD_HOST_BARRIERS"):
            return self.dev_transfer_and_sort_host_barriers()
        self._shard_step(with_potential=False)

    def _shard_step(self, with_potential):
        """cmcb200_shard_step_async + _result: the sort (and, with_potential, the potential rebuild queued behind the new tables on the second
        stream, overlapping the record exchange); one wait"""
        import math
        n_old = int(self.End[self.myid] - self.Start[self.myid] + 1)
        _lib.check(self.L.cmcb200_shard_step_async(self.ctx, self.clus.N_MAX_NEW, n_old, int(self.Start[self.myid]), self.clus.N_MAX, self.MIN_CHUNK_SIZE,
                                                   max(self.N_b_local, 0), self.clus.N_STAR if with_potential else 0, self.cenma.m, self.cenma.m_new),
                   "shard_step_async")
        keep = np.zeros(self.procs, np.int64)
        S, E, nb, pot = np.zeros(self.procs, np.int32), np.zeros(self.procs, np.int32), C.c_long(0), np.zeros(4)
        _lib.check(self.L.cmcb200_shard_step_result(self.ctx, keep.ctypes.data, S.ctypes.data, E.ctypes.data, C.byref(nb), pot.ctypes.data), "shard_step_result")
        self.Start, self.End = S, E
        self.clus.N_MAX = int(keep.sum())
        self.clus.N_MAX_NEW = int(E[self.myid] - S[self.myid] + 1)
        self.N_b_local = self.bin_slots_in_use = nb.value
        if with_potential:      # what cmcb200_dev_calc_potential_new hands back (reference cmc_utils.c:696-707, 2011)
            self.clus.N_MAX = int(pot[0])
            self.cenma.m, self.cenma.m_new, self.Mtotal = float(pot[1]), float(pot[2]), float(pot[3])
            self.Rtidal = self.orbit_r * math.pow(self.Mtotal, 1.0 / 3.0)
            S2, E2 = find_limits(self.clus.N_MAX, self.MIN_CHUNK_SIZE, self.procs)
            self.mpiBegin, self.mpiEnd = int(S2[self.myid]), int(E2[self.myid])

    def dev_transfer_and_sort_host_barriers(self):
        import torch
        n_old = int(self.End[self.myid] - self.Start[self.myid] + 1)
        first_global, N_MAX_old = int(self.Start[self.myid]), self.clus.N_MAX
        keep = C.c_long(0)
        _lib.check(self.L.cmcb200_shard_sort_local(self.ctx, self.clus.N_MAX_NEW, C.byref(keep)), "shard_sort_local")
        # the allgather of the counts is also the barrier after the key push (stream-ordered on every rank)
        with torch.cuda.stream(self._tstream):
            mine = torch.tensor([keep.value], dtype=torch.int64, device=f"cuda:{self.device}")
            allk = torch.empty(self.procs, dtype=torch.int64, device=f"cuda:{self.device}")
            self.dist.all_gather_into_tensor(allk, mine)
            n_keep_all = allk.cpu().numpy().astype(np.int64)
        S, E = plan_partition(n_keep_all, self.MIN_CHUNK_SIZE)
        _lib.check(self.L.cmcb200_shard_exchange(self.ctx, n_keep_all.ctypes.data, S.ctypes.data, E.ctypes.data, n_old, first_global, N_MAX_old),
                   "shard_exchange")
        self.barrier()
        nl = int(E[self.myid] - S[self.myid] + 1)
        nb = C.c_long(0)
        _lib.check(self.L.cmcb200_shard_publish(self.ctx, nl, int(S[self.myid]), max(self.N_b_local, 0), C.byref(nb)), "shard_publish")
        self.barrier()
        _lib.check(self.L.cmcb200_shard_finish(self.ctx), "shard_finish")
        self.Start, self.End = S, E
        self.clus.N_MAX = int(n_keep_all.sum())
        self.clus.N_MAX_NEW = nl
        self.N_b_local = self.bin_slots_in_use = nb.value

    def dev_qsorts_new(self):
        raise NotImplementedError("the sharded path sorts on rnew with the transfer fused (dev_transfer_and_sort)")

    def dev_calc_potential_new(self):
        super().dev_calc_potential_new()        # replicated on every rank, like potential_calculate (reference cmc_utils.c:660,716)
        S, E = find_limits(self.clus.N_MAX, self.MIN_CHUNK_SIZE, self.procs)   # mpiFindIndicesCustom (cmc_utils.c:2011)
        self.mpiBegin, self.mpiEnd = int(S[self.myid]), int(E[self.myid])

    def dev_compute_energy(self):
        _lib.check(self.L.cmcb200_dev_compute_energy(self.ctx, self.clus.N_MAX_NEW, int(self.Start[self.myid])), "dev_compute_energy")

    def dev_energy_conservation3(self):
        """energy_conservation3 -> set_velocities3 on P ranks (reference src/cmc/cmc_utils.c:254-363): the first loop on every shard, then the
        excess each shard could not absorb travels to the next rank until none is left.  Returns the number of rounds of the second loop."""
        import torch
        p, left = self.orbit_params(), C.c_double()
        _lib.check(self.L.cmcb200_dev_set_velocities3(self.ctx, C.byref(p), C.byref(left)), "dev_set_velocities3")
        with torch.cuda.stream(self._tstream):
            E, rounds = carry_excess_between_ranks(self.dist, left.value, self.dev_set_velocities3_pass, device=f"cuda:{self.device}")
        self.Eoops += -E * (1.0 / self.clus.N_STAR)
        return rounds

    def _ghost_velocities(self, p):
        """what calc_sigma_r exchanges between neighbouring ranks (reference src/cmc/cmc_dynamics_helper.c:2862-2900): vr, vt of the first and
        last p local stars, read from the resident records, all-gathered (MPI_Send / MPI_Recv pairs in the reference)"""
        import torch
        n = int(self.End[self.myid] - self.Start[self.myid] + 1)
        rec = np.zeros(self.star.shape[1], np.uint8)
        mine = np.zeros(4 * p)
        from .layout import STAR_OFFSETS
        for i in range(p):
            for half, j in ((0, 1 + i), (2, n - p + 1 + i)):
                if 1 <= j <= n:
                    _lib.check(self.L.cmcb200_dev_get(self.ctx, rec.ctypes.data, j, None, 0), "dev_get")
                    mine[half * p + i] = rec[STAR_OFFSETS["vr"]:STAR_OFFSETS["vr"] + 8].view(np.float64)[0]
                    mine[(half + 1) * p + i] = rec[STAR_OFFSETS["vt"]:STAR_OFFSETS["vt"] + 8].view(np.float64)[0]
        with torch.cuda.stream(self._tstream):
            t = torch.from_numpy(mine).cuda(self.device)
            parts = [torch.empty_like(t) for _ in range(self.procs)]
            self.dist.all_gather(parts, t)
            allv = [x.cpu().numpy() for x in parts]
        g = {}
        if self.myid > 0:
            g["prev_vr"], g["prev_vt"] = allv[self.myid - 1][2 * p:3 * p], allv[self.myid - 1][3 * p:4 * p]
        if self.myid < self.procs - 1:
            g["next_vr"], g["next_vt"] = allv[self.myid + 1][0:p], allv[self.myid + 1][p:2 * p]
        return g

    def dev_calc_sigma_r(self, p=None, mave=False, ghosts=None, force_sequential=False):
        """calc_sigma_r on P ranks: every rank its own chain from an empty window p stars before its first star, ghost stars from its neighbours"""
        p = int(p or self.DYNAMICS["AVEKERNEL"])
        return super().dev_calc_sigma_r(p, mave=mave, ghosts=self._ghost_velocities(p) if ghosts is None else ghosts, force_sequential=force_sequential)

    def dev_dynamics_apply(self, dt, draws=None, seed=0, **switches):
        """dynamics_apply on P ranks: pairs are formed inside each rank's shard (reference cmc_dynamics.c:216); the scattering-angle statistics
        are summed over the ranks (MPI_Reduce, :514-515)"""
        res = super().dev_dynamics_apply(dt, draws=draws, seed=seed, **switches)
        tot = self._allreduce_sum([res.Nrel] + list(res.Nrelbeta) + list(res.qaverelbeta) + list(res.maverelbeta) + list(res.raverelbeta))
        res.global_Nrel, res.global_Nrelbeta = int(tot[0]), [int(x) for x in tot[1:5]]
        res.global_qaverelbeta, res.global_maverelbeta, res.global_raverelbeta = tot[5:9], tot[9:13], tot[13:17]
        return res

    def _allreduce_min(self, x):
        """MPI_Allreduce(MPI_MIN) of simul_relax_new (reference cmc_dynamics_helper.c:2577)"""
        import torch
        with torch.cuda.stream(self._tstream):
            t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{self.device}")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            return float(t.cpu()[0])

    def _allreduce_sum(self, values):
        import torch
        with torch.cuda.stream(self._tstream):
            t = torch.tensor(list(values), dtype=torch.float64, device=f"cuda:{self.device}")
            self.dist.all_reduce(t)
            return [float(x) for x in t.cpu()]

    def orbit_params(self, draws=None, draw_off=None, seed=0, fuse_energy=False):
        n_old = int(self.End[self.myid] - self.Start[self.myid] + 1)
        n_new_stars = max(0, self.clus.N_MAX_NEW - n_old)
        p = _lib.OrbitParams(self.clus.N_MAX_NEW, n_old, int(self.Start[self.myid]), self.clus.N_MAX, self.clus.N_STAR,
                             self.clus.N_MAX + 2 + n_new_stars, self.cenma.m, self.Rtidal, None, None, seed, self.tcount, int(fuse_energy))
        if draws is not None:
            self._draws = np.ascontiguousarray(draws, np.float64)
            self._draw_off = np.ascontiguousarray(draw_off, np.int64)
            p.draws, p.draw_off = self._draws.ctypes.data, self._draw_off.ctypes.data
        return p

    def resident_begin(self, dirty_words=None):
        """Cluster.resident_begin for a shard: the records of this rank, and of the replicated tables only this rank's slice -- the ranks hand
        the slices to each other over NVLink (cmcb200_resident_begin_sharded)"""
        n, nb = self.clus.N_MAX_NEW, self._n_bin_in_use()
        n_new_stars = max(0, n - int(self.End[self.myid] - self.Start[self.myid] + 1))
        n_table = min(self.N_TABLE_DIM, self.clus.N_MAX + 2 + n_new_stars)
        first = int(self.Start[self.myid])
        n_slice = (n_table - first) if self.myid == self.procs - 1 else int(self.End[self.myid]) - first + 1
        _lib.check(self.L.cmcb200_resident_begin_sharded(self.ctx, self.star[1:].ctypes.data, n, self.binary[1:].ctypes.data, nb, self.star_r.ctypes.data,
                                                         self.star_m.ctypes.data, self.star_phi.ctypes.data, n_table,
                                                         self.ALL_WORDS if dirty_words is None else dirty_words, first, n_slice), "resident_begin_sharded")
        self.N_b_local = nb

    def host_step(self, seed=0, draws=None, draw_off=None, dirty_words=None):
        """Cluster.host_step on this rank's shard: its records and the tables cross this rank's host link once each way; max_r is reduced
        over the ranks like new_orbits_calculate does (reference cmc_utils.c:2206)."""
        import os
        if os.environ.get("CMCB200_SHARD_HOST_BARRIERS"):
            res = super().host_step(seed, draws, draw_off, dirty_words)
        else:
            self.resident_begin(dirty_words)
            p = self.orbit_params(draws, draw_off, seed)
            _lib.check(self.L.cmcb200_dev_new_orbits_async(self.ctx, C.byref(p), 0), "dev_new_orbits_async")
            self._shard_step(with_potential=True)
            res = _lib.OrbitResult()
            _lib.check(self.L.cmcb200_dev_orbit_result(self.ctx, C.byref(p), C.byref(res)), "dev_orbit_result")
            self.last_orbit = res
            self.resident_end()
            self.tcount += 1
        self.max_r = self._allreduce_max(res.max_r)
        return res

    def _allreduce_max(self, x):
        import torch
        with torch.cuda.stream(self._tstream):
            m = torch.tensor([x], dtype=torch.float64, device=f"cuda:{self.device}")
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX)
            return float(m.item())

    def dev_step(self, seed=0, draws=None, draw_off=None, fused=True):
        """orbit step on the local shard -> sharded sort (records to their final rank and slot over NVLink) -> replicated potential.
        max_r is reduced over the ranks like new_orbits_calculate does (reference cmc_utils.c:2206)."""
        import os
        import torch
        if os.environ.get("CMCB200_SHARD_HOST_BARRIERS"):
            res = self.dev_new_orbits(draws, draw_off, seed, fuse_energy=True)
            self.dev_transfer_and_sort()
        else:
            # the orbit step is only queued; its statistics are read once the sort's single wait has drained the stream
            p = self.orbit_params(draws, draw_off, seed, fuse_energy=True)
            _lib.check(self.L.cmcb200_dev_new_orbits_async(self.ctx, C.byref(p), 0), "dev_new_orbits_async")
            self._shard_step(with_potential=True)
            res = _lib.OrbitResult()
            _lib.check(self.L.cmcb200_dev_orbit_result(self.ctx, C.byref(p), C.byref(res)), "dev_orbit_result")
            self.max_r, self.last_orbit = res.max_r, res
        if os.environ.get("CMCB200_SHARD_HOST_BARRIERS"):
            self.dev_calc_potential_new()
        with torch.cuda.stream(self._tstream):
            m = torch.tensor([res.max_r], dtype=torch.float64, device=f"cuda:{self.device}")
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX)
        self._max_r_dev = m                        # read lazily: no host sync in the step
        self.tcount += 1
        return res
