"""Worker of tests/test_gpu_sharded.py and tools/: one rank of a radius-sharded run (one process per GPU, NCCL) checked against a
single-GPU run of the same cluster on the same device.  Philox draws are keyed by the star's GLOBAL index, so the sharded run must
reproduce the single-GPU run bit for bit: every record of the shard, the replicated tables, the binaries."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(rank, world, port, N, nb, steps, kind="plummer", masses="imf", host_barriers=False):
    import torch
    import torch.distributed as dist
    import common as cm
    from cmc_cosmic_b200.cluster import Cluster
    from cmc_cosmic_b200.sharded import ShardedCluster

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if host_barriers:       # the older form of the sharded sort: barriers and the count exchange through torch.distributed
        os.environ["CMCB200_SHARD_HOST_BARRIERS"] = "1"
    else:
        os.environ.pop("CMCB200_SHARD_HOST_BARRIERS", None)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    r, vr, vt, sm = cm.cluster_state(kind, N, seed=17, masses=masses)
    rng = np.random.default_rng(3)
    binind = np.zeros(N, np.int64)
    if nb:
        binind[rng.choice(N, nb, replace=False)] = rng.permutation(np.arange(1, nb + 1))
    gbin = np.zeros((nb + 1, 464), np.uint8)
    if nb:
        gbin[1:] = (np.arange(1, nb + 1, dtype=np.int64)[:, None] * 64 + np.arange(58, dtype=np.int64)[None, :]).view(np.uint8).reshape(nb, 464)

    # ---- sharded run
    c = ShardedCluster(N, dist, device=rank, n_bin_dim=nb + 8)
    c.load_global(r, vr, vt, sm, binind_global=binind if nb else None)
    if nb:
        src = c._shard_binary_source
        c.binary[1:len(src) + 1] = gbin[src]
    c.calc_potential_new()
    c.upload()
    for _ in range(steps):
        c.dev_step(seed=5)
    c.dev_compute_energy()
    c.download()
    c.sync()
    lo, hi, nl = int(c.Start[rank]), int(c.End[rank]), c.clus.N_MAX_NEW
    shard = c.star[1:nl + 1].copy()
    sb = c.binary[1:c.N_b_local + 1].copy()
    tabs = (c.star_r[:c.clus.N_MAX + 2].copy(), c.star_m[:c.clus.N_MAX + 2].copy(), c.star_phi[:c.clus.N_MAX + 2].copy())
    nmax, nbl = c.clus.N_MAX, c.N_b_local

    # ---- the same cluster on this GPU alone
    s = Cluster(N, n_star_dim=N + 64, n_bin_dim=nb + 8, device=rank)
    s.load(r, vr, vt, sm, binind=binind if nb else None)
    s.binary[1:nb + 1] = gbin[1:]
    s.calc_potential_new()
    s.upload()
    for _ in range(steps):
        s.dev_step(seed=5)
    s.dev_compute_energy()
    s.download()
    assert nmax == s.clus.N_MAX, (nmax, s.clus.N_MAX)
    assert hi - lo + 1 == nl
    for a, b, name in zip(tabs, (s.star_r, s.star_m, s.star_phi), ("star_r", "star_m", "star_phi")):
        assert np.array_equal(a.view(np.uint64), b[:nmax + 2].view(np.uint64)), f"rank {rank}: {name} differs"
    want = s.star[lo:hi + 1].copy()
    from cmc_cosmic_b200.layout import star_field
    gb, lb = star_field(want, "binind").copy(), star_field(shard, "binind").copy()
    assert np.array_equal(gb > 0, lb > 0), f"rank {rank}: binary flags differ"
    star_field(want, "binind")[:] = 0
    star_field(shard, "binind")[:] = 0
    bad = np.nonzero((want != shard).any(axis=1))[0]
    assert bad.size == 0, f"rank {rank}: {bad.size} of {nl} records differ, first local {bad[0]}"
    if nb:
        assert nbl == int((lb > 0).sum())
        assert np.array_equal(lb[lb > 0], np.arange(1, nbl + 1)), f"rank {rank}: binind not renumbered 1.. in star order"
        assert np.array_equal(sb, s.binary[gb[gb > 0]]), f"rank {rank}: binary records differ"
    # ---- one more step on HOST arrays (resident_begin: this rank's records and ITS slice of the tables up, the slices handed on over NVLink;
    # resident_end: records and tables down), against the same on the single GPU
    for x in (c, s):
        x.compute_energy()
    c.host_step(seed=9)
    s.host_step(seed=9)
    lo, hi, nl = int(c.Start[rank]), int(c.End[rank]), c.clus.N_MAX_NEW
    assert c.clus.N_MAX == s.clus.N_MAX and c.max_r == s.max_r
    for name in ("star_r", "star_m", "star_phi"):
        assert np.array_equal(getattr(c, name)[:c.clus.N_MAX + 2].view(np.uint64), getattr(s, name)[:c.clus.N_MAX + 2].view(np.uint64)), f"rank {rank}: {name} after host_step"
    want, shard = s.star[lo:hi + 1].copy(), c.star[1:nl + 1].copy()
    star_field(want, "binind")[:] = 0
    star_field(shard, "binind")[:] = 0
    bad = np.nonzero((want != shard).any(axis=1))[0]
    assert bad.size == 0, f"rank {rank}: host_step: {bad.size} of {nl} records differ, first local {bad[0]}"
    c.close()
    s.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world}: sharded == single GPU (N={N}, binaries={nb}, steps={steps} + 1 on host arrays, shard {lo}..{hi})", flush=True)


def run_sv3(rank, world, port, N):
    """energy_conservation3 -> set_velocities3 on `world` GPUs (first loop per shard on the device, the excess carried rank to rank by the
    host layer, second-loop passes on the device) against the oracle run as all ranks in turn -- which tests/ref_multirank_check.py pins
    against the ranks of the compiled reference."""
    import torch
    import torch.distributed as dist
    import common as cm
    import oraclelib as ol
    from cmc_cosmic_b200.sharded import ShardedCluster

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=23, masses="imf")
    pot = cm.oracle_potential(r, sm, N)
    f = cm.sv3_fields(r, vr, vt, pot["phi"][1:N + 1], seed=world)
    c = ShardedCluster(N, dist, device=rank)
    c.load_global(r, vr, vt, sm)
    c.star_phi[:N + 2] = pot["phi"][:N + 2]
    S, E = c.Start.copy(), c.End.copy()
    lo, hi = int(S[rank]), int(E[rank])
    nl = hi - lo + 1
    for k, v in f.items():
        c.field(k)[1:nl + 1] = v[lo - 1:hi]
    c.upload()
    rounds = c.dev_energy_conservation3()
    c.download()
    c.sync()
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
    shards = [(int(S[q]), int(E[q] - S[q] + 1)) for q in range(world)]
    want, eoops, want_rounds = ol.set_velocities3(t, shards, [{k: v[S[q] - 1:E[q]] for k, v in f.items()} for q in range(world)])
    assert rounds == want_rounds, (rounds, want_rounds)
    assert c.Eoops == eoops[rank], (c.Eoops, eoops[rank])
    for k in ol.SV3_IO:
        cm.assert_bits_equal(c.field(k)[1:nl + 1], want[rank][k], f"rank {rank}: {k}")
    c.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world}: set_velocities3 == oracle ranks (N={N}, rounds={rounds}, shard {lo}..{hi})", flush=True)


def run_tidal(rank, world, port, N):
    """tidally_strip_stars on `world` GPUs (criterion per shard on the device, the reference's loop in the host layer, the stripped mass and
    the escaper sums all-reduced) against the oracle run as all ranks -- pinned against the ranks of the compiled reference in
    tests/ref_multirank_check.py."""
    import ctypes
    import torch
    import torch.distributed as dist
    import common as cm
    import oraclelib as ol
    from cmc_cosmic_b200.sharded import ShardedCluster

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    orbit_r = 6.0
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=29, masses="imf")
    pot = cm.oracle_potential(r, sm, N, orbit_r=orbit_r)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    rng = np.random.default_rng(8)
    fg = dict(E=E, J=J, Eint=rng.uniform(0.0, 1e-7, N), r_apo=r * rng.uniform(1.0, 3.0, N), rnew=r.copy(), vrnew=vr.copy(), vtnew=vt.copy())
    c = ShardedCluster(N, dist, device=rank, orbit_r=orbit_r)
    c.load_global(r, vr, vt, sm)
    c.star_phi[:N + 2] = pot["phi"][:N + 2]
    S, En = c.Start.copy(), c.End.copy()
    lo, hi = int(S[rank]), int(En[rank])
    nl = hi - lo + 1
    for k, v in fg.items():
        c.field(k)[1:nl + 1] = v[lo - 1:hi]
    c.Mtotal = pot["Mtotal"]
    c.upload()
    stripped = c.dev_tidally_strip_stars(tidal_treatment=1)
    c.download()
    c.sync()
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
    parts = [{k: v[S[q] - 1:En[q]].copy() for k, v in fg.items()} for q in range(world)]
    states = [dict(Mtotal=pot["Mtotal"], TidalMassLoss=0.0, OldTidalMassLoss=0.0, Eescaped=0.0, Jescaped=0.0, Eintescaped=0.0, Ebescaped=0.0) for _ in range(world)]
    ol.tidally_strip_loop_ranks([(lambda thr, q=q: ol.tidally_strip_select(1, thr, parts[q])) for q in range(world)],
                                lambda x: ol.lib().orc_potential(ctypes.byref(t), x),
                                [(lambda i, q=q: pot["star_m"][S[q] + i - 1]) for q in range(world)],
                                [(lambda i, q=q: dict(E=parts[q]["E"][i - 1], J=parts[q]["J"][i - 1], Eint=parts[q]["Eint"][i - 1], binind=0)) for q in range(world)],
                                [None] * world, states, N, orbit_r, 1, 0.11)
    me = states[rank]
    assert stripped == me["stripped"] and me["passes"] >= 2 and sum(len(s["stripped"]) for s in states) > 20
    for k in ("Rtidal", "TidalMassLoss", "Eescaped", "Jescaped", "Eintescaped", "Etidal"):
        assert getattr(c, k) == me[k], (k, getattr(c, k), me[k])
    for k in ("rnew", "vrnew", "vtnew"):
        cm.assert_bits_equal(c.field(k)[1:nl + 1], parts[rank][k], f"rank {rank}: {k}")
    c.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world}: tidally_strip_stars == oracle ranks (N={N}, {len(stripped)} stripped here, {me['passes']} passes)", flush=True)


if __name__ == "__main__":
    # torchrun entry: RANK / WORLD_SIZE / MASTER_PORT from the environment
    run(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("MASTER_PORT", "29511")),
        int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000, int(sys.argv[2]) if len(sys.argv) > 2 else 5000, int(sys.argv[3]) if len(sys.argv) > 3 else 3)


def run_dyn(rank, world, port, N):
    """calc_sigma_r and dynamics_apply's pair loop on `world` GPUs: every rank its own running sums with ghost stars from its neighbours,
    pairs inside each shard -- against the oracle run as each rank (pinned against the ranks of the compiled reference in
    tests/ref_multirank_check.py for calc_sigma_r)."""
    import torch
    import torch.distributed as dist
    import common as cm
    import oraclelib as ol
    from cmc_cosmic_b200.sharded import ShardedCluster

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=29, masses="imf")
    pot = cm.oracle_potential(r, sm, N)
    c = ShardedCluster(N, dist, device=rank)
    c.load_global(r, vr, vt, sm)
    c.star_phi[:N + 2] = pot["phi"][:N + 2]
    c.compute_energy()
    lo, hi = int(c.Start[rank]), int(c.End[rank])
    nl = hi - lo + 1
    c.field("threebb_interacted")[1:nl + 1] = 0
    c.upload()
    for mave in (False, True):
        a, sg = c.dev_calc_sigma_r(20, mave=mave)
        want_a, want_s = ol.calc_sigma_r(20, pot["star_r"], pot["star_m"], N, N, vr, vt, first_global=lo, n_local=nl, mave=mave)
        cm.assert_bits_equal(a[1:], want_a[1:], f"rank {rank}: r / mean mass")
        cm.assert_bits_equal(sg[1:], want_s[1:], f"rank {rank}: sigma")
    draws = np.random.default_rng(100 + rank).random(3 * (nl // 2))
    dt = 2e-3
    res = c.dev_dynamics_apply(dt, draws=draws)
    c.download()
    c.sync()
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N)
    E, J = cm.energies(pot["phi"][1:N + 1], r, vr, vt)
    f0 = dict(vr=vr[lo - 1:hi].copy(), vt=vt[lo - 1:hi].copy(), E=E[lo - 1:hi].copy(), J=J[lo - 1:hi].copy(), Y=np.zeros(nl), binind=np.zeros(nl, np.int64),
              se_k=np.zeros(nl, np.int32), rad=np.zeros(nl))
    want = ol.dynamics_apply_pairs(t, lo, dt, ol.dyn_params(), draws, f0)
    assert res.Nrel == want["Nrel"] == nl // 2 and res.n_encounters == 0
    cm.assert_bits_equal(c.field("Y")[1:nl + 1], f0["Y"], "Y")
    vs = np.sqrt(f0["vr"] ** 2 + f0["vt"] ** 2)
    for k, scale in (("vr", vs), ("vt", vs), ("E", np.abs(f0["E"]) + vs ** 2), ("J", np.abs(f0["J"]) + 1e-300)):
        assert (np.abs(c.field(k)[1:nl + 1] - f0[k]) / scale).max() <= 1e-12, (rank, k)
    S, E_ = c.Start, c.End
    assert res.global_Nrel == sum(int(E_[q] - S[q] + 1) // 2 for q in range(world))
    c.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world}: calc_sigma_r bit-exact, dynamics_apply within 1e-12 of the oracle rank (N={N}, shard {lo}..{hi})", flush=True)


def run_evolve(rank, world, port, N, t_end):
    """A whole relaxation-driven run (Cluster.evolve: calc_timestep, dynamics_apply, energy scheme, orbits, sort, potential per timestep) on
    `world` GPUs against the same run on one GPU: the partition's chunks are multiples of 2 AVEKERNEL, so the ranks form the same pairs and
    blocks, the Philox draws are keyed by global indices -- the sharded run must leave the same records, bit for bit."""
    import torch
    import torch.distributed as dist
    import common as cm
    from cmc_cosmic_b200.cluster import Cluster
    from cmc_cosmic_b200.sharded import ShardedCluster
    from cmc_cosmic_b200.layout import star_field

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=41)
    pcs = np.array([0.01, 0.1, 0.5, 0.9])
    c = ShardedCluster(N, dist, device=rank)
    c.load_global(r, vr, vt, sm)
    c.calc_potential_new()
    c.compute_energy()
    c.upload()
    hist = c.evolve(t_end=t_end, max_steps=400, seed=13, tidally_strip=None, GAMMA=0.1)
    c.download()
    c.sync()
    lo, hi, nl = int(c.Start[rank]), int(c.End[rank]), c.clus.N_MAX_NEW
    s = Cluster(N, n_star_dim=N + 64, device=rank)
    s.load(r, vr, vt, sm)
    s.calc_potential_new()
    s.compute_energy()
    s.upload()
    hist1 = s.evolve(t_end=t_end, max_steps=400, seed=13, tidally_strip=None, GAMMA=0.1)
    s.download()
    assert len(hist) == len(hist1) and len(hist) > 20, (len(hist), len(hist1))
    assert [h[:3] for h in hist] == [h[:3] for h in hist1], "the time steps of the sharded run differ"
    assert c.clus.N_MAX == s.clus.N_MAX
    want, shard = s.star[lo:hi + 1].copy(), c.star[1:nl + 1].copy()
    bad = np.nonzero((want != shard).any(axis=1))[0]
    assert bad.size == 0, f"rank {rank}: {bad.size} of {nl} records differ after {len(hist)} timesteps, first local {bad[0]}"
    for name, top in (("star_r", 2), ("star_m", 1), ("star_phi", 2)):      # star_m[N_MAX + 1] is whatever the last removed star left (the reference never sets it)
        assert np.array_equal(getattr(c, name)[:c.clus.N_MAX + top].view(np.uint64), getattr(s, name)[:c.clus.N_MAX + top].view(np.uint64)), name
    c.close()
    s.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world}: {len(hist)} timesteps to t = {hist[-1][0]:.4f} sharded == single GPU (N={N}, shard {lo}..{hi})", flush=True)
