"""Host-side logic of the radius-sharded path, run on CPU as 2 and 3 gloo ranks: the partition arithmetic, the global-rank rule the
exchange kernel implements (restated in numpy: cmc_cosmic_b200.sharded.host_global_ranks), the metadata flow (allgather of the kept
counts -> findLimits -> destinations).  The records travel through gloo here (on the GPU they are stored into peer memory by the
library's kernels); the result must be the oracle's global sort, trimmed and partitioned like the reference's sample_sort +
load_balance (src/cmc/cmc_sort.c:375-791) leaves it -- cf. tests/ref_multirank_check.py for the reference itself."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, n_strip, ties):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oraclelib as ol
    from cmc_cosmic_b200.sharded import destinations, host_global_ranks, plan_partition
    from cmc_cosmic_b200.partition import find_limits
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)                       # same stream on every rank: everybody knows the whole input
    key = rng.integers(1, 60, N).astype(np.float64) if ties else rng.uniform(1e-3, 1e2, N)
    key[rng.choice(N, n_strip, replace=False)] = 1.0e10   # stripped stars
    ident = np.arange(1, N + 1)
    S0, E0 = find_limits(N, 40, world)
    lo, hi = int(S0[rank]), int(E0[rank])
    my_key, my_id = key[lo - 1:hi], ident[lo - 1:hi]
    # 1. local stable sort, trailing stripped stars trimmed (remove_stripped_stars, cmc_sort.c:343-354)
    order = np.argsort(my_key, kind="stable")
    sk, sid = my_key[order], my_id[order]
    n_keep = int((sk < 1.0e10).sum())
    sk, sid = sk[:n_keep], sid[:n_keep]
    # 2. everybody gets every run (on the GPU: keys pushed into each rank's key table) and the kept counts
    runs = [None] * world
    dist.all_gather_object(runs, sk)
    n_keep_all = [len(x) for x in runs]
    S, E = plan_partition(n_keep_all, 40)
    S_or, E_or = ol.find_limits(int(sum(n_keep_all)), 40, world)
    assert np.array_equal(S, S_or) and np.array_equal(E, E_or)
    # 3. global ranks -> destination rank and slot
    g = host_global_ranks(runs, rank)
    d, pos = destinations(g, S)
    assert np.all((d >= 0) & (d < world)) and np.all(pos >= 0) and np.all(pos <= (E - S)[d])
    # 4. the exchange (gloo stands in for the peer stores)
    parcels = [None] * world
    dist.all_gather_object(parcels, [(int(q), pos[d == q], sk[d == q], sid[d == q]) for q in range(world)])
    nl = int(E[rank] - S[rank] + 1)
    out_k, out_id, filled = np.zeros(nl), np.zeros(nl, np.int64), np.zeros(nl, bool)
    for src in parcels:
        q, p, k, i = src[rank]
        assert q == rank and not filled[p].any()
        out_k[p], out_id[p], filled[p] = k, i, True
    assert filled.all(), "every slot of the new shard is written exactly once"
    # 5. against the oracle: stable global sort of the concatenated shards, trimmed, cut at the partition
    n_new = int((key < 1.0e10).sum())
    go = np.argsort(key, kind="stable")[:n_new]
    assert sum(n_keep_all) == n_new
    assert np.array_equal(out_id, ident[go][S[rank] - 1:E[rank]]) and np.array_equal(out_k, key[go][S[rank] - 1:E[rank]])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,N,n_strip,ties", [(2, 10007, 0, False), (2, 5000, 137, True), (3, 20011, 45, False), (4, 1000, 3, True)])
def test_sharded_sort_plan_on_gloo(world, N, n_strip, ties):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, 29600 + world * 7 + (N % 50), N, n_strip, ties), nprocs=world, join=True)


def test_global_rank_rule_is_a_stable_sort_of_the_concatenation():
    sys.path.insert(0, ROOT)
    from cmc_cosmic_b200.sharded import host_global_ranks
    rng = np.random.default_rng(1)
    for P in (1, 2, 5, 8):
        runs = [np.sort(rng.integers(0, 40, rng.integers(0, 300)).astype(np.float64)) for _ in range(P)]
        allv = np.concatenate(runs)
        inv = np.empty(len(allv), np.int64)
        inv[np.argsort(allv, kind="stable")] = np.arange(len(allv))
        got = np.concatenate([host_global_ranks(runs, g) for g in range(P)]) if len(allv) else inv
        assert np.array_equal(got, inv)


def _sv3_worker(rank, world, port, N):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import common as cm
    import oraclelib as ol
    from cmc_cosmic_b200.sharded import carry_excess_between_ranks
    from cmc_cosmic_b200.partition import find_limits
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r, vr, vt, sm = cm.cluster_state("plummer", N, seed=5, masses="imf")
    pot = cm.oracle_potential(r, sm, N)
    f = cm.sv3_fields(r, vr, vt, pot["phi"][1:N + 1], seed=world)
    t = ol.tables(pot["star_r"], pot["star_m"], pot["phi"], N, N, 0.0)
    S, E = find_limits(N, 40, world)
    shards = [(int(S[q]), int(E[q] - S[q] + 1)) for q in range(world)]
    parts = [{k: v[S[q] - 1:E[q]] for k, v in f.items()} for q in range(world)]
    want_io, want_eoops, want_rounds = ol.set_velocities3(t, shards, parts)           # the oracle as all ranks in turn
    g = np.arange(shards[rank][0], shards[rank][0] + shards[rank][1])
    io, flags, left = ol.set_velocities3_first(t, g, parts[rank])                       # this rank's first loop
    Eex, rounds = carry_excess_between_ranks(dist, left, lambda e: ol.set_velocities3_second(t, g, parts[rank]["interacted"], flags, e, io))
    assert rounds == want_rounds and -Eex * t.madhoc == want_eoops[rank]
    for k in ol.SV3_IO:
        cm.assert_bits_equal(io[k], want_io[rank][k], k)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 10000), (3, 10007), (4, 50000)])
def test_excess_energy_carry_between_ranks_on_gloo(world, N):
    """the message pattern of set_velocities3's do-while loop (reference cmc_utils.c:313-347) as the sharded host layer runs it, with the
    oracle's loops standing in for the kernels: same rounds, same velocities, same Eoops as the ranks of the reference (which
    tests/ref_multirank_check.py pins against the same oracle driver)"""
    import torch.multiprocessing as mp
    mp.spawn(_sv3_worker, args=(world, 29700 + world * 5 + (N % 40), N), nprocs=world, join=True)
