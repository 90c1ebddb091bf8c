"""World-size-2 test of the m-cyclic sharding plan on CPU (gloo): each rank takes the Mmn slices the C ABI's
plan assigns to it, computes its partial RPA epsilon / Sigma_c diagonal with the oracle, and the partials are
combined with the same collectives the GPU path uses (sum all-reduce).  Must equal the unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import rpa as orpa
    from oracle import threecenter
    from votca_b200._capi import capi
    api = capi()
    rng = np.random.default_rng(3)
    naux, mtotal, ntotal, homo = 30, 12, 19, 6
    tc = threecenter.TCMatrix(naux, 0, mtotal - 1, 0, ntotal - 1)
    tc.M = rng.standard_normal((mtotal, ntotal, naux)) / 5
    e = np.sort(rng.uniform(-1, 2, ntotal))
    mine = [m for m in range(mtotal) if api.gwbse_shard_owner(m, world) == rank]
    assert len(mine) == api.gwbse_shard_local_count(mtotal, rank, world)
    assert [api.gwbse_shard_local_index(m, world) for m in mine] == list(range(len(mine)))
    # partial epsilon over the occupied levels this rank owns (rpa.cc:92 loop restricted to the shard)
    n_occ, n_unocc = homo + 1, ntotal - homo - 1
    part = np.zeros((naux, naux))
    for m in mine:
        if m >= n_occ:
            continue
        Mv = tc[m][ntotal - n_unocc:, :]
        dE = e[ntotal - n_unocc:] - e[m]
        d = 4.0 * dE / (dE * dE + 0.25)
        part += Mv.T @ (d[:, None] * Mv)
    t = torch.from_numpy(part)
    dist.all_reduce(t)
    eps = t.numpy() + np.eye(naux)
    r = orpa.RPA(tc)
    r.configure(homo, 0, ntotal - 1)
    r.set_rpa_input_energies(e)
    ref = r.calculate_epsilon_i(0.5)
    err = np.abs(eps - ref).max()
    # every level is owned exactly once
    owned = torch.zeros(mtotal, dtype=torch.float64)
    owned[mine] = 1.0
    dist.all_reduce(owned)
    ok = bool(torch.all(owned == 1.0))
    out[rank] = (err, ok)
    dist.destroy_process_group()


def test_sharded_epsilon_matches_unsharded_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        err, ok = out[rank]
        assert err < 1e-12 and ok


def test_shard_plan_pure_functions():
    from votca_b200._capi import capi
    api = capi()
    for world in (1, 2, 3, 8):
        for total in (1, 7, 431):
            counts = [api.gwbse_shard_local_count(total, r, world) for r in range(world)]
            assert sum(counts) == total
            assert max(counts) - min(counts) <= 1
            for m in range(total):
                r = api.gwbse_shard_owner(m, world)
                assert 0 <= r < world and api.gwbse_shard_local_index(m, world) < counts[r]


def _fill_worker(rank, world, port, out):
    """Aux-sharded fill (gwbse_mmn_fill_begin/end): rank g contracts its aux share for all m, the exchange moves
    segment [dest][chi_loc][m_loc][n] to rank dest - emulated with the oracle contraction and gloo."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import threecenter
    from votca_b200._capi import capi
    api = capi()
    rng = np.random.default_rng(11)
    N, naux, mtotal = 13, 11, 7
    mos = rng.standard_normal((N, N))
    ao = rng.standard_normal((naux, N, N))
    ao = ao + ao.transpose(0, 2, 1)
    ref = threecenter.TCMatrix(naux, 0, mtotal - 1, 0, N - 1)
    ref.fill_3c_mo(ao, mos)
    lo, hi = api.gwbse_shard_aux_begin(naux, rank, world), api.gwbse_shard_aux_begin(naux, rank + 1, world)
    part = threecenter.TCMatrix(hi - lo, 0, mtotal - 1, 0, N - 1)
    part.fill_3c_mo(ao[lo:hi], mos)  # all m, own aux functions: part.M[m, n, chi_loc]
    mlmax = (mtotal + world - 1) // world
    cnt_max = max(api.gwbse_shard_aux_begin(naux, r + 1, world) - api.gwbse_shard_aux_begin(naux, r, world)
                  for r in range(world))
    send = np.zeros((world, cnt_max, mlmax, N))
    for m in range(mtotal):
        send[api.gwbse_shard_owner(m, world), :hi - lo, api.gwbse_shard_local_index(m, world), :] = part.M[m].T
    gathered = [torch.zeros(send.shape, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(send))
    # what this rank receives: from every source its segment [rank]
    local = np.zeros((naux, mlmax, N))
    for src in range(world):
        slo, shi = api.gwbse_shard_aux_begin(naux, src, world), api.gwbse_shard_aux_begin(naux, src + 1, world)
        local[slo:shi] = gathered[src].numpy()[rank, :shi - slo]
    err = 0.0
    for m in range(mtotal):
        if api.gwbse_shard_owner(m, world) == rank:
            err = max(err, np.abs(local[:, api.gwbse_shard_local_index(m, world), :].T - ref.M[m]).max())
    out[rank] = err
    dist.destroy_process_group()


def test_aux_sharded_fill_plan_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_fill_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world and all(out[r] < 1e-12 for r in range(world))


def test_aux_partition_pure_function():
    from votca_b200._capi import capi
    api = capi()
    for world in (1, 2, 3, 8):
        for naux in (1, 5, 104, 3177):
            b = [api.gwbse_shard_aux_begin(naux, r, world) for r in range(world + 1)]
            assert b[0] == 0 and b[-1] == naux and all(b[i] <= b[i + 1] for i in range(world))
            sizes = [b[i + 1] - b[i] for i in range(world)]
            assert max(sizes) - min(sizes) <= 1
