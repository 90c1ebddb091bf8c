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
