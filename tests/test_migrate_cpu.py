"""Slab marker ownership + migration (pylamp_b200/migrate.py) on CPU tensors: the compaction plan
on its own, and the whole exchange under gloo with world sizes 2 and 3 -- after `migrate` every
marker sits on its owner, the global multiset of (position, properties, velocity) rows is unchanged
bit for bit, aliased property columns stay aliased, arrays grow and shrink through the spare
capacity, and per-slab trac2grid sums of the migrated cloud reproduce the oracle's single-process
result once the shared boundary rows are added."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import ROOT
from pylamp_b200 import migrate as MG


def test_slab_bounds_and_owner():
    assert MG.slab_bounds(4096, 8) == [512 * r for r in range(9)]
    assert MG.slab_bounds(10, 3) == [0, 3, 6, 10]
    nz, Lz = 11, 2.0
    z = torch.tensor([0.0, 0.59999, 0.6, 1.19, 1.2, 1.9999, 2.0, -0.1, 2.5], dtype=torch.float64)
    own = MG.owner_of(z, nz, Lz, MG.slab_bounds(nz - 1, 3))
    ie = np.clip(np.floor((nz - 1) * z.numpy() / Lz).astype(int), 0, nz - 2)      # pylamp2.py:588
    assert own.tolist() == [int(np.searchsorted([3, 6], i, side="right")) for i in ie]
    assert own.tolist()[:2] == [0, 0] and own.tolist()[-3:] == [2, 0, 2]


@pytest.mark.parametrize("seed", range(6))
def test_compaction_plan_random(seed):
    rng = np.random.default_rng(seed)
    for _ in range(50):
        M = int(rng.integers(1, 60))
        leave = rng.random(M) < rng.random()
        n_arr = int(rng.integers(0, 40))
        data = torch.arange(M, dtype=torch.float64)
        arrivals = 1000.0 + torch.arange(n_arr, dtype=torch.float64)
        mask = torch.as_tensor(leave)
        holes = torch.nonzero(mask).flatten()
        M_new, fill, src, dst = MG.compaction_plan(M, holes, mask, n_arr)
        assert M_new == M - int(leave.sum()) + n_arr
        t = MG.resize_rows(data, max(M, M_new))
        if n_arr:
            t.index_copy_(0, fill, arrivals)
        if src.numel():
            assert int(dst.max()) < M_new <= int(src.min())
            t.index_copy_(0, dst, t.index_select(0, src))
        t = t[:M_new]
        want = sorted(data[~mask].tolist() + arrivals.tolist())
        assert sorted(t.tolist()) == want


def test_resize_rows_reuses_spare_capacity():
    t = MG.empty_rows(1000, (2,), torch.float64, "cpu")
    t[:] = 1.0
    ptr = t.data_ptr()
    small = MG.resize_rows(t, 900)
    again = MG.resize_rows(small, 1010)                  # fits the spare room: no new allocation
    assert again.data_ptr() == ptr and again.shape == (1010, 2) and bool((again[:900] == 1.0).all())
    big = MG.resize_rows(again, 100000)                  # does not fit: reallocated, prefix kept
    assert big.data_ptr() != ptr and bool((big[:1010] == again).all())


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import pylamp_oracle as O
    from pylamp_b200 import migrate as MG, setups
    ncell = 24
    nx, L, tr_x, tr_f, opts = setups.convection(ncell=ncell)
    M = tr_x.shape[0]
    bounds = MG.slab_bounds(ncell, world)
    ie = np.clip(np.floor(ncell * tr_x[:, 0] / L[0]).astype(int), 0, ncell - 1)
    own = np.searchsorted(bounds[1:-1], ie, side="right")
    mine = own == rank
    s = SimpleNamespace(nx=nx, L=L, tr_x=torch.as_tensor(tr_x[mine].copy()),
                        cols=[torch.as_tensor(np.ascontiguousarray(tr_f[mine, k])) for k in range(O.NFTRAC)],
                        trac_vel=torch.as_tensor(np.random.default_rng(rank).random((int(mine.sum()), 2))))
    s.cols[O.TR_IHT] = s.cols[O.TR_MRK]                       # aliased columns, like setups.convection_device
    ids0 = s.cols[O.TR__ID].clone()
    rows = []
    ok = MG.check_ownership(s) == 0
    rng = np.random.default_rng(100 + rank)
    dz = L[0] / ncell
    log = []
    for step in range(6):
        # displace: mostly < half a cell; step 3 pushes everything one way (growth on one rank, shrink on the other)
        d = (rng.random(s.tr_x.shape[0]) - 0.5) * 0.9 * dz
        if step == 3:
            d = np.full(s.tr_x.shape[0], 1.3 * dz)
        if step == 4:
            d = np.full(s.tr_x.shape[0], -2.1 * dz)              # more than one slab boundary away is fine too
        s.tr_x[:, 0] = torch.clamp(s.tr_x[:, 0] + torch.as_tensor(d), 2.0 ** -10, L[0] - 2.0 ** -10)
        before = torch.cat([s.tr_x, torch.stack(s.cols, 1), s.trac_vel], 1)
        st = MG.migrate(s)
        log.append((st["sent"], st["received"]))
        ok = ok and MG.check_ownership(s) == 0 and s.tr_x.shape[0] == st["markers"]
        ok = ok and all(c.shape[0] == st["markers"] for c in s.cols) and s.trac_vel.shape[0] == st["markers"]
        ok = ok and s.cols[O.TR_IHT].data_ptr() == s.cols[O.TR_MRK].data_ptr()
        after = torch.cat([s.tr_x, torch.stack(s.cols, 1), s.trac_vel], 1)
        # global multiset of rows unchanged (bit for bit): gather on rank 0, sort by the unique id
        gathered = [None] * world
        dist.gather_object((before.numpy(), after.numpy()), gathered if rank == 0 else None, dst=0)
        if rank == 0:
            b = np.concatenate([g[0] for g in gathered])
            a = np.concatenate([g[1] for g in gathered])
            kb, ka = np.argsort(b[:, 2 + O.TR__ID], kind="stable"), np.argsort(a[:, 2 + O.TR__ID], kind="stable")
            ok = ok and a.shape == b.shape and np.array_equal(a[ka], b[kb]) and a.shape[0] == M
    # trac2grid of the migrated, slab-owned cloud: per-slab raw sums + all-reduce == oracle on the whole cloud
    grid = O.make_grids(nx, L)[0]
    x, f = s.tr_x.numpy(), s.cols[O.TR_TMP].numpy()
    iz = np.floor((nx[0] - 1) * x[:, 0] / L[0]).astype(int)
    jx = np.floor((nx[1] - 1) * x[:, 1] / L[1]).astype(int)
    az = (x[:, 0] - grid[0][iz]) / (grid[0][iz + 1] - grid[0][iz])
    ax = (x[:, 1] - grid[1][jx]) / (grid[1][jx + 1] - grid[1][jx])
    w = [(1 - ax) * (1 - az), (1 - ax) * az, ax * (1 - az), ax * az]
    wsum, fsum = np.zeros(nx), np.zeros(nx)
    for c, (di, dj) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        np.add.at(wsum, (iz + di, jx + dj), w[c])
        np.add.at(fsum, (iz + di, jx + dj), w[c] * f)
    # a slab's markers only touch its own node rows [bounds[r], bounds[r+1]] (the last one shared)
    touched = np.nonzero(wsum.sum(axis=1))[0]
    ok = ok and touched.min() >= bounds[rank] and touched.max() <= bounds[rank + 1]
    t = torch.as_tensor(np.stack([wsum, fsum]))
    dist.all_reduce(t)
    allx = [None] * world
    dist.gather_object((x, f), allx if rank == 0 else None, dst=0)
    if rank == 0:
        X, F = np.concatenate([a[0] for a in allx]), np.concatenate([a[1] for a in allx])
        ref = [np.zeros(nx)]
        O.trac2grid(X, F[:, None], None, grid, ref, nx, avgscheme=[5])
        err = float(np.abs((t[1] / t[0]).numpy() - ref[0]).max() / np.abs(ref[0]).max())
        q.put((bool(ok), err, log))
    else:
        q.put((bool(ok), 0.0, log))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(2, 29551), (3, 29553)])
def test_migration_gloo(world, port):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] for r in res), res
    assert max(r[1] for r in res) < 1e-13
    logs = [r[2] for r in res]
    assert any(sent > 0 for lg in logs for sent, _ in lg)              # markers did migrate
    assert any(rcv > snt for lg in logs for snt, rcv in lg) and any(rcv < snt for lg in logs for snt, rcv in lg)


def _inject_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import pylamp_oracle as O
    from pylamp_b200 import markers, migrate as MG
    rng = np.random.default_rng(3)
    nx, L = [13, 7], [1.0, 0.5]
    M = 900
    tr_x = rng.random((M, 2)) * L
    tr_x[:, 0] = tr_x[:, 0] ** 2            # uneven density: under-populated and empty cells in the lower slabs
    tr_f = rng.random((M, O.NFTRAC))
    tr_f[:, O.TR__ID] = np.arange(M)
    bounds = MG.slab_bounds(nx[0] - 1, world)
    ie = np.clip(np.floor((nx[0] - 1) * tr_x[:, 0] / L[0]).astype(int), 0, nx[0] - 2)
    mine = np.searchsorted(bounds[1:-1], ie, side="right") == rank
    grid = O.make_grids(nx, L)[0]
    kel, cnt = O.cell_index_count(tr_x[mine], nx, L)
    s = SimpleNamespace(nx=nx, L=L, grid=grid, tr_x=torch.as_tensor(tr_x[mine].copy()),
                        cols=[torch.as_tensor(np.ascontiguousarray(tr_f[mine, k])) for k in range(O.NFTRAC)],
                        kelem=torch.as_tensor(kel), count=torch.as_tensor(cnt))
    m0 = int(mine.sum())
    n = markers.inject_markers(s, 12, 6, cell_rows=(bounds[rank], bounds[rank + 1]))
    new_x = s.tr_x.numpy()[m0:]
    new_f = np.stack([c.numpy()[m0:] for c in s.cols], axis=1)
    ok = new_x.shape[0] == n and MG.check_ownership(s) == 0            # injected markers lie in the own slab
    gathered = [None] * world
    dist.gather_object((new_x, new_f), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        so = O.State(nx, L, tr_x.copy(), tr_f.copy())
        so.kelem, so.count = O.cell_index_count(so.tr_x, nx, L)
        np.random.seed(0)
        n_ref = O.inject_markers(so, 12, 6)
        X = np.concatenate([g[0] for g in gathered])
        F = np.concatenate([g[1] for g in gathered])
        ok = ok and X.shape[0] == n_ref > 0
        # same ids (the reference's running-maximum numbering, continued over the ranks in cell order),
        # same cell-mean properties, same cells -- rows in the reference's order already
        ok = ok and np.allclose(F, so.tr_f[M:], rtol=1e-12, atol=0, equal_nan=True)
        k_ref = O.cell_index_count(so.tr_x[M:], nx, L)[0]
        k_new = O.cell_index_count(X, nx, L)[0]
        ok = ok and np.array_equal(k_new, k_ref)
    q.put(bool(ok))
    dist.destroy_process_group()


def test_slab_injection_gloo_world3():
    """Marker injection with slab-owned markers: every rank serves the cells of its slab, ids continue
    over the ranks like the reference's single loop (pylamp2.py:599-633)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_inject_worker, args=(r, 3, 29557, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res)
