"""Slab-wise node sums (pylamp_b200/slabgrid.py) under gloo, world sizes 2 and 3: every rank
scatters only its own (slab-owned) markers -- emulated with NumPy exactly like the scatter kernel:
raw sums of weights and weight*value on the extended target axes -- then boundary rows are
exchanged with the neighbours, every rank finalises its own rows, rows are all-gathered.  The
result must equal the oracle's trac2grid of the whole cloud on the node target and on the
z-staggered target (one ghost row prepended), and no sums may lie beyond the halo rows."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT


def _raw_sums(x, f, axz, axx):
    nze, nxe = len(axz), len(axx)
    iz = np.clip(np.searchsorted(axz, x[:, 0], side="right") - 1, 0, nze - 2)
    jx = np.clip(np.searchsorted(axx, x[:, 1], side="right") - 1, 0, nxe - 2)
    az = (x[:, 0] - axz[iz]) / (axz[iz + 1] - axz[iz])
    ax = (x[:, 1] - axx[jx]) / (axx[jx + 1] - axx[jx])
    w = [(1 - ax) * (1 - az), (1 - ax) * az, ax * (1 - az), ax * az]
    planes = np.zeros((2, nze, nxe))
    for c, (di, dj) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        np.add.at(planes[0], (iz + di, jx + dj), w[c] * f)
        np.add.at(planes[1], (iz + di, jx + dj), w[c])
    return planes


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import pylamp_oracle as O
    from pylamp_b200 import migrate as MG, setups, slabgrid as SG
    from pylamp_b200.pylamp_trac import _extended_axis
    ncell = 24
    nx, L, tr_x, tr_f, opts = setups.convection(ncell=ncell)
    rng = np.random.default_rng(5)
    tr_x = np.clip(tr_x + (rng.random(tr_x.shape) - 0.5) * 1.7 * L[0] / ncell, 2.0 ** -10, L[0] - 2.0 ** -10)
    bounds = MG.slab_bounds(ncell, world)
    ie = np.clip(np.floor(ncell * tr_x[:, 0] / L[0]).astype(int), 0, ncell - 1)
    mine = np.searchsorted(bounds[1:-1], ie, side="right") == rank
    x, f = tr_x[mine], tr_f[mine, O.TR_TMP]
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    ok, errs = True, []
    for target in ([grid[0], grid[1]], [gridmp[0], grid[1]], [gridmp[0], gridmp[1]]):
        # extended axes from the GLOBAL marker extent (plb_marker_minmax all-reduces it)
        axz, lz, _ = _extended_axis(target[0], tr_x[:, 0].min(), tr_x[:, 0].max())
        axx, lx, _ = _extended_axis(target[1], tr_x[:, 1].min(), tr_x[:, 1].max())
        planes = torch.as_tensor(_raw_sums(x, f, axz, axx))
        p = SG.row_partition(bounds, lz, len(axz))
        SG.exchange_boundary_rows(planes, p, rank, world, check=True)
        # finalise the own rows, crop like the reference (pylamp_trac.py:313-316), all-gather
        qrows = [0] + bounds[1:-1] + [nx[0]]
        out = torch.zeros(tuple(nx), dtype=torch.float64)
        lo, hi = qrows[rank], qrows[rank + 1]
        with np.errstate(invalid="ignore", divide="ignore"):
            val = (planes[0] / planes[1])[lz + lo:lz + hi, lx:lx + nx[1]]
        out[lo:hi] = val
        SG.gather_rows([out], qrows, rank, world)
        ref = [np.zeros(nx)]
        with np.errstate(invalid="ignore", divide="ignore"):
            O.trac2grid(tr_x, tr_f[:, [O.TR_TMP]], None, target, ref, nx, avgscheme=[5])
        same_nan = np.array_equal(np.isnan(out.numpy()), np.isnan(ref[0]))
        err = np.nanmax(np.abs(out.numpy() - ref[0])) / np.nanmax(np.abs(ref[0]))
        ok = ok and same_nan and err < 1e-13
        errs.append(float(err))
    q.put((bool(ok), errs))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(2, 29561), (3, 29563)])
def test_slab_node_sums_gloo(world, port):
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


def test_row_partition():
    from pylamp_b200 import slabgrid as SG
    assert SG.row_partition([0, 8, 16, 24], 1, 26) == [0, 9, 17, 26]
    assert SG.row_partition([0, 24], 0, 25) == [0, 25]
