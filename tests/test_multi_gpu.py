"""Multi-GPU path.  GPU part (needs >= 2 GPUs, skipped otherwise): torchrun launches
scripts/multi_gpu_driver_check.py -- distributed timestep vs the oracle, 1e-8 on fields, 1e-10 on
positions.  CPU part (gloo, world_size 2): the host-side logic of the marker-parallel scheme --
shares partition the cloud, and all-reduced per-share node sums reproduce the single-process
trac2grid of the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


@pytest.mark.gpu
@pytest.mark.parametrize("mode,port", [("index", 29533), ("slab", 29535), ("slab reduce", 29537), ("slab local", 29539)])
def test_two_gpu_timestep_matches_oracle(mode, port):
    """index: every rank keeps its markers, fields replicated, raw node sums all-reduced.  slab / slab reduce:
    slab-owned markers with migration, node sums by all-reduce / by boundary-row exchange + all-gather.
    slab local (bench.py's multi-GPU default): slab-owned markers AND slab-local grid fields -- boundary-row
    accumulate, halo rows, all-reduced scalars; no full-plane collective in the step.  The first three ran
    green on 2 B200s in round 2 (profiles/r02_multi_gpu.log)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "multi_gpu_driver_check.py"), "64", "3"] + mode.split()
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "MULTI_GPU_PARITY OK" in r.stdout


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import pylamp_oracle as O
    from pylamp_b200 import setups
    nx, L, tr_x, tr_f, opts = setups.convection(ncell=16)
    M = tr_x.shape[0]
    lo, hi = (rank * M) // world, ((rank + 1) * M) // world
    grid = O.make_grids(nx, L)[0]
    # per-share raw sums (what each rank's scatter kernel produces), then the all-reduce
    x, f = tr_x[lo:hi], tr_f[lo:hi, O.TR_RH0]
    ie = np.floor((nx[0] - 1) * x[:, 0] / L[0]).astype(int)
    je = np.floor((nx[1] - 1) * x[:, 1] / L[1]).astype(int)
    az = (x[:, 0] - grid[0][ie]) / (grid[0][ie + 1] - grid[0][ie])
    ax = (x[:, 1] - grid[1][je]) / (grid[1][je + 1] - grid[1][je])
    w = [(1 - ax) * (1 - az), (1 - ax) * az, ax * (1 - az), ax * az]
    wsum, fsum = np.zeros(nx), np.zeros(nx)
    for c, (di, dj) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        np.add.at(wsum, (ie + di, je + dj), w[c])
        np.add.at(fsum, (ie + di, je + dj), w[c] * f)
    t = torch.as_tensor(np.stack([wsum, fsum]))
    dist.all_reduce(t)
    mm = torch.tensor([-x[:, 0].min(), x[:, 0].max(), -x[:, 1].min(), x[:, 1].max()])
    dist.all_reduce(mm, op=dist.ReduceOp.MAX)           # the (-min, max) trick of plb_marker_minmax
    cnt = torch.tensor([hi - lo])
    dist.all_reduce(cnt)
    if rank == 0:
        ref = [np.zeros(nx)]
        O.trac2grid(tr_x, tr_f[:, [O.TR_RH0]], None, grid, ref, nx, avgscheme=[5])
        out = (t[1] / t[0]).numpy()
        q.put((float(np.abs(out - ref[0]).max() / np.abs(ref[0]).max()), int(cnt.item()) == M,
               bool(np.isclose(-mm[0].item(), tr_x[:, 0].min()) and np.isclose(mm[3].item(), tr_x[:, 1].max()))))
    dist.destroy_process_group()


def test_marker_parallel_host_logic_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29541, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, count_ok, mm_ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-13 and count_ok and mm_ok


def test_slab_partition_rules():
    """Rows per rank of the z-slab solver (mirrors build_levels in csrc/stokes.cu): a level is
    distributed while every rank keeps an even number (>= 4) of cell rows."""
    def dist_levels(ncell, R):
        cells = [ncell]
        while cells[-1] % 2 == 0 and cells[-1] // 2 >= 4:
            cells.append(cells[-1] // 2)
        n = 0
        for l, c in enumerate(cells):
            if l < len(cells) - 1 and c % R == 0 and c // R >= 4 and (c // R) % 2 == 0:
                n += 1
            else:
                break
        return n
    assert dist_levels(4096, 8) == 8 and dist_levels(4096, 2) == 10 and dist_levels(64, 2) == 4
    assert dist_levels(200, 8) == 0        # 25 rows per rank: not even -> plb_stokes_create refuses
