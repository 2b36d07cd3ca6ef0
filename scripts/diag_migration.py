"""Do markers cross slab boundaries in the benchmark run, and are they migrated?  (torchrun, 2+ GPUs)
Prints per step and rank: dt limiter, largest marker displacement in cells, markers outside the rank's slab before /
after migration, the migration counts.   python -m torch.distributed.run --nproc-per-node 2 scripts/diag_migration.py [ncell] [nsteps]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pylamp_b200 import _lib, driver, migrate, setups  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctx = _lib.default_context(local)
ctx.init_comm()
nx, L, tr_x, cols, opts = setups.convection_device(ncell=ncell, per_side=4, device="cuda:%d" % local, rank=rank, world=world)
s = driver.State(nx, L, tr_x, cols, device=local)
o = driver.Options(**opts)
o.marker_ownership, o.slab_local, o.resort_every = "slab", True, 16
o.stokes_rtol, o.stokes_params = bench.DEFAULTS["stokes_rtol"], bench.stokes_params()
o.heat_rtol = bench.DEFAULTS["heat_rtol"]
b = migrate.slab_bounds(nx[0] - 1, world)
dz = L[0] / ncell
for it in range(nsteps):
    x0 = s.tr_x.clone()
    M0 = x0.shape[0]
    driver.timestep(s, o, want_kelem=False)
    row = torch.floor(s.tr_x[:, 0] / dz).long()
    stray = int(((row < b[rank]) | (row >= b[rank + 1])).sum().item())
    disp = float((s.tr_x[:min(M0, s.tr_x.shape[0])] - x0[:min(M0, s.tr_x.shape[0])]).abs().max().item()) / dz if s.tr_x.shape[0] == M0 else -1.0
    zmin, zmax = float(s.tr_x[:, 0].min().item()) / dz, float(s.tr_x[:, 0].max().item()) / dz
    print("step %d rank %d limiter %s dt %.3e max|dx| %.4f cells  rows [%d,%d) z in [%.4f, %.4f] strays after migration %d  %s  iters %d"
          % (it + 1, rank, s.limiter, s.tstep, disp, b[rank], b[rank + 1], zmin, zmax, stray, s.stats.get("migrated"), s.stats["stokes_iters"]), flush=True)
dist.destroy_process_group()
