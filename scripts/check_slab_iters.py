"""Does the slab-distributed Stokes solve need the same number of iterations as the single-GPU one?
torchrun, 2+ GPUs:  python -m torch.distributed.run --nproc-per-node 2 scripts/check_slab_iters.py [ncell=2048] [nsteps=10] [shift_cells=0.3]
Steps the analytic C4-type fields (setups.convection_fields) through nsteps time levels with bench.py's solver
settings, first on every rank alone, then slab-distributed, and prints iterations / relres per step for both."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pylamp_b200 import _lib, pylamp_stokes as S, setups  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
shift = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3      # cells per step the fields move
dev = torch.device("cuda", local)
ctx = _lib.default_context(local)
cache = {}
for t in range(-nsteps, 1):
    nx, L, grid, gridmp, es, en, rho = setups.convection_fields(ncell, t, shift_cells=shift)
    cache[t] = [torch.as_tensor(a).to(dev) for a in (es, en, rho)]


def run(tag):
    work = [a.clone() for a in cache[-nsteps]]
    A = S.StokesOperator(nx, grid, *work, [1, 1, 1, 1], ctx=ctx)
    A.warn_unconverged = False
    for k, v in bench.stokes_params().items():
        A.set_param(k, v)
    log = []
    for t in range(-nsteps, 1):
        for w, a in zip(work, cache[t]):
            w.copy_(a)
        A.set_coeffs(*work)
        if t == 0 and rank == 0:
            os.environ["PLB_DEBUG_FGMRES"] = "1"
        A.solve(None, rtol=1e-9, maxit=200, raise_on_fail=False)
        os.environ.pop("PLB_DEBUG_FGMRES", None)
        log.append((A.stats["iterations"], "%.2e" % A.stats["relres"]))
    if rank == 0:
        print(tag, log, flush=True)
    A.close()


run("single GPU:")
ctx.init_comm()
run("%d slabs:   " % world)
dist.destroy_process_group()
