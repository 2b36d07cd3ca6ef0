"""Multi-GPU parity (torchrun, one process per GPU): the distributed timestep -- marker-parallel
ranks, all-reduced node sums, z-slab Stokes solve -- versus the oracle's single-process loop body.
  torchrun --nproc-per-node 2 scripts/multi_gpu_driver_check.py [ncell] [nsteps]"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pylamp_oracle as O
from pylamp_b200 import _lib, driver, setups

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = _lib.default_context(local)
ctx.init_comm()
nx, L, tr_x, tr_f, opts = setups.convection(ncell=ncell)
M = tr_x.shape[0]
lo, hi = (rank * M) // world, ((rank + 1) * M) // world          # contiguous share of the markers
sg = driver.State(nx, L, tr_x[lo:hi], tr_f[lo:hi], device=local)
og = driver.Options(**opts)
if rank == 0:
    so, oo = O.State(nx, L, tr_x.copy(), tr_f.copy()), O.Options(solve=O.solve_refined, **opts)
ok = True
for it in range(nsteps):
    driver.timestep(sg, og)
    if rank == 0:
        O.timestep(so, oo)
        rel = lambda a, b: float(np.linalg.norm(a.cpu().numpy() - b) / np.linalg.norm(b))
        e = {"vz": rel(sg.newvel[0], so.newvel[0]), "vx": rel(sg.newvel[1], so.newvel[1]),
             "P": rel(sg.newpres, so.newpres), "T": rel(sg.newtemp, so.newtemp), "rho": rel(sg.f_rho, so.f_rho),
             "x": rel(sg.tr_x, so.tr_x[lo:hi]), "Tm": rel(sg.cols[O.TR_TMP], so.tr_f[lo:hi, O.TR_TMP]),
             "count": int(np.abs(sg.count.cpu().numpy() - so.count).max())}
        print("step", it + 1, "world", world, sg.stats, {k: ("%.1e" % v if k != "count" else v) for k, v in e.items()}, flush=True)
        ok = ok and all(e[k] < 1e-8 for k in ("vz", "vx", "P", "T")) and e["x"] < 1e-10 and e["rho"] < 1e-10
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
if rank == 0:
    print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
sys.exit(0 if flag.item() else 1)
