"""Multi-GPU parity (torchrun, one process per GPU): the distributed timestep -- marker-parallel
ranks, all-reduced node sums, z-slab Stokes solve -- versus the oracle's single-process loop body.
  torchrun --nproc-per-node 2 scripts/multi_gpu_driver_check.py [ncell] [nsteps] [index|slab] [reduce]
With "slab" every rank starts with the markers of its own cell rows and markers migrate between the
slabs after every step (pylamp_b200/migrate.py); markers are then matched with the oracle's by id."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pylamp_oracle as O
from pylamp_b200 import _lib, driver, setups

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ownership = sys.argv[3] if len(sys.argv) > 3 else "index"
slab_reduce = len(sys.argv) > 4 and sys.argv[4] == "reduce"      # slab only: slabgrid.py instead of the all-reduce
slab_local = len(sys.argv) > 4 and sys.argv[4] == "local"        # slab only: slab-local grid fields (no full-plane collective)
host_group = dist.new_group(backend="gloo")          # object gathers of the id-matched comparison
ctx = _lib.default_context(local)
ctx.init_comm()
nx, L, tr_x, tr_f, opts = setups.convection(ncell=ncell)
M = tr_x.shape[0]
if ownership == "slab":
    from pylamp_b200 import migrate
    b = migrate.slab_bounds(ncell, world)
    ie = np.clip(np.floor(ncell * tr_x[:, 0] / L[0]).astype(int), 0, ncell - 1)
    sel = np.nonzero(np.searchsorted(b[1:-1], ie, side="right") == rank)[0]
else:
    sel = np.arange((rank * M) // world, ((rank + 1) * M) // world)   # contiguous share of the markers
sg = driver.State(nx, L, tr_x[sel], tr_f[sel], device=local)
og = driver.Options(marker_ownership=ownership, slab_reduce=slab_reduce, slab_local=slab_local, **opts)
if rank == 0:
    so, oo = O.State(nx, L, tr_x.copy(), tr_f.copy()), O.Options(solve=O.solve_refined, **opts)
ok = True
for it in range(nsteps):
    driver.timestep(sg, og)
    # this rank's markers matched with the oracle's by id (they change slot and rank when they migrate)
    ids = sg.cols[O.TR__ID].cpu().numpy().astype(np.int64)
    mine = [None] * world
    dist.gather_object((ids, sg.tr_x.cpu().numpy(), sg.cols[O.TR_TMP].cpu().numpy()), mine if rank == 0 else None,
                       dst=0, group=host_group)
    if ownership == "slab":
        ok = ok and migrate.check_ownership(sg) == 0
    # slab-local fields: every rank holds its own rows (+ halo rows); put the pieces together for the comparison
    F = {k: driver.full_field(sg, f) for k, f in (("vz", sg.newvel[0]), ("vx", sg.newvel[1]), ("P", sg.newpres),
                                                  ("T", sg.newtemp), ("rho", sg.f_rho))}
    count = sg.count.clone()
    if slab_local:
        dist.all_reduce(count)
        # the halo rows must hold the neighbours' values: compare them with the assembled fields
        i0, i1, h = sg.ctx.slab
        lo, hi = max(i0 - h, 0), min(i1 + h, nx[0])
        for k, f in (("vz", sg.newvel[0]), ("T", sg.newtemp), ("rho", sg.f_rho), ("etas", sg.f_etas)):
            ref = F[k] if k in F else driver.full_field(sg, f)
            if not torch.equal(f[lo:hi], ref[lo:hi]):
                print("rank", rank, "halo rows of", k, "differ from the owners' rows", flush=True)
                ok = False
    if rank == 0:
        O.timestep(so, oo)
        all_ids = np.concatenate([m[0] for m in mine])
        ok = ok and np.array_equal(np.sort(all_ids), np.arange(M))          # nobody lost, nobody duplicated
        gx, gT = np.concatenate([m[1] for m in mine]), np.concatenate([m[2] for m in mine])
        rel = lambda a, b: float(np.linalg.norm(a.cpu().numpy() - b) / np.linalg.norm(b))
        e = {"vz": rel(F["vz"], so.newvel[0]), "vx": rel(F["vx"], so.newvel[1]),
             "P": rel(F["P"], so.newpres), "T": rel(F["T"], so.newtemp), "rho": rel(F["rho"], so.f_rho),
             "x": float(np.linalg.norm(gx - so.tr_x[all_ids]) / np.linalg.norm(so.tr_x)),
             "Tm": float(np.linalg.norm(gT - so.tr_f[all_ids, O.TR_TMP]) / np.linalg.norm(so.tr_f[:, O.TR_TMP])),
             "count": int(np.abs(count.cpu().numpy() - so.count).max())}
        print("step", it + 1, "world", world, ownership, "local" if slab_local else ("reduce" if slab_reduce else ""), sg.stats, {k: ("%.1e" % v if k != "count" else v) for k, v in e.items()}, flush=True)
        ok = ok and all(e[k] < 1e-8 for k in ("vz", "vx", "P", "T")) and e["x"] < 1e-10 and e["rho"] < 1e-10
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)        # rank 0 holds the parity verdict, every rank its ownership check
ok = bool(flag.item())
dist.destroy_process_group()
if rank == 0:
    print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
sys.exit(0 if flag.item() else 1)
