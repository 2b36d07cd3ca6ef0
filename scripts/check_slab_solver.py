"""Multi-GPU check (run under torchrun, one process per GPU): the slab-distributed Stokes solve
versus the single-GPU solve of the same system, plus timing.
  torchrun --nproc-per-node 2 scripts/test_slab_solver.py 512"""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from pylamp_b200 import _lib, pylamp_stokes as S
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) + 1
params = dict(kv.split('=') for kv in sys.argv[2:])
rtol = float(params.pop('rtol', 1e-11))
dev = torch.device('cuda', local)
g = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
gm = (g[1:] + g[:-1]) / 2
gm = torch.cat([gm, gm[-1:] + (gm[-1] - gm[-2])])
def fields(z, x):
    T = 273 + 1350 * z + 0.05 * 1350 * torch.sin(np.pi * z) * torch.cos(np.pi * x)
    eta = torch.clamp(1e20 * torch.exp(120e3 / (8.31446 * T) - 120e3 / (8.31446 * 1623)), 1e17, 1e23)
    return eta, 3300 / (3.5e-5 * (T - 1623) + 1)
zs, xs = torch.meshgrid(g, g, indexing='ij'); zc, xc = torch.meshgrid(gm, gm, indexing='ij')
etas, rho = fields(zs, xs); etan, _ = fields(zc, xc)
grid = [g.cpu().numpy() * 1e6, g.cpu().numpy() * 1e6]
ctx = _lib.default_context(local)
# reference: single-GPU solve (no communicator yet), on every rank
A1 = S.StokesOperator([n, n], grid, etas, etan, rho, [1, 1, 1, 1], ctx=ctx)
for k, v in params.items(): A1.set_param(k, float(v))
torch.cuda.synchronize(); t = time.time()
x1 = A1.solve(None, rtol=rtol, maxit=400)
torch.cuda.synchronize(); t1 = time.time() - t
it1 = A1.iterations
A1.close()
ctx.init_comm()
A = S.StokesOperator([n, n], grid, etas, etan, rho, [1, 1, 1, 1], ctx=ctx)
for k, v in params.items(): A.set_param(k, float(v))
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize(); t = time.time()
    x = A.solve(None, rtol=rtol, maxit=400)
    torch.cuda.synchronize(); tn = time.time() - t
    if rank == 0: print('   slab rep', rep, A.stats, '%.3fs' % tn, flush=True)
err = [float(torch.linalg.norm(x[k::3] - x1[k::3]) / torch.linalg.norm(x1[k::3])) for k in range(3)]
print("rank %d/%d n=%d: single-GPU %d its %.3fs | slab %d its %.3fs relres %.2e | rel diff vz,vx,P %s" %
      (rank, world, n, it1, t1, A.iterations, tn, A.relres, ["%.1e" % e for e in err]), flush=True)
assert max(err) < 1e-7, err
dist.destroy_process_group()
