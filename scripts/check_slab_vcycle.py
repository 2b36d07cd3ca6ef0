"""Compare one V-cycle: slab-distributed vs single GPU (torchrun, 2+ GPUs)."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from pylamp_b200 import _lib, pylamp_stokes as S
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) + 1
dev = torch.device('cuda', local)
g = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
gm = (g[1:] + g[:-1]) / 2; gm = torch.cat([gm, gm[-1:] + (gm[-1] - gm[-2])])
def fields(z, x):
    T = 273 + 1350 * z + 0.05 * 1350 * torch.sin(np.pi * z) * torch.cos(np.pi * x)
    return torch.clamp(1e20 * torch.exp(120e3 / (8.31446 * T) - 120e3 / (8.31446 * 1623)), 1e17, 1e23), 3300 / (3.5e-5 * (T - 1623) + 1)
zs, xs = torch.meshgrid(g, g, indexing='ij'); zc, xc = torch.meshgrid(gm, gm, indexing='ij')
etas, rho = fields(zs, xs); etan, _ = fields(zc, xc)
grid = [g.cpu().numpy() * 1e6] * 2
ctx = _lib.default_context(local)
gen = torch.Generator(device=dev); gen.manual_seed(5)
b = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
b[0, 1:n-1, 1:n-2] = torch.randn((n-2, n-3), generator=gen, dtype=torch.float64, device=dev)
b[1, 1:n-2, 1:n-1] = torch.randn((n-3, n-2), generator=gen, dtype=torch.float64, device=dev)
A1 = S.StokesOperator([n, n], grid, etas, etan, rho, [1, 1, 1, 1], ctx=ctx)
x1 = A1.vcycle(b).reshape(2, n, n).clone()
A1.close()
ctx.init_comm()
A = S.StokesOperator([n, n], grid, etas, etan, rho, [1, 1, 1, 1], ctx=ctx)
x = A.vcycle(b).reshape(2, n, n)
d = (x - x1).abs()
rows = d.amax(dim=(0, 2)) / x1.abs().max()
bad = torch.nonzero(rows > 1e-10).flatten().cpu().numpy()
if rank == 0:
    print("n", n, "max rel diff %.2e" % float(rows.max()), "rows differing:", bad[:20], "... count", len(bad))
dist.destroy_process_group()
