"""Stokes-solver micro-benchmark on analytic convection-like fields (no markers).
usage: python scripts/bench_solver.py NCELL [key=value ...]   (solver params via set_param)"""
import sys, time, json, numpy as np, torch
sys.path.insert(0, '.')
from pylamp_b200 import _lib, pylamp_stokes as S
from bench import CLASS_NAMES
n = int(sys.argv[1]) + 1
params = dict(kv.split('=') for kv in sys.argv[2:])
reps = int(params.pop('reps', 2))
dev = torch.device('cuda')
g = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
gm = (g[1:] + g[:-1]) / 2
gm = torch.cat([gm, gm[-1:] + (gm[-1] - gm[-2])])
def fields(z, x, t=0.0):
    T = 273 + 1350 * z + 0.05 * 1350 * torch.sin(np.pi * z) * torch.cos(np.pi * (x + t))
    eta = torch.clamp(1e20 * torch.exp(120e3 / (8.31446 * T) - 120e3 / (8.31446 * 1623)), 1e17, 1e23)
    rho = 3300 / (3.5e-5 * (T - 1623) + 1)
    return eta, rho
zs, xs = torch.meshgrid(g, g, indexing='ij'); zc, xc = torch.meshgrid(gm, gm, indexing='ij')
etas, rho = fields(zs, xs); etan, _ = fields(zc, xc)
grid = [g.cpu().numpy() * 1e6, g.cpu().numpy() * 1e6]
A = S.StokesOperator([n, n], grid, etas, etan, rho, [1, 1, 1, 1])
for k, v in params.items(): A.set_param(k, float(v))
ctx = A.ctx
for rep in range(reps):
    if rep > 0:   # slightly shifted fields, like a next time step
        e2, r2 = fields(zs, xs, 1e-3 * rep); en2, _ = fields(zc, xc, 1e-3 * rep)
        etas.copy_(e2); etan.copy_(en2); rho.copy_(r2)
        A.set_coeffs(etas, etan, rho)
    ctx.profile(True)
    torch.cuda.synchronize(); t = time.time()
    x = A.solve(None, rtol=1e-12, maxit=400, raise_on_fail=False)
    torch.cuda.synchronize(); dt = time.time() - t
    prof = ctx.profile_read(); ctx.profile(False)
    print('rep', rep, 'n', n, params, 'iters', A.iterations, 'relres %.2e' % A.relres, 'time %.3fs' % dt)
    tot = 0
    for k, (c, ms, by) in sorted(prof.items()):
        tot += ms
        print('   %-58s %5d  %8.2f ms  %s' % (CLASS_NAMES[k][:58], c, ms, ('%.0f GB/s' % (by / ms / 1e6)) if by else ''))
    print('   profiled total %.1f ms of %.1f ms' % (tot, dt * 1e3))
