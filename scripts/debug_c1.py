import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from oracle import pylamp_oracle as O
from pylamp_b200 import pylamp_stokes as S, setups
nx, L, tr_x, tr_f, opts = setups.c1_shipped(1234)
s = O.State(nx, L, tr_x, tr_f)
O.update_properties(tr_f, False, False, 1623, 1e17, 1e23)
O.trac2grid(tr_x, tr_f[:, [O.TR_RHO, O.TR_ETA]], s.mesh, s.grid, [s.f_rho, s.f_etas], nx, avgscheme=[5, 6])
O.trac2grid(tr_x, tr_f[:, [O.TR_ETA]], s.meshmp, s.gridmp, [s.f_etan], nx, avgscheme=[2])
Aref, rref = O.makeStokesMatrix(nx, s.grid, s.f_etas, s.f_etan, s.f_rho, [1,1,1,1])
xref = O.solve_refined(Aref, rref)
err = lambda x: ["%.1e" % (np.linalg.norm(x[k::3]-xref[k::3])/np.linalg.norm(xref[k::3])) for k in range(3)]
for params in [dict(), dict(reorth=1), dict(gcr_m=100), dict(gcr_m=100, reorth=1), dict(gcr_m=200, reorth=1), dict(coarsen_wide=0, gcr_m=100, reorth=1), dict(nu=4, gcr_m=100, reorth=1), dict(cheb_ratio=20, nu=5, gcr_m=100, reorth=1)]:
    A, rhs = S.makeStokesMatrix(nx, s.grid, s.f_etas, s.f_etan, s.f_rho, [1,1,1,1])
    for k, v in params.items(): A.set_param(k, v)
    torch.cuda.synchronize(); t = time.time()
    x = A.solve(rhs, maxit=800, raise_on_fail=False)
    torch.cuda.synchronize(); dt = time.time() - t
    print(params, 'iters', A.iterations, 'relres %.1e' % A.relres, 'err', err(x), '%.2fs' % dt, flush=True)
