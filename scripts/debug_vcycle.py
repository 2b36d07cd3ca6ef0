"""Debug: compare the CUDA V-cycle / dense coarse solve with the NumPy prototype (GPU box)."""
import sys, numpy as np, torch, scipy.sparse.linalg as spla
sys.path.insert(0, '.')
from oracle import mg_prototype as P, pylamp_oracle as O
from pylamp_b200 import pylamp_stokes as S, setups

def planar_from_reduced(lv, y, nz, nxx):
    full = lv.E @ np.concatenate([y, np.zeros(lv.np_)]) if hasattr(lv,'np_') else None
    out = np.zeros((2, nz*nxx))
    out[0] = full[0::3]; out[1] = full[1::3]
    return out

def reduced_from_planar(lv, b2, nz, nxx):
    I = lv.I[:lv.nv]
    node, eq = I//3, I%3
    return b2.reshape(2,-1)[eq, node]

def run(name, nx, grid, etas, etan, rho, bc, wide):
    nz, nxx = nx
    A, rhs = S.makeStokesMatrix(nx, grid, etas, etan, rho, bc)
    A.set_param("coarsen_wide", wide)
    mg = P.MG2(nz, nxx, grid[0], grid[1], etas, etan, rho, bc, ms='fw', mn='4x4' if wide else '2x2', smoother='cheb', nu=3)
    lv = mg.levels[0]
    rng = np.random.default_rng(0)
    y = rng.normal(size=lv.nv)
    b = lv.K @ y
    # planar rhs (zero off-rows)
    I = lv.I[:lv.nv]; node, eq = I//3, I%3
    b2 = np.zeros((2, nz*nxx)); b2[eq, node] = b
    xg = A.vcycle(b2).cpu().numpy().reshape(2,-1)
    xr = xg[eq, node]
    xp = mg.vcycle(0, b)
    print(name, 'levels', [(l.nz,l.nxx) for l in mg.levels], 'lmax', np.round(mg.lmax,3))
    print('  |V_gpu b - V_proto b|/|V_proto b| = %.3e   |V_gpu b - y|/|y| = %.3e  |V_proto b - y|/|y| = %.3e' % (
        np.linalg.norm(xr-xp)/np.linalg.norm(xp), np.linalg.norm(xr-y)/np.linalg.norm(y), np.linalg.norm(xp-y)/np.linalg.norm(y)))
    # slaves consistent?
    full = lv.E @ np.concatenate([xr, np.zeros(lv.np_)])
    print('  slave mismatch', np.abs(full[0::3]-xg[0]).max(), np.abs(full[1::3]-xg[1]).max())

g = np.load('tests/golden/kernels.npz')
run('golden', list(g['nx']), [g['st_gz'], g['st_gx']], g['st_etas'], g['st_etan'], g['st_rho'], [1,1,1,1], 1)
run('golden-noslip', list(g['nx']), [g['st_gz'], g['st_gx']], g['st_etas'], g['st_etan'], g['st_rho'], [0,1,0,1], 1)
for n in (17, 33, 65):
    nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
    run('const%d'%n, nx, grid, np.ones_like(etas), np.ones_like(etan), rho, [1,1,1,1], 1)
    run('solcx%d-wide'%n, nx, grid, etas, etan, rho, [1,1,1,1], 1)
    run('solcx%d-narrow'%n, nx, grid, etas, etan, rho, [1,1,1,1], 0)
