"""Design study (CPU, NumPy prototype of the GPU solver: oracle/mg_prototype.py): does the FGMRES +
multigrid solver need the free-surface stabilisation terms (pylamp_stokes.py:422-426) inside the
preconditioner, or is it enough to add them to the outer operator?  Sticky-air free surface with a
cosine topography, dt from the advective criterion of the unstabilised solve (pylamp2.py:364-366).
  python scripts/proto_surfstab.py [n=65] [viscosity contrast=100]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mg_prototype as P, pylamp_oracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65
nx, L = [n, n], [1.0, 1.0]
grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
zs, xs = np.meshgrid(grid[0], grid[1], indexing="ij")
zc, xc = np.meshgrid(gridmp[0], gridmp[1], indexing="ij")
surf = lambda x: 0.2 + 0.03 * np.cos(2 * np.pi * x)           # z grows downwards: air above the surface
eta_rock, eta_air, rho_rock, rho_air = float(sys.argv[2]) if len(sys.argv) > 2 else 1e2, 1.0, 1000.0, 1.0
etas = np.where(zs < surf(xs), eta_air, eta_rock)
etan = np.where(zc < surf(xc), eta_air, eta_rock)
rho = np.where(zs < surf(xs), rho_air, rho_rock)
bc = [1, 1, 1, 1]
A0, b0 = O.makeStokesMatrix(nx, grid, etas, etan, rho, bc)
x0 = O.solve_refined(A0, b0)
(vz, vx), _ = O.x2vp(x0, nx)
dx = L[0] / (n - 1)
dt = 0.67 * dx / np.max([vz, vx])
print("n", n, "max v", float(np.max([vz, vx])), "dt", dt)
mg = P.MG2(n, n, grid[0], grid[1], etas, etan, rho, bc, nu=3)
lv = mg.levels[0]
_, its_plain, _ = P.solve_scaled(mg, tol=1e-10, m=40, maxit=300)
print("unstabilised system, standard preconditioner: iterations", its_plain)
for mult in (1.0, 10.0):
    A1, b1 = O.makeStokesMatrix(nx, grid, etas, etan, rho, bc, surfstab=True, tstep=mult * dt)
    x1 = O.solve_refined(A1, b1)
    change = np.linalg.norm(x1 - x0) / np.linalg.norm(x0)
    Ar1 = (lv.S @ A1.tocsr() @ lv.E).tocsr()
    keep = (lv.Ar, lv.K, lv.Kdiag, lv.G)
    # (a) stabilisation terms in the outer operator only
    lv.Ar = Ar1
    xa, its_a, _ = P.solve_scaled(mg, tol=1e-10, m=40, maxit=300)
    err_a = np.linalg.norm(xa - x1) / np.linalg.norm(x1)
    # (b) also in the finest-level velocity block of the preconditioner (smoother diagonal + residuals)
    lv.K = Ar1[:lv.nv][:, :lv.nv].tocsr()
    lv.Kdiag = lv.K.diagonal()
    xb, its_b, _ = P.solve_scaled(mg, tol=1e-10, m=40, maxit=300)
    err_b = np.linalg.norm(xb - x1) / np.linalg.norm(x1)
    lv.Ar, lv.K, lv.Kdiag, lv.G = keep
    print("dt x %g: solution changes by %.2e; outer-only: %d iterations (err %.1e); + finest level of the "
          "preconditioner: %d iterations (err %.1e)" % (mult, change, its_a, err_a, its_b, err_b))
