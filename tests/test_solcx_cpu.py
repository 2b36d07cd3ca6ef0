"""The analytic SolCx-type solution (pylamp_b200/solcx.py) against the oracle's direct solve of the reference's
discrete system (BASELINE.json configs[1], CPU-sized grids): it satisfies the boundary and interface conditions,
and the discrete solution converges to it at first order (the reference puts the viscosity jump on a node line)."""
import numpy as np

from oracle import pylamp_oracle as O
from pylamp_b200 import setups, solcx


def test_analytic_solution_conditions():
    f = solcx.solution(eta_right=1e6)
    z = np.linspace(0.05, 0.95, 7)
    for xw in (0.0, 1.0):                                  # no normal flow, no shear stress on the x-walls
        h = 1e-6
        xin = xw + (h if xw == 0 else -h)
        assert np.all(np.abs(f(z, np.full_like(z, xw))[1]) < 1e-12)
        dvz = (f(z, np.full_like(z, xin))[0] - f(z, np.full_like(z, xw))[0]) / (xin - xw)
        assert np.all(np.abs(dvz) < 1e-4 * np.abs(f(z, np.full_like(z, 0.3))[0]).max())     # O(h) one-sided difference, h = 1e-6
    e = 1e-9                                               # velocities continuous across the viscosity jump
    a, b = f(z, np.full_like(z, 0.5 - e)), f(z, np.full_like(z, 0.5 + e))
    vmax = np.abs(f(z, np.full_like(z, 0.3))[0]).max()
    assert np.allclose(a[0], b[0], rtol=0, atol=1e-6 * vmax) and np.allclose(a[1], b[1], rtol=0, atol=1e-6 * vmax)
    zz = np.array([0.0, 1.0])                              # no normal flow on the z-walls
    assert np.all(np.abs(f(zz, np.array([0.3, 0.7]))[0]) < 1e-15)


def test_oracle_converges_to_analytic_at_first_order():
    errs = []
    for n in (33, 65, 129):
        nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
        A, b = O.makeStokesMatrix(nx, grid, etas, etan, rho, [1, 1, 1, 1])
        vel, p = O.x2vp(O.solve_refined(A, b), nx)
        errs.append(solcx.errors(nx, grid, gridmp, vel[0], vel[1], p, O.stokes_scaling(grid, etas, etan)[0]))
    for a, b in zip(errs[:-1], errs[1:]):
        order = [np.log2(x / y) for x, y in zip(a, b)]
        assert all(0.9 < o < 1.2 for o in order), (errs, order)
    assert max(errs[-1]) < 6e-2
