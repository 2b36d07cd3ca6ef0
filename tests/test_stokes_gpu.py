"""GPU parity of the Stokes path through the C-ABI: the matrix-free operator versus the oracle's
assembled matrix (pinned bit-exactly to the reference's makeStokesMatrix), and the GPU solver
versus the oracle's direct solve.

Tolerances.  Operator: 1e-13 relative (summation order only).  Solve: north_star asks for 1e-8
relative L2 on vz, vx, P~ against scipy's direct solution.  A sparse direct solve is itself only
reproducible to cond(A)*eps: for every case the test measures the oracle's own
noise floor -- the distance between raw `spsolve` and `spsolve` + one refinement step -- and
asserts   err <= max(1e-8, 3 * floor)   against the refined solution.  For the smooth / moderate
contrast cases the floor is below 1e-8 and the bare 1e-8 bound applies.
"""
import numpy as np
import pytest
import torch

from oracle import pylamp_oracle as O
from pylamp_b200 import setups

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from pylamp_b200 import pylamp_stokes
    return pylamp_stokes


def _comp_err(x, ref):
    return [np.linalg.norm(x[k::3] - ref[k::3]) / np.linalg.norm(ref[k::3]) for k in range(3)]


def test_operator_matches_reference_matrix(S, golden_kernels):
    g = golden_kernels
    nx = list(g["nx"])
    grid = [g["st_gz"], g["st_gx"]]          # non-uniform rectilinear spacing
    rng = np.random.default_rng(0)
    for bc in g["st_bc"]:
        Aref, rref = O.makeStokesMatrix(nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], list(bc))
        A, rhs = S.makeStokesMatrix(nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], list(bc))
        assert A.shape == Aref.shape
        assert np.allclose(A.scaling, O.stokes_scaling(grid, g["st_etas"], g["st_etan"]), rtol=1e-15)
        assert np.allclose(rhs, rref, rtol=1e-15, atol=0)
        for _ in range(3):
            x = rng.normal(size=A.shape[0])
            y, yref = A @ x, Aref @ x
            assert np.linalg.norm(y - yref) <= 1e-13 * np.linalg.norm(yref), list(bc)
            assert np.allclose(y, yref, rtol=1e-11, atol=1e-12 * np.abs(yref).max())


def test_operator_large_random_viscosity(S):
    rng = np.random.default_rng(1)
    nx, L = [97, 131], [1.0, 2.0]
    grid = O.make_grids(nx, L)[0]
    etas, etan = 10 ** rng.uniform(-3, 3, nx), 10 ** rng.uniform(-3, 3, nx)
    rho = rng.uniform(1, 2, nx)
    Aref, rref = O.makeStokesMatrix(nx, grid, etas, etan, rho, [1, 1, 1, 1])
    A, rhs = S.makeStokesMatrix(nx, grid, etas, etan, rho, [1, 1, 1, 1])
    x = rng.normal(size=A.shape[0])
    assert np.allclose(rhs, rref, rtol=1e-15, atol=0)
    assert np.linalg.norm(A @ x - Aref @ x) <= 1e-13 * np.linalg.norm(Aref @ x)


def test_unsupported_bcs_raise(S, golden_kernels):
    g = golden_kernels
    nx = list(g["nx"])
    grid = [g["st_gz"], g["st_gx"]]
    with pytest.raises(Exception, match="not supported"):
        S.makeStokesMatrix(nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], [1, 0, 1, 1])
    with pytest.raises(Exception, match="surface stabilization"):
        S.makeStokesMatrix(nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], [1, 1, 1, 1], surfstab=True)
    with pytest.raises(Exception, match="FLOWTHRU"):
        S.makeStokesMatrix(nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], [4, 1, 1, 1])


def _solve_case(S, nx, grid, etas, etan, rho, bc, wide=1, maxit=600, label="", gcr_m=None):
    from pylamp_b200 import solve
    Aref, rref = O.makeStokesMatrix(nx, grid, etas, etan, rho, bc)
    xref = O.solve_refined(Aref, rref)
    xraw = O.spsolve(Aref, rref)
    floor = _comp_err(xraw, xref)
    A, rhs = S.makeStokesMatrix(nx, grid, etas, etan, rho, bc)
    A.set_param("coarsen_wide", wide)
    if gcr_m:
        A.set_param("gcr_m", gcr_m)
    x = solve.spsolve(A, rhs, maxit=maxit)
    err = _comp_err(x, xref)
    res = np.linalg.norm(rref - Aref @ x) / np.linalg.norm(rref)
    print("%s %s: iters %d  err(vz,vx,P) %s  oracle floor %s  residual %.1e" %
          (label, nx, A.iterations, ["%.1e" % e for e in err], ["%.1e" % e for e in floor], res))
    for e, f in zip(err, floor):
        assert e <= max(1e-8, 3 * f), (label, err, floor)
    # ghosts and anchor exactly as the reference leaves them
    vel, p = S.x2vp(x, nx)
    assert np.all(vel[0][:, -1] == 0) and np.all(vel[1][-1, :] == 0)
    assert np.all(p[-1, :] == 0) and np.all(p[:, -1] == 0) and p[3, 2] == 0
    return A, x, xref


def test_solve_golden_small_nonuniform(S, golden_kernels):
    g = golden_kernels
    nx = list(g["nx"])
    grid = [g["st_gz"], g["st_gx"]]
    for bc in g["st_bc"]:
        _solve_case(S, nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], list(bc), label="golden bc=%s" % list(bc),
                    gcr_m=200)      # node-wise random viscosity: Krylov needs a long recurrence


@pytest.mark.parametrize("n", [65, 129, 257])
def test_solve_solcx(S, n):
    """BASELINE.json configs[1]: SolCx-type, viscosity jump 1e6 (2x2 coarsening is exact here)."""
    nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
    _solve_case(S, nx, grid, etas, etan, rho, [1, 1, 1, 1], wide=0, label="solcx")


def test_solve_solcx_default_coarsening(S):
    nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(65)
    _solve_case(S, nx, grid, etas, etan, rho, [1, 1, 1, 1], wide=1, label="solcx-wide")


def test_solve_rayleigh_taylor_fields(S):
    nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(129)
    zs, xs = np.meshgrid(grid[0], grid[1], indexing="ij")
    zc, xc = np.meshgrid(gridmp[0], gridmp[1], indexing="ij")
    up = lambda z, x: z < 0.5 + 0.02 * np.cos(np.pi * x)
    etas, etan = np.where(up(zs, xs), 1e21, 1e20), np.where(up(zc, xc), 1e21, 1e20)
    rho = np.where(up(zs, xs), 3300.0, 3200.0)
    for bc in ([1, 1, 1, 1], [0, 1, 0, 1]):
        _solve_case(S, nx, grid, etas, etan, rho, bc, label="RT bc=%s" % bc)


def test_solve_arrhenius_convection_fields(S):
    n = 129
    nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
    zs, xs = np.meshgrid(grid[0], grid[1], indexing="ij")
    zc, xc = np.meshgrid(gridmp[0], gridmp[1], indexing="ij")
    T = lambda z, x: 273 + 1350 * z + 0.05 * 1350 * np.sin(np.pi * z) * np.cos(np.pi * x)
    eta = lambda t: np.clip(1e20 * np.exp(120e3 / (8.31446 * t) - 120e3 / (8.31446 * 1623)), 1e17, 1e23)
    _solve_case(S, nx, grid, eta(T(zs, xs)), eta(T(zc, xc)), 3300 / (3.5e-5 * (T(zs, xs) - 1623) + 1),
                [1, 1, 1, 1], label="arrhenius")


def test_solve_c1_shipped_fields(S):
    """BASELINE.json configs[0]: the fields of pylamp2.py as shipped (201x41 nodes, eta 1e2/1e12)."""
    nx, L, tr_x, tr_f, opts = setups.c1_shipped(1234)
    s = O.State(nx, L, tr_x, tr_f)
    O.update_properties(tr_f, False, False, 1623, 1e17, 1e23)
    O.trac2grid(tr_x, tr_f[:, [O.TR_RHO, O.TR_ETA]], s.mesh, s.grid, [s.f_rho, s.f_etas], nx, avgscheme=[5, 6])
    O.trac2grid(tr_x, tr_f[:, [O.TR_ETA]], s.meshmp, s.gridmp, [s.f_etan], nx, avgscheme=[2])
    _solve_case(S, nx, s.grid, s.f_etas, s.f_etan, s.f_rho, [1, 1, 1, 1], label="C1")


def test_device_resident_solve_and_x2vp(S):
    nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(65, eta_right=1e3)
    dev = lambda a: torch.as_tensor(a).cuda()
    A, rhs = S.makeStokesMatrix(nx, grid, dev(etas), dev(etan), dev(rho), [1, 1, 1, 1])
    assert rhs.is_cuda
    x = A.solve(rhs)
    assert x.is_cuda
    vel, p = S.x2vp(x, nx)
    xh = x.cpu().numpy()
    velh, ph = S.x2vp(xh, nx)
    assert np.array_equal(vel[0].cpu().numpy(), velh[0]) and np.array_equal(p.cpu().numpy(), ph)
    Aref, rref = O.makeStokesMatrix(nx, grid, etas, etan, rho, [1, 1, 1, 1])
    assert max(_comp_err(xh, O.solve_refined(Aref, rref))) < 1e-8
