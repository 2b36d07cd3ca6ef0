"""GPU parity of the energy path: matrix-free heat operator versus the oracle's assembled matrix
(bit-exactly pinned to the reference's makeDiffusionMatrix) and the GPU solve versus spsolve.
Tolerance: operator 1e-13, temperature 1e-8 relative L2 (north_star) -- measured ~1e-12."""
import numpy as np
import pytest

from oracle import pylamp_oracle as O

pytestmark = pytest.mark.gpu


def test_operator_and_solve_golden(golden_kernels):
    from pylamp_b200 import pylamp_diff as Dm, solve
    g = golden_kernels
    nx = list(g["nx"])
    grid, gridmp = [g["st_gz"], g["st_gx"]], [g["df_gmz"], g["df_gmx"]]
    rng = np.random.default_rng(0)
    for bc in g["df_bc"]:
        args = (nx, grid, gridmp, g["df_T"], list(g["df_k"]), g["df_Cp"], g["st_rho"], g["df_H"], list(bc),
                list(g["df_bcval"]), float(g["df_tstep"]))
        Aref, rref = O.makeDiffusionMatrix(*args)
        A, rhs = Dm.makeDiffusionMatrix(*args)
        assert np.allclose(rhs, rref, rtol=1e-14, atol=0)
        x = rng.normal(size=A.shape[0])
        assert np.linalg.norm(A @ x - Aref @ x) <= 1e-13 * np.linalg.norm(Aref @ x)
        T = solve.spsolve(A, rhs)
        Tref = O.spsolve(Aref, rref)
        err = np.linalg.norm(T - Tref) / np.linalg.norm(Tref)
        print("bc", list(bc), "iters", A.iterations, "err %.1e" % err)
        assert err < 1e-8
        assert Dm.x2t(T, nx).shape == tuple(nx)


@pytest.mark.parametrize("n", [129, 257])
def test_solve_convection_like(n):
    """dt at the driver's heat limit (pylamp2.py:339-343), T 273..1623 K, variable k/rho/cp."""
    from pylamp_b200 import pylamp_diff as Dm
    rng = np.random.default_rng(n)
    nx, L = [n, n], [1e6, 1e6]
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    z = mesh[0] / L[0]
    T = 273 + 1350 * z + 30 * rng.normal(size=nx)
    k = [rng.uniform(2.5, 4.0, nx), rng.uniform(2.5, 4.0, nx)]
    cp, rho, H = rng.uniform(1000, 1250, nx), rng.uniform(2900, 3300, nx), rng.uniform(0, 1e-9, nx)
    dx = L[0] / (n - 1)
    dt = 0.67 * dx ** 2 / np.max(2 * k[0] / (rho * cp))
    bc, bcv = [0, 1, 0, 1], [273, 0, 1623, 0]
    Aref, rref = O.makeDiffusionMatrix(nx, grid, gridmp, T, k, cp, rho, H, bc, bcv, dt)
    A, rhs = Dm.makeDiffusionMatrix(nx, grid, gridmp, T, k, cp, rho, H, bc, bcv, dt)
    Tg, Tref = A.solve(rhs), O.spsolve(Aref, rref)
    err = np.linalg.norm(Tg - Tref) / np.linalg.norm(Tref)
    derr = np.linalg.norm((Tg - T.ravel()) - (Tref - T.ravel())) / np.linalg.norm(Tref - T.ravel())
    print("n", n, "iters", A.iterations, "err %.1e" % err, "increment err %.1e" % derr)
    assert err < 1e-8 and derr < 1e-6
