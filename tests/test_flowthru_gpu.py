"""the flow-through x = 0 wall the reference actually solves -- BC_TYPE_FLOWTHRU|FREESLIP (= 5):
dvx/dx = 0 on the wall rows (pylamp_stokes.py:268-273), free slip for vz (:249-255), pressure anchor moved to cell
(nz/2, 0) (:525-551), markers that leave through the wall are removed instead of fenced (pylamp2.py:565-581) and the
ghost ring of the cell-centre velocities copies vx (:521-522).  Operator, solve and time loop against the oracle, whose
assembly and solve are pinned to the reference for these walls (tests/golden/flowthru.npz)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import pylamp_oracle as O
from pylamp_b200 import setups

pytestmark = [pytest.mark.gpu]


def _comp_err(x, ref):
    return [np.linalg.norm(x[k::3] - ref[k::3]) / np.linalg.norm(ref[k::3]) for k in range(3)]


def test_operator_matches_reference_matrix(golden_kernels):
    from pylamp_b200 import pylamp_stokes as S
    g, f = golden_kernels, np.load(os.path.join(GOLDEN, "flowthru.npz"))
    nx = list(g["nx"])
    n = 3 * nx[0] * nx[1]
    grid = [g["st_gz"], g["st_gx"]]
    rng = np.random.default_rng(0)
    import scipy.sparse
    for ic, bc in enumerate(f["ft_bc"]):
        key = "ft_%d_" % ic
        Aref = scipy.sparse.csr_matrix((f[key + "val"], (f[key + "row"], f[key + "col"])), shape=(n, n))
        A, rhs = S.makeStokesMatrix(nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], list(bc))
        assert np.allclose(rhs, f[key + "rhs"], rtol=1e-15, atol=0)
        for _ in range(3):
            x = rng.normal(size=n)
            assert np.linalg.norm(A @ x - Aref @ x) <= 1e-13 * np.linalg.norm(Aref @ x)


def test_unsupported_flowthru_walls_raise(golden_kernels):
    from pylamp_b200 import pylamp_stokes as S
    g = golden_kernels
    nx, grid = list(g["nx"]), [g["st_gz"], g["st_gx"]]
    for bc in ([1, 4, 1, 1], [1, 1, 1, 5], [1, 5, 1, 5]):
        with pytest.raises(Exception, match="not supported"):
            S.makeStokesMatrix(nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], bc)
    with pytest.raises(Exception, match="FLOWTHRU"):
        S.makeStokesMatrix(nx, grid, g["st_etas"], g["st_etan"], g["st_rho"], [5, 1, 1, 1])


def test_solve_matches_reference_direct_solve():
    """The 33 x 25 fixture solved by the reference itself, and a 129^2 Arrhenius-type case against the oracle."""
    from pylamp_b200 import pylamp_stokes as S, solve
    f = np.load(os.path.join(GOLDEN, "flowthru.npz"))
    nx = list(f["fs_nx"])
    A, rhs = S.makeStokesMatrix(nx, [f["fs_gz"], f["fs_gx"]], f["fs_etas"], f["fs_etan"], f["fs_rho"], [1, 5, 1, 1])
    A.set_param("gcr_m", 100)
    x = solve.spsolve(A, rhs, maxit=800)
    err = _comp_err(x, f["fs_x"])
    print("flow-through 33x25: iters", A.iterations, "err", ["%.1e" % e for e in err], A.stats)
    assert max(err) < 1e-8
    vel, p = S.x2vp(x, nx)
    assert p[nx[0] // 2, 0] == 0 and np.all(vel[1][:-1, 0] == vel[1][:-1, 1])      # anchor; dvx/dx = 0 on the wall
    assert np.abs(vel[1][:, 0]).max() > 0.1 * np.abs(vel[0]).max()                 # flow does go through the wall
    n = 129
    nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
    zs, xs = np.meshgrid(grid[0], grid[1], indexing="ij")
    zc, xc = np.meshgrid(gridmp[0], gridmp[1], indexing="ij")
    T = lambda z, x: 273 + 1350 * z + 0.15 * 1350 * np.exp(-((z - 0.5) ** 2 + (x - 0.15) ** 2) / 0.02)
    eta = lambda t: np.clip(1e20 * np.exp(120e3 / (8.31446 * t) - 120e3 / (8.31446 * 1623)), 1e18, 1e23)
    etas, etan, rho = eta(T(zs, xs)), eta(T(zc, xc)), 3300 / (3.5e-5 * (T(zs, xs) - 1623) + 1)
    Aref, rref = O.makeStokesMatrix(nx, grid, etas, etan, rho, [1, 5, 1, 1])
    xref, xraw = O.solve_refined(Aref, rref), O.spsolve(Aref, rref)
    floor = _comp_err(xraw, xref)
    A, rhs = S.makeStokesMatrix(nx, grid, etas, etan, rho, [1, 5, 1, 1])
    x = solve.spsolve(A, rhs, maxit=800)
    err = _comp_err(x, xref)
    print("flow-through 129^2: iters", A.iterations, "err", ["%.1e" % e for e in err], "floor", ["%.1e" % e for e in floor])
    for e, fl in zip(err, floor):
        assert e <= max(1e-8, 3 * fl)


def test_time_loop_with_flowthru_wall_vs_oracle():
    """Three thermo-mechanical steps with an open x = 0 wall: fields, time step, the markers removed at the wall
    (same ids survive) and the survivors' positions against the oracle's loop."""
    from pylamp_b200 import driver
    nx, L, tr_x, tr_f, opts = setups.convection(ncell=32)
    # a hot anomaly next to the open wall so that material flows through it
    zn, xn = tr_x[:, 0] / L[0], tr_x[:, 1] / L[1]
    tr_f[:, O.TR_TMP] += 300 * np.exp(-((zn - 0.5) ** 2 + (xn - 0.1) ** 2) / 0.02)
    opts = dict(opts, bcstokes=[1, 5, 1, 1])
    so, oo = O.State(nx, L, tr_x.copy(), tr_f.copy()), O.Options(solve=O.solve_refined, **opts)
    sg, og = driver.State(nx, L, tr_x, tr_f), driver.Options(**opts)
    rel = lambda a, b: np.linalg.norm(a.cpu().numpy() - b) / np.linalg.norm(b)
    removed = 0
    for it in range(3):
        O.timestep(so, oo)
        driver.timestep(sg, og)
        e = {"vz": rel(sg.newvel[0], so.newvel[0]), "vx": rel(sg.newvel[1], so.newvel[1]), "P": rel(sg.newpres, so.newpres),
             "T": rel(sg.newtemp, so.newtemp)}
        print("step", it + 1, sg.stats, {k: "%.1e" % v for k, v in e.items()}, "markers", sg.ntrac, so.tr_x.shape[0])
        assert all(v < 1e-8 for v in e.values()), e
        assert abs(sg.tstep - so.tstep) <= 1e-9 * so.tstep
        assert sg.ntrac == so.tr_x.shape[0]
        removed += sg.stats.get("removed", 0)
        ids_g = sg.cols[O.TR__ID].cpu().numpy().astype(np.int64)
        ids_o = so.tr_f[:, O.TR__ID].astype(np.int64)
        og_, oo_ = np.argsort(ids_g), np.argsort(ids_o)
        assert np.array_equal(ids_g[og_], ids_o[oo_])
        assert np.allclose(sg.tr_x.cpu().numpy()[og_], so.tr_x[oo_], rtol=1e-10, atol=1e-10 * L[0])
        assert int(sg.count.sum().item()) == sg.ntrac and np.array_equal(np.sort(sg.kelem.cpu().numpy()), np.sort(so.kelem))
    assert removed > 0          # the wall is open: markers did leave
