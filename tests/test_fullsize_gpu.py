"""Full-size checks at the benchmark configuration (BASELINE.json configs[3]: 4096^2 cells, 16
markers per cell, 2.7e8 markers, 5.0e7 Stokes DOF) through size-independent properties -- the
oracle (SciPy SuperLU) cannot run this size:

  * trac2grid reproduces a constant exactly (partition of unity of the four corner weights), for the
    arithmetic and the geometric weighted mean, on the node target and on a staggered target;
  * grid2trac reproduces a linear field exactly at every marker;
  * RK4 in a uniform velocity field moves every marker by (4/6) dt v -- the reference's 1/6 weights,
    pylamp_trac.py:385;
  * the Stokes operator is linear; the solution of the time loop satisfies b - A x = 0 and the
    discrete continuity equation;
  * the heat system keeps the conductive steady state (linear T(z), uniform k, no sources) fixed;
  * the time loop conserves markers: per-cell counts add up to M, every marker stays inside the
    fence, positions and temperatures stay finite.

Needs ~90 GB of HBM (one B200); skipped on smaller devices.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NCELL = 4096


@pytest.fixture(scope="module")
def big():
    if torch.cuda.get_device_properties(0).total_memory < 120e9:
        pytest.skip("needs a 180 GB device")
    from pylamp_b200 import driver, setups
    nx, L, tr_x, cols, opts = setups.convection_device(ncell=NCELL, per_side=4, device="cuda:0")
    s = driver.State(nx, L, tr_x, cols, device=0)
    o = driver.Options(**opts)
    o.heat_rtol = 1e-11
    o.stokes_params = {"warm_start": 1, "nu": 2}
    yield s, o
    del s
    torch.cuda.empty_cache()


def _relmax(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_marker_kernels_exact_properties(big):
    s, o = big
    from pylamp_b200 import pylamp_trac as T
    from pylamp_b200.pylamp_const import IX, IZ
    ctx, nx, L = s.ctx, s.nx, s.L
    M = s.ntrac
    assert M == 16 * NCELL * NCELL
    # trac2grid of a constant
    c = torch.full((M,), 3.25e20, dtype=torch.float64, device="cuda")
    out = [torch.empty(tuple(nx), dtype=torch.float64, device="cuda") for _ in range(2)]
    T.trac2grid_device(ctx, s.tr_x, [c, c], [5, 6], s.grid, out)
    assert float((out[0] - 3.25e20).abs().max()) <= 1e-13 * 3.25e20
    assert float((out[1] - 3.25e20).abs().max()) <= 5e-12 * 3.25e20          # exp(mean(log)): 47 x the rounding of the mean
    T.trac2grid_device(ctx, s.tr_x, [c], [5], [s.gridmp[IZ], s.grid[IX]], out[:1])
    inner = out[0][:-1, :]                                                   # last row of a z-staggered target is a ghost row
    assert float((inner - 3.25e20).abs().max()) <= 1e-13 * 3.25e20
    del c
    # grid2trac of a linear field
    gz = torch.as_tensor(s.grid[IZ], device="cuda").view(-1, 1)
    gx = torch.as_tensor(s.grid[IX], device="cuda").view(1, -1)
    f = (2.0 + 3.0 * gz / L[IZ] - 1.5 * gx / L[IX]).contiguous()
    got = torch.empty(M, dtype=torch.float64, device="cuda")
    nbad = T.grid2trac_device(ctx, s.tr_x, s.grid, [f], nx, T.INTERP_METHOD_LINEAR, float("nan"), [got])
    assert nbad == 0
    want = 2.0 + 3.0 * s.tr_x[:, 0] / L[IZ] - 1.5 * s.tr_x[:, 1] / L[IX]
    assert float((got - want).abs().max()) < 1e-12
    del got, want, f
    # RK4 in a uniform velocity field
    dz = L[IZ] / (nx[IZ] - 1)
    pre = [s.gridmp[d][0] - (s.gridmp[d][1] - s.gridmp[d][0]) for d in range(2)]
    newgrid = [np.insert(s.gridmp[IZ], 0, pre[IZ]), np.insert(s.gridmp[IX], 0, pre[IX])]
    vz = torch.full((nx[IZ] + 1, nx[IX] + 1), 2.0e-9, dtype=torch.float64, device="cuda")
    vx = torch.full_like(vz, -1.0e-9)
    dt = 0.3 * dz / 2.0e-9
    v, x1 = T.rk4_device(ctx, s.tr_x, newgrid, vz, vx, [nx[IZ] + 1, nx[IX] + 1], dt)
    inside = (s.tr_x[:, 0] > 2 * dz) & (s.tr_x[:, 0] < L[IZ] - 2 * dz) & (s.tr_x[:, 1] > 2 * dz) & \
             (s.tr_x[:, 1] < L[IX] - 2 * dz)
    d = (x1 - s.tr_x)[inside]
    assert float((d[:, 0] - (4.0 / 6.0) * dt * 2.0e-9).abs().max()) < 1e-9 * dz
    assert float((d[:, 1] + (4.0 / 6.0) * dt * 1.0e-9).abs().max()) < 1e-9 * dz


def test_time_loop_conserves_markers_and_solves_the_system(big):
    s, o = big
    from pylamp_b200 import driver
    from pylamp_b200.pylamp_const import EPS, IX, IZ, TR_TMP
    M = s.ntrac
    for _ in range(2):
        driver.timestep(s, o, want_kelem=False)
        assert int(s.count.sum().item()) == M == s.ntrac
        assert float(s.tr_x[:, 0].min()) >= EPS and float(s.tr_x[:, 0].max()) <= s.L[IZ] - EPS
        assert float(s.tr_x[:, 1].min()) >= EPS and float(s.tr_x[:, 1].max()) <= s.L[IX] - EPS
        assert bool(torch.isfinite(s.cols[TR_TMP]).all())
        assert s.stats["stokes_relres"] < 1e-8 and s.stats["stokes_iters"] < 200
    # temperatures stay within the boundary values (+ the 1 % perturbation)
    assert 273.0 - 20 <= float(s.newtemp.min()) and float(s.newtemp.max()) <= 1623.0 + 20
    # b - A x of the last solve, recomputed with the operator kernel, and the discrete continuity equation
    op = s.stokes_op
    b = op.rhs(device=True)
    # the solver returned planar fields; rebuild the interleaved vector the reference layout uses
    x = torch.stack([s.newvel[IZ].reshape(-1), s.newvel[IX].reshape(-1), s.newpres.reshape(-1)], dim=1).reshape(-1)
    r = b - op.dot(x)
    res = float(r.norm() / b.norm())
    vz, vx = s.newvel[IZ], s.newvel[IX]
    dz, dx = s.dx[IZ], s.dx[IX]
    div = (vx[:-1, 1:] - vx[:-1, :-1]) / dx + (vz[1:, :-1] - vz[:-1, :-1]) / dz
    scale = float(torch.maximum(vz.abs().max(), vx.abs().max())) / dx
    divrel = float(div[1:-1, 1:-1].abs().max()) / scale
    print("full-size Stokes: |b-Ax|/|b| = %.2e, max|div v| dx/max|v| = %.2e, iterations %d" %
          (res, divrel, s.stats["stokes_iters"]))
    assert res < 1e-10 and divrel < 1e-10          # measured on B200: 4.4e-13 and 6.7e-14


def test_operators_linear_and_steady_state(big):
    s, o = big
    from pylamp_b200 import pylamp_diff
    from pylamp_b200.pylamp_const import IX, IZ
    if s.stokes_op is None:
        pytest.skip("time-loop test did not run")
    op = s.stokes_op
    n = op.shape[0]
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    lhs = op.dot(2.5 * x - 0.75 * y)
    rhs = 2.5 * op.dot(x) - 0.75 * op.dot(y)
    assert float((lhs - rhs).norm() / rhs.norm()) < 1e-13
    del x, y, lhs, rhs
    # conductive steady state of the heat system
    nx, L = s.nx, s.L
    gz = torch.as_tensor(s.grid[IZ], device="cuda").view(-1, 1)
    Tlin = (273.0 + (1623.0 - 273.0) * gz / L[IZ]).expand(nx[IZ], nx[IX]).contiguous()
    const = lambda v: torch.full(tuple(nx), float(v), dtype=torch.float64, device="cuda")
    A = pylamp_diff.DiffusionOperator(nx, s.grid, s.gridmp, Tlin, [const(4.0), const(4.0)], const(1250.0),
                                      const(3300.0), const(0.0), o.bcheat, o.bcheatvals, 1e12, ctx=s.ctx)
    Tnew = pylamp_diff.x2t(A.solve(None, rtol=1e-13), nx)
    assert _relmax(Tnew, Tlin) < 1e-10
