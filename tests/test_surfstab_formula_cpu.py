"""The free-surface stabilisation terms as the CUDA kernels compute them (stokes.cu: k_surfstab_planes
+ the SURF branches): z-momentum row (i,j) gains theta*dt*g_z*q(i,j), x-momentum row (i,j) gains
theta*dt*g_x*q(i,j) with q = Dz*vz(i,j) + Dx*vx(i,j) and the centred density gradients
  Dz(i,j) = (rho[i+1,j]+rho[i+1,j+1]-rho[i-1,j]-rho[i-1,j+1]) / 2 / (z[i+1]-z[i-1]),
  Dx(i,j) = (rho[i,j+1]+rho[i+1,j+1]-rho[i,j-1]-rho[i+1,j-1]) / 2 / (x[j+1]-x[j-1]),
restated in NumPy and held to the difference between the stabilised and the plain matrix of the
oracle (which is pinned bit-exactly to the reference's assembly, pylamp_stokes.py:422-426, :483-487).
Pins the formula, the row ranges and the columns on CPU; the kernels themselves are tested on the GPU."""
import numpy as np
import pytest

from oracle import pylamp_oracle as O


@pytest.mark.parametrize("gvec", [(9.81, 0.0), (9.81, 2.5)])
def test_kernel_formula_equals_matrix_difference(golden_kernels, gvec, monkeypatch):
    g = golden_kernels
    nx = list(g["nx"])
    nz, nxx = nx
    gz, gx = np.asarray(g["st_gz"]), np.asarray(g["st_gx"])
    rho = np.asarray(g["st_rho"])
    monkeypatch.setattr(O, "G", [gvec[0], gvec[1]])
    theta, dt = 0.5, 1e3
    A1, _ = O.makeStokesMatrix(nx, [gz, gx], g["st_etas"], g["st_etan"], rho, [1, 1, 1, 1], surfstab=True, tstep=dt,
                               surfstab_theta=theta)
    A0, _ = O.makeStokesMatrix(nx, [gz, gx], g["st_etas"], g["st_etan"], rho, [1, 1, 1, 1])
    # the planes of k_surfstab_planes
    Dz, Dx = np.zeros((nz, nxx)), np.zeros((nz, nxx))
    i, j = np.meshgrid(np.arange(1, nz - 1), np.arange(1, nxx - 1), indexing="ij")
    Dz[i, j] = 0.5 * (rho[i + 1, j] + rho[i + 1, j + 1] - rho[i - 1, j] - rho[i - 1, j + 1]) / (gz[i + 1] - gz[i - 1])
    Dx[i, j] = 0.5 * (rho[i, j + 1] + rho[i + 1, j + 1] - rho[i, j - 1] - rho[i + 1, j - 1]) / (gx[j + 1] - gx[j - 1])
    rng = np.random.default_rng(1)
    x = rng.normal(size=A0.shape[0])
    vz, vx = x[0::3].reshape(nz, nxx), x[1::3].reshape(nz, nxx)
    q = Dz * vz + Dx * vx
    # rows: the momentum rows of the reduced system (stencil.cuh is_vz_row / is_vx_row for the reference closure)
    yz, yx = np.zeros((nz, nxx)), np.zeros((nz, nxx))
    yz[1:nz - 1, 1:nxx - 2] = theta * dt * gvec[0] * q[1:nz - 1, 1:nxx - 2]
    yx[1:nz - 2, 1:nxx - 1] = theta * dt * gvec[1] * q[1:nz - 2, 1:nxx - 1]
    want = (A1 - A0) @ x
    got = np.zeros_like(want)
    got[0::3], got[1::3] = yz.ravel(), yx.ravel()
    assert np.linalg.norm(want) > 0
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)
