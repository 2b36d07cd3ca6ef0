"""Analytic solution of the SolCx-type Stokes benchmark of BASELINE.json configs[1] (the reference's README:22-24
points to Duretz et al. 2011 for its verification; the benchmark itself is not shipped with the reference).

Unit square, free slip on all walls, viscosity eta_left for x < xc and eta_right beyond, body force
f_z = rho*g = -sin(pi z) cos(pi x), f_x = 0 (setups.solcx_fields).  With the stream function
psi = Psi(x) sin(pi z), vx = d psi/dz, vz = -d psi/dx, each half obeys the constant-coefficient ODE
eta (D^2 - pi^2)^2 Psi = pi sin(pi x): a particular solution A sin(pi x) plus (c1 + c2 x) e^{pi x} +
(c3 + c4 x) e^{-pi x}.  The eight constants follow from Psi = Psi'' = 0 on x = 0, 1 (no normal flow, no shear
stress) and from the continuity of vx, vz, the shear stress and the normal stress sigma_xx = -p + 2 eta dvx/dx
across x = xc (an 8 x 8 linear system, solved in floating point).  Pressure: p = P(x) cos(pi z) with
P = [eta (Psi''' - pi^2 Psi') + cos(pi x)] / pi (z-momentum).  Test / verification utility (NumPy, host side).
"""
import numpy as np


def solution(eta_left=1.0, eta_right=1e6, xc=0.5):
    """Returns fields(z, x) -> (vz, vx, p) for arrays of coordinates."""
    k = np.pi

    def basis(x, d):
        ek, em = np.exp(k * x), np.exp(-k * x)
        return np.array([[ek, x * ek, em, x * em],
                         [k * ek, (1 + k * x) * ek, -k * em, (1 - k * x) * em],
                         [k * k * ek, (2 * k + k * k * x) * ek, k * k * em, (-2 * k + k * k * x) * em],
                         [k ** 3 * ek, (3 * k * k + k ** 3 * x) * ek, -k ** 3 * em, (3 * k * k - k ** 3 * x) * em]][d])

    def part(x, d, eta):
        A = np.pi / (eta * (np.pi ** 2 + k ** 2) ** 2)
        return A * [np.sin(np.pi * x), np.pi * np.cos(np.pi * x), -np.pi ** 2 * np.sin(np.pi * x),
                    -np.pi ** 3 * np.cos(np.pi * x)][d]

    eL, eR = float(eta_left), float(eta_right)
    M, r = np.zeros((8, 8)), np.zeros(8)
    M[0, :4], r[0] = basis(0.0, 0), -part(0.0, 0, eL)
    M[1, :4], r[1] = basis(0.0, 2), -part(0.0, 2, eL)
    M[2, 4:], r[2] = basis(1.0, 0), -part(1.0, 0, eR)
    M[3, 4:], r[3] = basis(1.0, 2), -part(1.0, 2, eR)
    for row, d in ((4, 0), (5, 1)):
        M[row, :4], M[row, 4:] = basis(xc, d), -basis(xc, d)
        r[row] = part(xc, d, eR) - part(xc, d, eL)
    shear = lambda d0, d2, eta: eta * (-k * k * d0 - d2)
    M[6, :4], M[6, 4:] = shear(basis(xc, 0), basis(xc, 2), eL), -shear(basis(xc, 0), basis(xc, 2), eR)
    r[6] = shear(part(xc, 0, eR), part(xc, 2, eR), eR) - shear(part(xc, 0, eL), part(xc, 2, eL), eL)
    normal = lambda d3, d1, eta: -eta * (d3 - k * k * d1) / k + 2 * eta * k * d1     # (the cos(pi x)/pi term is continuous)
    M[7, :4], M[7, 4:] = normal(basis(xc, 3), basis(xc, 1), eL), -normal(basis(xc, 3), basis(xc, 1), eR)
    r[7] = normal(part(xc, 3, eR), part(xc, 1, eR), eR) - normal(part(xc, 3, eL), part(xc, 1, eL), eL)
    sc = np.abs(M).max(axis=1)
    c = np.linalg.solve(M / sc[:, None], r / sc)

    def fields(z, x):
        z, x = np.asarray(z, dtype=np.float64), np.asarray(x, dtype=np.float64)
        left = x < xc
        eta = np.where(left, eL, eR)

        def ev(d):
            out = np.empty_like(x)
            for sl, cc, e in ((left, c[:4], eL), (~left, c[4:], eR)):
                out[sl] = np.tensordot(cc, basis(x[sl], d), axes=1) + part(x[sl], d, e)
            return out

        Psi, d1, d3 = ev(0), ev(1), ev(3)
        vx = k * Psi * np.cos(k * z)
        vz = -d1 * np.sin(k * z)
        p = (eta * (d3 - k * k * d1) + np.cos(np.pi * x)) / k * np.cos(k * z)
        return vz, vx, p

    return fields


def errors(nx, grid, gridmp, vz, vx, ptilde, Kcont, eta_right=1e6):
    """Relative L2 errors (vz, vx, p) of a discrete solution in the reference's staggered layout (vz at
    (z_i, x_{j+1/2}), vx at (z_{i+1/2}, x_j), P~ = P/Kcont at the cell centres, ghost row/column last) against
    the analytic solution; pressures are compared after removing their means over the real cells."""
    f = solution(eta_right=eta_right)
    zz, xx = np.meshgrid(grid[0], gridmp[1], indexing="ij")
    vza = f(zz, xx)[0][:, :-1]
    zz, xx = np.meshgrid(gridmp[0], grid[1], indexing="ij")
    vxa = f(zz, xx)[1][:-1, :]
    zz, xx = np.meshgrid(gridmp[0], gridmp[1], indexing="ij")
    pa = f(zz, xx)[2][:-1, :-1]
    pn = np.asarray(ptilde)[:-1, :-1] * Kcont
    pn, pa = pn - pn.mean(), pa - pa.mean()
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    return rel(np.asarray(vz)[:, :-1], vza), rel(np.asarray(vx)[:-1, :], vxa), rel(pn, pa)
