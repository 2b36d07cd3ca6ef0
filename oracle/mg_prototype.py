"""Design study (TEST INFRASTRUCTURE / NOT SHIPPED): NumPy/SciPy prototype of the GPU Stokes
solver algorithm -- FGMRES with a block-triangular preconditioner whose velocity block is one
geometric-multigrid V-cycle -- run on the matrices the oracle assembles, to pick smoothers,
coarsening rules and tolerances on the CPU before the CUDA kernels are written.

Nothing in pylamp_b200/ imports this file.
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import pylamp_oracle as O


def dof_sets(nz, nxx):
    """Index sets (full interleaved numbering) of the free unknowns = interior rows."""
    def g(i, j, eq):
        return (i * nxx + j) * 3 + eq
    i, j = np.meshgrid(np.arange(1, nz - 1), np.arange(1, nxx - 2), indexing="ij")
    vz = g(i, j, 0).ravel()
    i, j = np.meshgrid(np.arange(1, nz - 2), np.arange(1, nxx - 1), indexing="ij")
    vx = g(i, j, 1).ravel()
    i, j = np.meshgrid(np.arange(0, nz - 1), np.arange(0, nxx - 1), indexing="ij")
    corner = ((i == 0) | (i == nz - 2)) & ((j == 0) | (j == nxx - 2))
    anchor = (i == 3) & (j == 2)
    keep = ~(corner | anchor)
    p = g(i[keep], j[keep], 2).ravel()
    return vz, vx, p


class Level:
    def __init__(self, nz, nxx, gz, gx, etas, etan, rho, bc, need_full=False):
        self.nz, self.nxx, self.gz, self.gx = nz, nxx, gz, gx
        self.etas, self.etan = etas, etan
        A, rhs = O.makeStokesMatrix([nz, nxx], [gz, gx], etas, etan, rho, bc)
        A = A.tocsr()
        n = A.shape[0]
        vz, vx, p = dof_sets(nz, nxx)
        I = np.concatenate([vz, vx, p])
        isI = np.zeros(n, bool)
        isI[I] = True
        B = np.where(~isI)[0]
        A_BB = A[B][:, B].tocsc()
        A_BI = A[B][:, I]
        X = spla.spsolve(A_BB, A_BI.tocsc())           # slaves = -X @ masters
        E = sp.vstack([sp.identity(len(I), format="csr"), -sp.csr_matrix(X)]).tocsr()
        perm = np.concatenate([I, B])
        # E maps interior vector -> full vector (in perm order); build full-order version
        Pm = sp.csr_matrix((np.ones(n), (perm, np.arange(n))), shape=(n, n))
        self.E = (Pm @ E).tocsr()                      # (n_full, n_I)
        self.S = sp.csr_matrix((np.ones(len(I)), (np.arange(len(I)), I)), shape=(len(I), n))
        Ar = (self.S @ A @ self.E).tocsr()
        self.nv = len(vz) + len(vx)
        self.nvz = len(vz)
        self.np_ = len(p)
        self.Ar = Ar
        self.K = Ar[:self.nv][:, :self.nv].tocsr()
        self.G = Ar[:self.nv][:, self.nv:].tocsr()
        self.D = Ar[self.nv:][:, :self.nv].tocsr()
        self.C = Ar[self.nv:][:, self.nv:].tocsr()
        self.b = self.S @ rhs
        self.A_full, self.rhs_full = A, rhs
        self.I = I
        self.Kc = O.stokes_scaling([gz, gx], etas, etan)[0]
        self.Kdiag = self.K.diagonal()
        # cell viscosity for the Schur scaling, on the free pressure cells
        pi = (p // 3) // nxx
        pj = (p // 3) % nxx
        self.eta_p = etan[pi, pj]


class ProperLevel:
    """Coarse-level velocity operator with the standard staggered free-slip closure
    (zero shear stress on the wall faces); every in-domain staggered velocity is an unknown."""
    def __init__(self, nz, nxx, gz, gx, etas, etan):
        self.nz, self.nxx, self.gz, self.gx = nz, nxx, gz, gx
        self.etas, self.etan = etas, etan

        def g(i, j, eq):
            return (i * nxx + j) * 3 + eq
        T = O._Triplets()
        # vz rows
        i, j = [a.ravel() for a in np.meshgrid(np.arange(1, nz - 1), np.arange(0, nxx - 1), indexing="ij")]
        rows = g(i, j, 0)
        dzc = (gz[i + 1] - gz[i - 1]) / 2
        dzp, dzm = gz[i + 1] - gz[i], gz[i] - gz[i - 1]
        dxj = gx[j + 1] - gx[j]
        cN, cS = 2 * etan[i, j] / dzp / dzc, 2 * etan[i - 1, j] / dzm / dzc
        T.add(rows, g(i + 1, j, 0), cN), T.add(rows, g(i, j, 0), -cN)
        T.add(rows, g(i - 1, j, 0), cS), T.add(rows, g(i, j, 0), -cS)
        for side, jj in ((+1, j + 1), (-1, j)):          # shear stress at node (i, jj)
            m = (jj > 0) & (jj < nxx - 1)
            r, ii, jn = rows[m], i[m], jj[m]
            es = etas[ii, jn]
            dxn = (gx[jn + 1] - gx[jn - 1]) / 2          # distance between the two vz points
            c1 = side * es / dxn / dxj[m]
            T.add(r, g(ii, jn, 0), c1), T.add(r, g(ii, jn - 1, 0), -c1)
            c2 = side * es / dzc[m] / dxj[m]
            T.add(r, g(ii, jn, 1), c2), T.add(r, g(ii - 1, jn, 1), -c2)
        # vx rows
        i, j = [a.ravel() for a in np.meshgrid(np.arange(0, nz - 1), np.arange(1, nxx - 1), indexing="ij")]
        rows = g(i, j, 1)
        dxc = (gx[j + 1] - gx[j - 1]) / 2
        dxp, dxm = gx[j + 1] - gx[j], gx[j] - gx[j - 1]
        dzi = gz[i + 1] - gz[i]
        cE, cW = 2 * etan[i, j] / dxp / dxc, 2 * etan[i, j - 1] / dxm / dxc
        T.add(rows, g(i, j + 1, 1), cE), T.add(rows, g(i, j, 1), -cE)
        T.add(rows, g(i, j - 1, 1), cW), T.add(rows, g(i, j, 1), -cW)
        for side, ii in ((+1, i + 1), (-1, i)):          # shear stress at node (ii, j)
            m = (ii > 0) & (ii < nz - 1)
            r, inn, jj = rows[m], ii[m], j[m]
            es = etas[inn, jj]
            dzn = (gz[inn + 1] - gz[inn - 1]) / 2
            c1 = side * es / dzn / dzi[m]
            T.add(r, g(inn, jj, 1), c1), T.add(r, g(inn - 1, jj, 1), -c1)
            c2 = side * es / dxc[m] / dzi[m]
            T.add(r, g(inn, jj, 0), c2), T.add(r, g(inn, jj - 1, 0), -c2)
        n = nz * nxx * 3
        A = T.tocsr(n)
        i, j = np.meshgrid(np.arange(1, nz - 1), np.arange(0, nxx - 1), indexing="ij")
        vz = g(i, j, 0).ravel()
        i, j = np.meshgrid(np.arange(0, nz - 1), np.arange(1, nxx - 1), indexing="ij")
        vx = g(i, j, 1).ravel()
        I = np.concatenate([vz, vx])
        self.nv, self.nvz = len(I), len(vz)
        self.S = sp.csr_matrix((np.ones(len(I)), (np.arange(len(I)), I)), shape=(len(I), n))
        self.E = self.S.T.tocsr()
        self.K = (self.S @ A @ self.E).tocsr()
        self.Kdiag = self.K.diagonal()
        self.I = I


def interp1d_nodes(nf):
    """fine nodes (nf) from coarse nodes ((nf-1)/2+1): linear."""
    nc = (nf - 1) // 2 + 1
    P = sp.lil_matrix((nf, nc))
    for i in range(nf):
        if i % 2 == 0:
            P[i, i // 2] = 1.0
        else:
            P[i, i // 2] = 0.5
            P[i, i // 2 + 1] = 0.5
    return P.tocsr()


def interp1d_mids(nf):
    """fine midpoints (index 0..nf-2 real, nf-1 ghost) from coarse midpoints (0..nc-2 real):
    linear with constant extrapolation at the ends."""
    nc = (nf - 1) // 2 + 1
    P = sp.lil_matrix((nf, nc))
    for j in range(nf - 1):
        J = j // 2
        if j % 2 == 0:
            a, b = J - 1, J
            wa, wb = 0.25, 0.75
        else:
            a, b = J, J + 1
            wa, wb = 0.75, 0.25
        a = min(max(a, 0), nc - 2)
        b = min(max(b, 0), nc - 2)
        P[j, a] += wa
        P[j, b] += wb
    return P.tocsr()


def transfer(fine, coarse):
    """Reduced prolongation (fine interior velocities <- coarse interior velocities)."""
    nz, nxx, nzc, nxc = fine.nz, fine.nxx, coarse.nz, coarse.nxx
    Pzn, Pxn = interp1d_nodes(nz), interp1d_nodes(nxx)
    Pzm, Pxm = interp1d_mids(nz), interp1d_mids(nxx)
    Pvz = sp.kron(Pzn, Pxm).tocsr()        # vz: node in z, mid in x
    Pvx = sp.kron(Pzm, Pxn).tocsr()        # vx: mid in z, node in x
    nf, nc = nz * nxx, nzc * nxc
    # full interleaved prolongation on velocity dofs only
    rows, cols, vals = [], [], []
    for eq, Pm in ((0, Pvz), (1, Pvx)):
        Pm = Pm.tocoo()
        rows.append(Pm.row * 3 + eq), cols.append(Pm.col * 3 + eq), vals.append(Pm.data)
    Pfull = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(3 * nf, 3 * nc))
    Pr = (fine.S @ Pfull @ coarse.E).tocsr()[:fine.nv][:, :coarse.nv]
    return Pr.tocsr()


def coarsen_eta(etas, etan, mode="geom"):
    nz, nxx = etas.shape
    inject = mode.startswith("inj-")
    mode = mode.replace("inj-", "")
    nzc, nxc = (nz - 1) // 2 + 1, (nxx - 1) // 2 + 1
    f = {"geom": (np.log, np.exp), "arith": (lambda a: a, lambda a: a),
         "harm": (lambda a: 1 / a, lambda a: 1 / a)}[mode]
    # nodes: weighted 9-point average (full weighting) with edge clamping
    L = f[0](etas)
    Lp = np.pad(L, 1, mode="edge")
    w = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]]) / 16.0
    acc = np.zeros((nzc, nxc))
    for a in range(3):
        for b in range(3):
            acc += w[a, b] * Lp[a:a + nz:2, b:b + nxx:2][:nzc, :nxc]
    es = f[1](acc)
    if inject:
        es = etas[::2, ::2].copy()
    # centres: the 4 fine cells inside each coarse cell
    Lc = f[0](etan[:nz - 1, :nxx - 1])
    acc = 0.25 * (Lc[0::2, 0::2] + Lc[1::2, 0::2] + Lc[0::2, 1::2] + Lc[1::2, 1::2])
    en = np.ones((nzc, nxc)) * f[1](acc).mean()
    en[:nzc - 1, :nxc - 1] = f[1](acc)
    return es, en


class MG:
    def __init__(self, nz, nxx, gz, gx, etas, etan, rho, bc, min_cells=4, eta_mode="geom",
                 smoother="cheb", nu=3, coarse="lu", proper=True):
        self.levels = []
        self.P = []
        self.smoother, self.nu = smoother, nu
        self.proper = proper
        while True:
            if proper and self.levels:
                self.levels.append(ProperLevel(nz, nxx, gz, gx, etas, etan))
            else:
                self.levels.append(Level(nz, nxx, gz, gx, etas, etan, rho, bc))
            if (nz - 1) % 2 or (nxx - 1) % 2 or min(nz - 1, nxx - 1) // 2 < min_cells:
                break
            etas, etan = coarsen_eta(etas, etan, eta_mode)
            rho = rho[::2, ::2]
            gz, gx = gz[::2], gx[::2]
            nz, nxx = (nz - 1) // 2 + 1, (nxx - 1) // 2 + 1
        for a, b in zip(self.levels[:-1], self.levels[1:]):
            self.P.append(transfer(a, b))
        self.lu = spla.splu(self.levels[-1].K.tocsc())
        self.lmax = []
        for lv in self.levels:
            # estimate lambda_max(D^-1 K) by a few power iterations
            x = np.random.default_rng(0).normal(size=lv.nv)
            for _ in range(20):
                x = (lv.K @ x) / lv.Kdiag
                lam = np.linalg.norm(x)
                x /= lam
            self.lmax.append(lam * 1.1)
        if smoother == "gs":
            self.Ltri = [sp.tril(lv.K).tocsr() for lv in self.levels]
            self.Utri = [sp.triu(lv.K).tocsr() for lv in self.levels]

    def smooth(self, l, x, b, nu, post=False):
        lv = self.levels[l]
        if self.smoother == "jac":
            for _ in range(nu):
                x = x + 0.6 * (b - lv.K @ x) / lv.Kdiag
            return x
        if self.smoother == "gs":
            T = self.Utri[l] if post else self.Ltri[l]
            for _ in range(nu):
                x = x + spla.spsolve_triangular(T, b - lv.K @ x, lower=not post)
            return x
        # Chebyshev on D^-1 K over [lmax/alpha, lmax]
        lmax = self.lmax[l]
        lmin = lmax / 8.0
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho_ = 1.0 / sigma
        r = (b - lv.K @ x) / lv.Kdiag
        d = r / theta
        for k in range(nu):
            x = x + d
            if k == nu - 1:
                break
            r = (b - lv.K @ x) / lv.Kdiag
            rho_new = 1.0 / (2 * sigma - rho_)
            d = rho_new * rho_ * d + 2 * rho_new / delta * r
            rho_ = rho_new
        return x

    def vcycle(self, l, b):
        lv = self.levels[l]
        if l == len(self.levels) - 1:
            return self.lu.solve(b)
        x = self.smooth(l, np.zeros_like(b), b, self.nu)
        r = b - lv.K @ x
        if getattr(self, "scaled_restriction", False):
            Pn = self.P[l]
            wsum = Pn.T @ np.ones(Pn.shape[0])
            rc = 4.0 * self.levels[l + 1].Kdiag * ((Pn.T @ (r / lv.Kdiag)) / wsum)
        else:
            rc = 0.25 * (self.P[l].T @ r)
        ec = self.vcycle(l + 1, rc)
        x = x + self.P[l] @ ec
        x = self.smooth(l, x, b, self.nu, post=True)
        return x


def fgmres(Aop, b, Mop, x0=None, m=40, tol=1e-12, maxit=400, verbose=False):
    n = b.shape[0]
    x = np.zeros(n) if x0 is None else x0.copy()
    bnorm = np.linalg.norm(b)
    its = 0
    hist = []
    while its < maxit:
        r = b - Aop(x)
        beta = np.linalg.norm(r)
        hist.append(beta / bnorm)
        if beta / bnorm < tol:
            break
        V = [r / beta]
        Z = []
        H = np.zeros((m + 1, m))
        gvec = np.zeros(m + 1)
        gvec[0] = beta
        k_used = 0
        for k in range(m):
            z = Mop(V[k])
            w = Aop(z)
            for i in range(k + 1):
                H[i, k] = V[i] @ w
                w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w)
            V.append(w / H[k + 1, k])
            Z.append(z)
            its += 1
            k_used = k + 1
            y, res, _, _ = np.linalg.lstsq(H[:k + 2, :k + 1], gvec[:k + 2], rcond=None)
            rn = np.linalg.norm(H[:k + 2, :k + 1] @ y - gvec[:k + 2])
            hist.append(rn / bnorm)
            if verbose:
                print(its, rn / bnorm)
            if rn / bnorm < tol or its >= maxit:
                break
        x = x + sum(yi * zi for yi, zi in zip(y, Z[:k_used]))
    return x, its, hist


def solve(mg, tol=1e-12, m=40, maxit=300, schur_sign=1.0, verbose=False, inner=1):
    lv = mg.levels[0]
    nv = lv.nv
    sdiag = lv.Kc ** 2 / lv.eta_p * schur_sign

    def Aop(x):
        return lv.Ar @ x

    def Mop(r):
        dp = r[nv:] / sdiag
        rv = r[:nv] - lv.G @ dp
        dv = mg.vcycle(0, rv)
        for _ in range(inner - 1):
            dv = dv + mg.vcycle(0, rv - lv.K @ dv)
        return np.concatenate([dv, dp])

    x, its, hist = fgmres(Aop, lv.b, Mop, m=m, tol=tol, maxit=maxit, verbose=verbose)
    return lv.E @ x, its, hist


def compare(xfull, xref, nz, nxx):
    out = []
    for eq in range(3):
        a, b = xfull[eq::3], xref[eq::3]
        out.append(np.linalg.norm(a - b) / np.linalg.norm(b))
    return out


if __name__ == "__main__":
    from pylamp_b200 import setups
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65
    case = sys.argv[2] if len(sys.argv) > 2 else "solcx"
    smoother = sys.argv[3] if len(sys.argv) > 3 else "cheb"
    nu = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    mode = sys.argv[5] if len(sys.argv) > 5 else "geom"
    if case == "solcx":
        nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
    elif case == "rand":
        nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
        rng = np.random.default_rng(1)
        from scipy.ndimage import gaussian_filter
        f = gaussian_filter(rng.normal(size=(n, n)), 3)
        f = f / np.abs(f).max()
        etas = 10 ** (3 * f)
        etan = etas.copy()
        rho = 3000 + 100 * gaussian_filter(rng.normal(size=(n, n)), 2)
    elif case == "incl":
        nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
        zs, xs = np.meshgrid(grid[0], grid[1], indexing="ij")
        zc, xc = np.meshgrid(gridmp[0], gridmp[1], indexing="ij")
        etas = np.where((zs - 0.3) ** 2 + (xs - 0.5) ** 2 < 0.1 ** 2, 1e6, 1.0)
        etan = np.where((zc - 0.3) ** 2 + (xc - 0.5) ** 2 < 0.1 ** 2, 1e6, 1.0)
        rho = np.where((zs - 0.3) ** 2 + (xs - 0.5) ** 2 < 0.1 ** 2, 1.1, 1.0)
    t = time.time()
    mg = MG(nx[0], nx[1], grid[0], grid[1], etas, etan, rho, [1, 1, 1, 1], smoother=smoother,
            nu=nu, eta_mode=mode)
    print("levels", [(l.nz, l.nxx) for l in mg.levels], "setup %.1fs" % (time.time() - t),
          "lmax", np.round(mg.lmax, 3))
    lv = mg.levels[0]
    print("C block nnz", lv.C.nnz, " K sym err",
          abs(lv.K - lv.K.T).max() / abs(lv.K).max())
    t = time.time()
    xref = O.solve_refined(lv.A_full, lv.rhs_full)
    print("direct %.1fs" % (time.time() - t))
    for tol in (1e-10, 1e-12, 1e-13):
        t = time.time()
        x, its, hist = solve(mg, tol=tol)
        print("tol", tol, "its", its, "time %.1fs" % (time.time() - t), "err vz,vx,p",
              ["%.2e" % e for e in compare(x, xref, nx[0], nx[1])],
              "true relres %.2e" % (np.linalg.norm(lv.rhs_full - lv.A_full @ x) / np.linalg.norm(lv.rhs_full)))


def bicgstab(Aop, b, Mop, rtol, maxit):
    """Right-preconditioned BiCGStab from x0 = 0; returns (x, iterations, recurrence relres)."""
    x = np.zeros_like(b)
    r = b.copy()
    rhat = r.copy()
    bn = np.linalg.norm(b)
    rho = alpha = omega = 1.0
    v = np.zeros_like(b)
    p = np.zeros_like(b)
    for it in range(1, maxit + 1):
        rho_new = rhat @ r
        if rho_new == 0:
            break
        beta = (rho_new / rho) * (alpha / omega)
        p = r + beta * (p - omega * v)
        ph = Mop(p)
        v = Aop(ph)
        alpha = rho_new / (rhat @ v)
        s = r - alpha * v
        sh = Mop(s)
        t = Aop(sh)
        omega = (t @ s) / (t @ t)
        x = x + alpha * ph + omega * sh
        r = s - omega * t
        rho = rho_new
        if np.linalg.norm(r) <= rtol * bn:
            return x, it, np.linalg.norm(r) / bn
    return x, maxit, np.linalg.norm(r) / bn


def solve_refine(mg, err_tol=1e-10, inner_rtol=1e-5, max_outer=8, maxit=200, verbose=True):
    """Outer iterative refinement around BiCGStab; the stopping test is the preconditioned
    true residual M^-1 (b - A x) per component relative to the component's norm."""
    lv = mg.levels[0]
    nv, nvz = lv.nv, lv.nvz
    sdiag = lv.Kc ** 2 / lv.eta_p
    calls = [0]

    def Aop(x):
        return lv.Ar @ x

    def Mop(r):
        calls[0] += 1
        dp = r[nv:] / sdiag
        dv = mg.vcycle(0, r[:nv] - lv.G @ dp)
        return np.concatenate([dv, dp])

    def comps(z):
        return [np.linalg.norm(z[:nvz]), np.linalg.norm(z[nvz:nv]), np.linalg.norm(z[nv:])]

    x = np.zeros_like(lv.b)
    for outer in range(max_outer):
        r = lv.b - Aop(x)
        z = Mop(r)
        est = max(a / max(b_, 1e-300) for a, b_ in zip(comps(z), comps(x))) if outer else 1.0
        if verbose:
            print("  outer", outer, "est err %.1e" % est, "calls", calls[0])
        if est < err_tol:
            break
        d, its, rr = bicgstab(Aop, r, Mop, inner_rtol, maxit)
        x = x + d
    return lv.E @ x, calls[0]


# ---------------------------------------------------------------------------------------
# second design iteration: arithmetic coarsening variants, scaled Krylov norm, inner GCR
# (the CUDA solver in pylamp_b200/csrc/stokes.cu follows MG2 with ms='fw', mn='4x4'|'2x2',
#  smoother='cheb', nu=3 and solve_scaled(ncyc=1))
# ---------------------------------------------------------------------------------------
from pylamp_b200 import setups  # noqa: E402

def coarsen(etas, etan, ms, mn):
    nz, nxx = etas.shape
    nzc, nxc = (nz-1)//2+1, (nxx-1)//2+1
    # nodes
    Lp = np.pad(etas, 1, mode='edge')
    def win(a,b): return Lp[a:a+nz:2, b:b+nxx:2][:nzc,:nxc]
    if ms=='inj': es = etas[::2,::2].copy()
    elif ms=='fw':
        w = np.array([[1,2,1],[2,4,2],[1,2,1]])/16.
        es = sum(w[a,b]*win(a,b) for a in range(3) for b in range(3))
    elif ms=='max':
        es = np.max([win(a,b) for a in range(3) for b in range(3)],axis=0)
    elif ms=='geom':
        w = np.array([[1,2,1],[2,4,2],[1,2,1]])/16.
        es = np.exp(sum(w[a,b]*np.log(win(a,b)) for a in range(3) for b in range(3)))
    # centres: real cells [0:nz-1, 0:nxx-1]
    c = etan[:nz-1,:nxx-1]
    ncz, ncx = nzc-1, nxc-1
    if mn=='2x2':
        ec = 0.25*(c[0::2,0::2]+c[1::2,0::2]+c[0::2,1::2]+c[1::2,1::2])
    elif mn=='geom':
        ec = np.exp(0.25*(np.log(c[0::2,0::2])+np.log(c[1::2,0::2])+np.log(c[0::2,1::2])+np.log(c[1::2,1::2])))
    else:
        cp = np.pad(c, 1, mode='edge')   # cp[k] = c[k-1]
        def w4(a,b): return cp[a:a+2*ncz:2, b:b+2*ncx:2]  # rows 2I-1+a
        if mn=='4x4':
            w1 = np.array([1,3,3,1])/8.
            ec = sum(w1[a]*w1[b]*w4(a,b) for a in range(4) for b in range(4))
        elif mn=='max':
            ec = np.max([w4(a,b) for a in range(4) for b in range(4)],axis=0)
    en = np.ones((nzc,nxc))*ec.mean(); en[:ncz,:ncx]=ec
    return es, en

class MG2(MG):
    def __init__(self, nz, nxx, gz, gx, etas, etan, rho, bc, ms='fw', mn='4x4', smoother='cheb', nu=3, min_cells=4, minres=False, cheb_ratio=8.0):
        self.levels=[]; self.P=[]; self.smoother=smoother; self.nu=nu; self.minres=minres; self.cheb_ratio=cheb_ratio
        first=True
        while True:
            self.levels.append(Level(nz,nxx,gz,gx,etas,etan,rho,bc) if first else ProperLevel(nz,nxx,gz,gx,etas,etan))
            first=False
            if (nz-1)%2 or (nxx-1)%2 or min(nz-1,nxx-1)//2 < min_cells: break
            etas, etan = coarsen(etas, etan, ms, mn)
            rho = rho[::2,::2]; gz, gx = gz[::2], gx[::2]
            nz, nxx = (nz-1)//2+1, (nxx-1)//2+1
        for a,b in zip(self.levels[:-1], self.levels[1:]): self.P.append(transfer(a,b))
        self.lu = spla.splu(self.levels[-1].K.tocsc())
        self.lmax=[]
        for lv in self.levels:
            x=np.random.default_rng(0).normal(size=lv.nv)
            for _ in range(20):
                x=(lv.K@x)/lv.Kdiag; lam=np.linalg.norm(x); x/=lam
            self.lmax.append(lam*1.1)
        if smoother=='cgs':
            self.colors=[]
            for lv in self.levels:
                I=lv.I; node=I//3; eq=I%3; i=node//lv.nxx; j=node%lv.nxx
                col = eq*2 + (i+j)%2
                self.colors.append([ (np.where(col==c)[0], lv.K[np.where(col==c)[0]]) for c in range(4)])
    def smooth(self,l,x,b,nu,post=False):
        lv=self.levels[l]
        if self.smoother=='cgs':
            order = range(4) if not post else range(3,-1,-1)
            for _ in range(nu):
                for c in order:
                    idx,Kc = self.colors[l][c]
                    x[idx] += (b[idx]-Kc@x)/lv.Kdiag[idx]
            return x
        if self.smoother=='cheb':
            lmax=self.lmax[l]; lmin=lmax/self.cheb_ratio
            theta,delta=0.5*(lmax+lmin),0.5*(lmax-lmin); sigma=theta/delta; rho_=1/sigma
            r=(b-lv.K@x)/lv.Kdiag; d=r/theta
            for k in range(nu):
                x=x+d
                if k==nu-1: break
                r=(b-lv.K@x)/lv.Kdiag
                rho_new=1/(2*sigma-rho_); d=rho_new*rho_*d+2*rho_new/delta*r; rho_=rho_new
            return x
        return MG.smooth(self,l,x,b,nu,post)
    def vcycle(self,l,b):
        lv=self.levels[l]
        if l==len(self.levels)-1: return self.lu.solve(b)
        x=self.smooth(l,np.zeros_like(b),b,self.nu)
        r=b-lv.K@x
        ec=self.vcycle(l+1, 0.25*(self.P[l].T@r))
        e=self.P[l]@ec
        if self.minres:
            Ke=lv.K@e; a=(r@Ke)/(Ke@Ke); e=a*e
        x=x+e
        return self.smooth(l,x,b,self.nu,post=True)

def fields(case,n):
    nx, L, grid, gridmp, etas, etan, rho = setups.solcx_fields(n)
    zs, xs = np.meshgrid(grid[0], grid[1], indexing="ij")
    zc, xc = np.meshgrid(gridmp[0], gridmp[1], indexing="ij")
    if case=='const': etas[:]=1; etan[:]=1
    if case=='incl':
        f=lambda z,x: np.where((z-0.3)**2+(x-0.5)**2<0.1**2,1e6,1.0)
        etas=f(zs,xs); etan=f(zc,xc); rho=np.where((zs-0.3)**2+(xs-0.5)**2<0.1**2,1.1,1.0)
    if case=='sinkers':
        rng=np.random.default_rng(3); cz=rng.uniform(0.1,0.9,8); cx=rng.uniform(0.1,0.9,8)
        def f(z,x):
            m=np.zeros_like(z,bool)
            for a,b in zip(cz,cx): m|=((z-a)**2+(x-b)**2<0.05**2)
            return m
        etas=np.where(f(zs,xs),1e4,1.0); etan=np.where(f(zc,xc),1e4,1.0); rho=np.where(f(zs,xs),1.2,1.0)
    if case=='arrh':
        T=lambda z,x: 273+1350*z+0.05*1350*np.sin(np.pi*z)*np.cos(np.pi*x)
        e=lambda T: np.clip(1e20*np.exp(120e3/(8.314*T)-120e3/(8.314*1623)),1e17,1e23)
        etas=e(T(zs,xs)); etan=e(T(zc,xc)); rho=3300/(3.5e-5*(T(zs,xs)-1623)+1)
    if case=='smallincl':
        f=lambda z,x: np.where((z-0.3)**2+(x-0.5)**2<(2.2/(n-1))**2,1e10,1.0)
        etas=f(zs,xs); etan=f(zc,xc); rho=np.where((zs-0.3)**2+(xs-0.5)**2<(2.2/(n-1))**2,1.1,1.0)
    return nx,L,grid,gridmp,etas,etan,rho

def solve_scaled(mg, tol=1e-12, m=50, maxit=150, ncyc=1, scaled=True, inner_gmres=0):
    lv=mg.levels[0]; nv=lv.nv
    sd = lv.Kc**2/lv.eta_p
    if scaled:
        W=np.concatenate([1/np.sqrt(np.abs(lv.Kdiag)), 1/np.sqrt(sd)])
    else:
        W=np.ones(lv.Ar.shape[0])
    def Aop(y): return W*(lv.Ar@(W*y))
    def Ksolve(rv):
        dv=mg.vcycle(0,rv)
        for _ in range(ncyc-1): dv=dv+mg.vcycle(0,rv-lv.K@dv)
        return dv
    def M(r):
        dp=r[nv:]/sd; dv=Ksolve(r[:nv]-lv.G@dp); return np.concatenate([dv,dp])
    def Mop(rh): return M(rh/W)/W
    xh,its,hist=fgmres(Aop, W*lv.b, Mop, m=m, tol=tol, maxit=maxit)
    return lv.E@(W*xh), its, hist

def gcr(Aop, b, Mop, k, rtol=0.0):
    """k steps of GCR (flexible), x0=0"""
    x=np.zeros_like(b); r=b.copy(); Zs=[]; Cs=[]; r0=np.linalg.norm(b)
    for i in range(k):
        z=Mop(r); c=Aop(z)
        for zj,cj in zip(Zs,Cs):
            a=cj@c; c=c-a*cj; z=z-a*zj
        nc=np.linalg.norm(c); c/=nc; z/=nc
        a=c@r; x+=a*z; r-=a*c; Zs.append(z); Cs.append(c)
        if np.linalg.norm(r)<rtol*r0: break
    return x, i+1

def solve_inner(mg, tol=1e-12, m=50, maxit=150, kin=4, rtol_in=0.0, mode='upper'):
    lv=mg.levels[0]; nv=lv.nv
    sd = lv.Kc**2/lv.eta_p
    W=np.concatenate([1/np.sqrt(np.abs(lv.Kdiag)), 1/np.sqrt(sd)])
    Wv=W[:nv]
    nV=[0]
    def Aop(y): return W*(lv.Ar@(W*y))
    def Vc(r): nV[0]+=1; return mg.vcycle(0,r)
    def Ksolve(rv):
        if kin<=1: return Vc(rv)
        # scaled inner GCR: minimise ||Wv (rv - K dv)||
        dv,_=gcr(lambda y: Wv*(lv.K@(Wv*y)), Wv*rv, lambda r: Vc(r/Wv)/Wv, kin, rtol_in)
        return Wv*dv
    def M(r):
        dp=r[nv:]/sd; dv=Ksolve(r[:nv]-lv.G@dp); return np.concatenate([dv,dp])
    def Mop(rh): return M(rh/W)/W
    xh,its,hist=fgmres(Aop, W*lv.b, Mop, m=m, tol=tol, maxit=maxit)
    return lv.E@(W*xh), its, hist, nV[0]
