"""CPU oracle for the PyLamp per-timestep hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the reference algorithm
(larskaislaniemi/PyLamp).  It is the *checker* for the CUDA path, never the
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
``pylamp_b200`` must never import anything from ``oracle/``.

Pinning.  The reference ships no tests, golden vectors or fixtures,
so the oracle is pinned against the *reference itself*: ``oracle/make_golden.py``
imports the unmodified reference modules from ``/root/reference`` (with the shims
in ``oracle/ref_shims.py``), runs them on seeded inputs and commits the outputs
under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function
below against those vectors (bit-exact for the marker routines and the matrix
assembly).  The linear solve is SciPy's vendored SuperLU (``spsolve``), a
third-party dependency the reference does not pin (no requirements file); the
oracle calls the same SciPy and adds one step of iterative refinement
 -- the distance to raw ``spsolve`` is reported by the tests.

Each function cites the reference file:line it follows.  Layout conventions are
the reference's: arrays are float64, C-contiguous, index [i (z), j (x)];
``tr_x`` is (M,2) [z,x]; ``tr_f`` is (M,13) with the TR_* columns.
"""

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

# --------------------------------------------------------------------------------------
# constants: pylamp_const.py:6-46
# --------------------------------------------------------------------------------------
DIM = 2
IZ, IX = 0, 1
IP = DIM
G = [9.81, 0]
SECINYR = 60 * 60 * 24 * 365.25
SECINKYR = SECINYR * 1e3
SECINMYR = SECINYR * 1e6
GASR = 8.31446
NFTRAC = 13
(TR_RHO, TR_ETA, TR_MRK, TR_TMP, TR_HCD, TR_HCP, TR_RH0, TR_ALP, TR_MAT, TR_ACE, TR_ET0,
 TR_IHT, TR__ID) = range(13)
EPS = 2 ** (-10)

# pylamp_trac.py:11-22
INTERP_AVG_ARITHMETIC = 1
INTERP_AVG_GEOMETRIC = 2
INTERP_AVG_WEIGHTED = 4
INTERP_AVG_ARITHW = 5
INTERP_AVG_GEOMW = 6
INTERP_METHOD_ELEM = 4
INTERP_METHOD_NEAREST = 8
INTERP_METHOD_LINEAR = 16
INTERP_METHOD_VELDIV = 32

# pylamp_stokes.py:17-20, pylamp_diff.py:12-13
BC_TYPE_NOSLIP = 0
BC_TYPE_FREESLIP = 1
BC_TYPE_CYCLIC = 2
BC_TYPE_FLOWTHRU = 4
BC_TYPE_FIXTEMP = 0
BC_TYPE_FIXFLOW = 1


# --------------------------------------------------------------------------------------
# marker -> grid    (pylamp_trac.py:161-318, INTERP_METHOD_ELEM branch)
# --------------------------------------------------------------------------------------
def _extended_axis(coords, lo, hi):
    """Ghost-node extension of one grid axis, pylamp_trac.py:207-217.

    Returns (axis, n_added_left, n_added_right)."""
    ax = np.array(coords, dtype=np.float64, copy=True)
    nleft = nright = 0
    while lo < ax[0]:
        ax = np.concatenate([[ax[0] - (ax[1] - ax[0])], ax])
        nleft += 1
    while hi > ax[-1]:
        ax = np.concatenate([ax, [ax[-1] + (ax[-1] - ax[-2])]])
        nright += 1
    return ax, nleft, nright


def marker_cell(coord, axis):
    """floor((n-1)*(x-Lmin)/L) -- pylamp_trac.py:46-47, 226-227.  IEEE mul then div."""
    n = axis.shape[0]
    lmin = axis[0]
    length = axis[-1] - axis[0]
    return np.floor((n - 1) * (coord - lmin) / length).astype(np.int64)


def trac2grid(tr_x, tr_f, mesh, grid, gridfield, nx, distweight=None, avgscheme=None,
              method=INTERP_METHOD_ELEM, debug=False):
    """Marker-to-node averaging; mutates ``gridfield[k][:, :]``.  pylamp_trac.py:161-318.

    ``np.add.at`` in the reference accumulates corner 0 for all markers, then corner 1,
    ... sequentially into one array (257-260, 276-279); ``np.bincount`` over the
    concatenated corner index lists adds in exactly that order, so sums are bit-identical.
    """
    assert len(gridfield) == tr_f.shape[1]
    if avgscheme is None:
        avgscheme = [INTERP_AVG_ARITHW] * len(gridfield)
    assert isinstance(avgscheme, list) and len(avgscheme) == len(gridfield)
    if not (method & INTERP_METHOD_ELEM):
        raise NotImplementedError("oracle restates the ELEM method only")

    axes, nl, nr = [], [], []
    for d in range(DIM):
        ax, a, b = _extended_axis(grid[d], np.min(tr_x[:, d]), np.max(tr_x[:, d]))
        axes.append(ax), nl.append(a), nr.append(b)
    nz, nxx = axes[IZ].shape[0], axes[IX].shape[0]

    ie = marker_cell(tr_x[:, IZ], axes[IZ])
    je = marker_cell(tr_x[:, IX], axes[IX])
    az = (tr_x[:, IZ] - axes[IZ][ie]) / (axes[IZ][ie + 1] - axes[IZ][ie])   # :247
    ax_ = (tr_x[:, IX] - axes[IX][je]) / (axes[IX][je + 1] - axes[IX][je])
    bz = 1 - az                                                            # :249
    bx = 1 - ax_
    # corner order of the reference: (i,j), (i+1,j), (i,j+1), (i+1,j+1)      :252-260
    flat = np.concatenate([ie * nxx + je, (ie + 1) * nxx + je, ie * nxx + je + 1,
                           (ie + 1) * nxx + je + 1])
    nnode = nz * nxx
    any_w = any(s & INTERP_AVG_WEIGHTED for s in avgscheme)
    any_c = any(not (s & INTERP_AVG_WEIGHTED) for s in avgscheme)
    if any_w:
        w = [(1 - ax_) * (1 - az), (1 - ax_) * (1 - bz), (1 - bx) * (1 - az), (1 - bx) * (1 - bz)]
        wsum = np.bincount(flat, weights=np.concatenate(w), minlength=nnode).reshape(nz, nxx)
    if any_c:
        cnt = np.bincount(flat, minlength=nnode).astype(np.float64).reshape(nz, nxx)

    for k, scheme in enumerate(avgscheme):
        f = tr_f[:, k]
        if scheme & INTERP_AVG_ARITHMETIC:
            val = f
        elif scheme & INTERP_AVG_GEOMETRIC:
            with np.errstate(divide="ignore", invalid="ignore"):
                val = np.log(f)
        else:
            print("!!! ERROR INVALID AVERAGING SCHEME")                     # :309
            continue
        if scheme & INTERP_AVG_WEIGHTED:
            contrib = np.concatenate([val * w[0], val * w[1], val * w[2], val * w[3]])
            den = wsum
        else:
            contrib = np.concatenate([val, val, val, val])
            den = cnt
        s = np.bincount(flat, weights=contrib, minlength=nnode).reshape(nz, nxx)
        with np.errstate(divide="ignore", invalid="ignore"):
            if scheme & INTERP_AVG_GEOMETRIC and not (scheme & INTERP_AVG_ARITHMETIC):
                s[np.isinf(s)] = 0                                          # :301
                out = np.exp(s / den)                                       # :304, 306
            else:
                out = s / den                                               # :281, 287
        gridfield[k][:, :] = out[nl[IZ]:nz - nr[IZ], nl[IX]:nxx - nr[IX]]    # :313-316
    return


# --------------------------------------------------------------------------------------
# grid -> marker    (pylamp_trac.py:30-158)
# --------------------------------------------------------------------------------------
def grid2trac(tr_x, tr_f, grid, gridfield, nx, defval=np.nan, method=INTERP_METHOD_LINEAR,
              stopOnError=False):
    """Bilinear (LINEAR) / divergence-conserving (VELDIV) / NEAREST interpolation of node
    fields to markers; writes ``tr_f[:, k]`` in place.  pylamp_trac.py:30-158."""
    assert len(gridfield) == tr_f.shape[1]
    assert method & (INTERP_METHOD_LINEAR | INTERP_METHOD_NEAREST | INTERP_METHOD_VELDIV)
    gz, gx = grid[IZ], grid[IX]
    z, x = tr_x[:, IZ], tr_x[:, IX]
    ie = np.floor((nx[IZ] - 1) * (z - gz[0]) / (gz[-1] - gz[0])).astype(np.int64)    # :46
    je = np.floor((nx[IX] - 1) * (x - gx[0]) / (gx[-1] - gx[0])).astype(np.int64)    # :47
    bad = (ie < 0) | (ie > nx[IZ] - 1) | (je < 0) | (je > nx[IX] - 1)               # :52
    nbad = int(np.sum(bad))
    if stopOnError and nbad > 0:
        raise Exception("stopOnError in grid2trac")
    if nbad > 0:
        print("!!! Warning, grid2trac(): Using default value for extrapolation in ", nbad, "tracers")
    ie[bad] = 0
    je[bad] = 0
    # signed distances to the low and high node of the cell, :71-75
    dz0, dz1 = z - gz[ie], -(z - gz[ie + 1])
    dx0, dx1 = x - gx[je], -(x - gx[je + 1])
    nfield = len(gridfield)

    if method & INTERP_METHOD_NEAREST:                                             # :77-85
        d2 = np.stack([dz0 ** 2 + dx0 ** 2, dz0 ** 2 + dx1 ** 2, dz1 ** 2 + dx0 ** 2,
                       dz1 ** 2 + dx1 ** 2], axis=1)
        c = np.argmin(d2, axis=1)
        for k in range(nfield):
            tr_f[:, k] = gridfield[k][ie + c // 2, je + c % 2]
            tr_f[bad, k] = defval
        return

    dxn = dx0 / (dx0 + dx1)                                                        # :89
    dzn = dz0 / (dz0 + dz1)                                                        # :90

    def bilin(f):                                                                  # :92-96
        return ((1 - dxn) * (1 - dzn) * f[ie, je] + dxn * (1 - dzn) * f[ie, je + 1]
                + (1 - dxn) * dzn * f[ie + 1, je] + dxn * dzn * f[ie + 1, je + 1])

    if method & INTERP_METHOD_LINEAR:
        for k in range(nfield):
            tr_f[:, k] = bilin(gridfield[k])
            tr_f[bad, k] = defval
        return

    # VELDIV: Meyer & Jenny (2004) correction, :98-154
    if nfield != 2:
        raise Exception("grid2trac(): method INTERP_METHOD_VELDIV only works in 2D and "
                        "expects field to be (vz,vx)")
    fz, fx = gridfield[IZ], gridfield[IX]
    hz = (gz[1:] - gz[:-1])[ie]
    hx = (gx[1:] - gx[:-1])[je]
    c10 = (0.5 * hx / hz) * (fz[ie, je] - fz[ie + 1, je] + fz[ie + 1, je + 1] - fz[ie, je + 1])
    c20 = (0.5 * hz / hx) * (fx[ie, je] - fx[ie, je + 1] + fx[ie + 1, je + 1] - fx[ie + 1, je])
    ux = bilin(fx)
    uz = bilin(fz)
    tr_f[:, IX] = ux + dxn * (1 - dxn) * c10                                       # :150,153
    tr_f[:, IZ] = uz + dzn * (1 - dzn) * c20                                       # :151,154
    tr_f[bad, :] = defval
    return


def RK(tr_x, grids, vels, nx, tstep, order=4):
    """4-stage marker advection with the reference's unweighted (1/6)(k1+k2+k3+k4) update.
    pylamp_trac.py:321-388 (order 4 only; the order-2 branch of the reference is dead)."""
    if order != 2 and order != 4:
        raise Exception("Sorry, don't know how to do that")
    if len(nx) != 2:
        raise Exception("Sorry, only 2D supported at the moment")
    if order == 2:
        raise NotImplementedError("reference RK2 branch references undefined names (332-345)")
    nx1 = [nx[IZ] + 1, nx[IX] + 1]
    k = []
    loc = tr_x
    for stage in range(4):
        v = np.zeros((tr_x.shape[0], DIM))
        grid2trac(loc, v, grids, vels, nx1, defval=0, method=INTERP_METHOD_VELDIV)
        k.append(v)
        if stage < 3:
            h = 0.5 if stage < 2 else 1.0
            loc = np.empty_like(tr_x)
            for d in range(DIM):
                loc[:, d] = tr_x[:, d] + (h * tstep if h != 1.0 else tstep) * v[:, d]   # :366,372,378
    xf = np.empty_like(tr_x)
    vf = np.empty_like(tr_x)
    for d in range(DIM):
        xf[:, d] = tr_x[:, d] + (1 / 6) * tstep * (k[0][:, d] + k[1][:, d] + k[2][:, d] + k[3][:, d])
        vf[:, d] = (xf[:, d] - tr_x[:, d]) / tstep
    return vf, xf


# --------------------------------------------------------------------------------------
# Stokes system    (pylamp_stokes.py:22-35, 86-101, 104-563)
# --------------------------------------------------------------------------------------
def stokes_gidx(idxs, nx, dim=DIM):
    """pylamp_stokes.py:22-35"""
    if len(idxs) != dim:
        raise Exception("num of idxs != dimensions")
    return idxs[IZ] * nx[IX] * (dim + 1) + idxs[IX] * (dim + 1)


def x2vp(x, nx):
    """pylamp_stokes.py:86-101 -- pressure is returned as P/Kcont, ghosts kept."""
    nxt = tuple(nx)
    return [x[0::3].reshape(nxt), x[1::3].reshape(nxt)], x[2::3].reshape(nxt)


def stokes_scaling(grid, f_etas, f_etan):
    """Kcont, Kbond of pylamp_stokes.py:116-122 (note L/n, not L/(n-1))."""
    mineta = min(np.min(f_etas), np.min(f_etan))
    avgdx = (grid[IX][-1] - grid[IX][0]) / grid[IX].shape[0]
    avgdz = (grid[IZ][-1] - grid[IZ][0]) / grid[IZ].shape[0]
    return 2 * mineta / (avgdx + avgdz), 4 * mineta / (avgdx + avgdz) ** 2


class _Triplets:
    def __init__(self):
        self.r, self.c, self.v = [], [], []

    def add(self, rows, cols, vals):
        rows = np.asarray(rows, dtype=np.int64)
        rows, cols, vals = np.broadcast_arrays(rows, np.asarray(cols, dtype=np.int64),
                                               np.asarray(vals, dtype=np.float64))
        self.r.append(rows.ravel()), self.c.append(cols.ravel()), self.v.append(vals.ravel())

    def tocsr(self, n, drop_rows=()):
        r, c, v = np.concatenate(self.r), np.concatenate(self.c), np.concatenate(self.v)
        if len(drop_rows):
            keep = ~np.isin(r, np.asarray(drop_rows))
            r, c, v = r[keep], c[keep], v[keep]
        return scipy.sparse.csr_matrix((v, (r, c)), shape=(n, n))


def makeStokesMatrix(nx, grid, f_etas, f_etan, f_rho, bc, surfstab=False, tstep=None,
                     surfstab_theta=0.5):
    """Assemble the reference's (3N x 3N) Stokes/continuity system.  pylamp_stokes.py:104-563.

    Returns (A csr_matrix, rhs).  The reference returns a lil_matrix; values and pattern
    are identical (checked bit-exactly against tests/golden).  Every row class of the
    reference is written exactly once (its ``lc`` counter, :112, :555-561), so a triplet
    list reproduces the lil overwrite semantics, apart from the anchor row which the
    reference zeroes first (:548) -- handled by dropping that row's earlier entries.
    """
    nz, nxx = int(nx[IZ]), int(nx[IX])
    gz, gx = np.asarray(grid[IZ], dtype=np.float64), np.asarray(grid[IX], dtype=np.float64)
    dof = nz * nxx * 3
    Kc, Kb = stokes_scaling(grid, f_etas, f_etan)
    rhs = np.zeros(dof)
    T = _Triplets()
    bz0, bx0, bz1, bx1 = bc[DIM * 0 + IZ], bc[DIM * 0 + IX], bc[DIM * 1 + IZ], bc[DIM * 1 + IX]

    def g(i, j, eq):
        return (i * nxx + j) * 3 + eq

    ii, jj = np.arange(nz), np.arange(nxx)
    # ghosts, :128-152
    T.add(g(ii, nxx - 1, IZ), g(ii, nxx - 1, IZ), Kc)
    T.add(g(ii, nxx - 1, IP), g(ii, nxx - 1, IP), Kc)
    T.add(g(nz - 1, jj, IX), g(nz - 1, jj, IX), Kc)
    T.add(g(nz - 1, jj[:-1], IP), g(nz - 1, jj[:-1], IP), Kc)

    # z-walls, :158-233 (tests use ==)
    for wall, b in ((0, bz0), (1, bz1)):
        j = np.arange(1, nxx - 1)
        if wall == 0:
            i, i1, i2, ib, iw = 0, 1, 2, 0, 0          # row i; neighbour rows; wall node
        else:
            i, i1, i2, ib, iw = nz - 2, nz - 3, None, nz - 1, nz - 1
        if b == BC_TYPE_NOSLIP:                        # :163-168, :202-207
            if wall == 0:
                d2, d1 = gz[2] - gz[0], gz[1] - gz[0]
            else:
                d2, d1 = gz[nz - 3] - gz[nz - 1], gz[nz - 2] - gz[nz - 1]
            T.add(g(i, j, IX), g(i, j, IX), Kc * (-1 / d2 + (-1) / d1))
            T.add(g(i, j, IX), g(i1, j, IX), Kc * (1 / d2))
        elif b == BC_TYPE_FREESLIP:                    # :170-175, :209-214
            T.add(g(i, j, IX), g(i, j, IX), Kc)
            T.add(g(i, j, IX), g(i1, j, IX), -Kc)
        elif b == BC_TYPE_CYCLIC:
            raise NotImplementedError("CYCLIC walls: SURVEY.md §8f-4 (next)")
        else:
            raise NotImplementedError("unsupported z-wall BC %r" % (b,))
        j = np.arange(0, nxx - 1)                       # vz = 0, :190-194, :229-233
        T.add(g(iw, j, IZ), g(iw, j, IZ), Kc)

    # x-walls, :237-326 (tests use &, so NOSLIP=0 never matches: rows stay empty)
    for wall, b in ((0, bx0), (1, bx1)):
        i = np.arange(1, nz - 1)
        if b & BC_TYPE_FREESLIP:                       # :249-255, :296-301
            if wall == 0:
                T.add(g(i, 0, IZ), g(i, 0, IZ), Kc)
                T.add(g(i, 0, IZ), g(i, 1, IZ), -Kc)
            else:
                T.add(g(i, nxx - 2, IZ), g(i, nxx - 2, IZ), Kc)
                T.add(g(i, nxx - 2, IZ), g(i, nxx - 3, IZ), -Kc)
        elif b & BC_TYPE_CYCLIC:
            raise NotImplementedError("CYCLIC walls: the reference's own matrix is singular (tests/test_reference_bc_probe.py)")
        elif b & BC_TYPE_FLOWTHRU:
            raise NotImplementedError("FLOWTHRU without FREESLIP leaves the vz wall rows empty in the reference (singular)")
        # (b == NOSLIP: the reference writes nothing here -> singular system, quirk 4)
        i = np.arange(0, nz - 1)
        jw = 0 if wall == 0 else nxx - 1
        if b & BC_TYPE_FLOWTHRU:                        # dvx/dx = 0, :268-273, :314-319
            jn, sgn = (1, -1.0) if wall == 0 else (nxx - 2, 1.0)
            T.add(g(i, jw, IX), g(i, jw, IX), sgn * Kc)
            T.add(g(i, jw, IX), g(i, jn, IX), -sgn * Kc)
        else:                                           # vx = 0, :277-281, :322-326
            T.add(g(i, jw, IX), g(i, jw, IX), Kc)

    # continuity: boundary cells without corners (:333-354) and interior (:496-518)
    ci, cj = np.meshgrid(np.arange(0, nz - 1), np.arange(0, nxx - 1), indexing="ij")
    corner = ((ci == 0) | (ci == nz - 2)) & ((cj == 0) | (cj == nxx - 2))
    i, j = ci[~corner], cj[~corner]
    rows = g(i, j, IP)
    T.add(rows, g(i, j + 1, IX), Kc / (gx[j + 1] - gx[j]))
    T.add(rows, g(i, j, IX), -Kc / (gx[j + 1] - gx[j]))
    T.add(rows, g(i + 1, j, IZ), Kc / (gz[i + 1] - gz[i]))
    T.add(rows, g(i, j, IZ), -Kc / (gz[i + 1] - gz[i]))

    # corner cells: horizontal pressure symmetry, :358-369
    for i in (0, nz - 2):
        T.add(g(i, 0, IP), g(i, 1, IP), Kb)
        T.add(g(i, 0, IP), g(i, 0, IP), -Kb)
        T.add(g(i, nxx - 2, IP), g(i, nxx - 3, IP), Kb)
        T.add(g(i, nxx - 2, IP), g(i, nxx - 2, IP), -Kb)

    if surfstab and tstep is None:
        raise Exception("surface stabilization needs predetermined tstep")

    # interior z-momentum, :376-429
    i, j = [a.ravel() for a in np.meshgrid(np.arange(1, nz - 1), np.arange(1, nxx - 2), indexing="ij")]
    rows = g(i, j, IZ)
    dzc = gz[i + 1] - gz[i - 1]
    dzp, dzm = gz[i + 1] - gz[i], gz[i] - gz[i - 1]
    dxj = gx[j + 1] - gx[j]
    dxp2, dxm2 = gx[j + 2] - gx[j], gx[j + 1] - gx[j - 1]
    diag = (-4 * f_etan[i, j] / dzp / dzc + -4 * f_etan[i - 1, j] / dzm / dzc
            + -2 * f_etas[i, j + 1] / dxp2 / dxj + -2 * f_etas[i, j] / dxm2 / dxj)
    offd_vx = np.zeros_like(diag)
    if surfstab:                                       # :422-426
        offd_vx = surfstab_theta * tstep * G[IZ] * 0.5 * (f_rho[i, j + 1] + f_rho[i + 1, j + 1] - f_rho[i, j - 1] - f_rho[i + 1, j - 1]) / (gx[j + 1] - gx[j - 1])
        diag = diag + surfstab_theta * tstep * G[IZ] * 0.5 * (f_rho[i + 1, j] + f_rho[i + 1, j + 1] - f_rho[i - 1, j] - f_rho[i - 1, j + 1]) / (gz[i + 1] - gz[i - 1])
    T.add(rows, g(i, j, IZ), diag)
    T.add(rows, g(i + 1, j, IZ), 4 * f_etan[i, j] / dzp / dzc)
    T.add(rows, g(i - 1, j, IZ), 4 * f_etan[i - 1, j] / dzm / dzc)
    T.add(rows, g(i, j + 1, IZ), 2 * f_etas[i, j + 1] / dxp2 / dxj)
    T.add(rows, g(i, j - 1, IZ), 2 * f_etas[i, j] / dxm2 / dxj)
    T.add(rows, g(i, j + 1, IX), 2 * f_etas[i, j + 1] / dzc / dxj)
    T.add(rows, g(i - 1, j + 1, IX), -2 * f_etas[i, j + 1] / dzc / dxj)
    T.add(rows, g(i, j, IX), -2 * f_etas[i, j] / dzc / dxj + offd_vx)
    T.add(rows, g(i - 1, j, IX), 2 * f_etas[i, j] / dzc / dxj)
    T.add(rows, g(i, j, IP), -2 * Kc / dzc)
    T.add(rows, g(i - 1, j, IP), 2 * Kc / dzc)
    rhs[rows] = -0.5 * (f_rho[i, j] + f_rho[i, j + 1]) * G[IZ]

    # interior x-momentum, :435-490
    i, j = [a.ravel() for a in np.meshgrid(np.arange(1, nz - 2), np.arange(1, nxx - 1), indexing="ij")]
    rows = g(i, j, IX)
    dxc = gx[j + 1] - gx[j - 1]
    dxp, dxm = gx[j + 1] - gx[j], gx[j] - gx[j - 1]
    dzi = gz[i + 1] - gz[i]
    dzp2, dzm2 = gz[i + 2] - gz[i], gz[i + 1] - gz[i - 1]
    diag = (-4 * f_etan[i, j] / dxp / dxc + -4 * f_etan[i, j - 1] / dxm / dxc
            + -2 * f_etas[i + 1, j] / dzp2 / dzi + -2 * f_etas[i, j] / dzm2 / dzi)
    offd_vz = np.zeros_like(diag)
    if surfstab:                                       # :483-487
        diag = diag + surfstab_theta * tstep * G[IX] * 0.5 * (f_rho[i, j + 1] + f_rho[i + 1, j + 1] - f_rho[i, j - 1] - f_rho[i + 1, j - 1]) / (gx[j + 1] - gx[j - 1])
        offd_vz = surfstab_theta * tstep * G[IX] * 0.5 * (f_rho[i + 1, j] + f_rho[i + 1, j + 1] - f_rho[i - 1, j] - f_rho[i - 1, j + 1]) / (gz[i + 1] - gz[i - 1])
    T.add(rows, g(i, j, IX), diag)
    T.add(rows, g(i, j + 1, IX), 4 * f_etan[i, j] / dxp / dxc)
    T.add(rows, g(i, j - 1, IX), 4 * f_etan[i, j - 1] / dxm / dxc)
    T.add(rows, g(i + 1, j, IX), 2 * f_etas[i + 1, j] / dzp2 / dzi)
    T.add(rows, g(i - 1, j, IX), 2 * f_etas[i, j] / dzm2 / dzi)
    T.add(rows, g(i + 1, j, IZ), 2 * f_etas[i + 1, j] / dxc / dzi)
    T.add(rows, g(i + 1, j - 1, IZ), -2 * f_etas[i + 1, j] / dxc / dzi)
    T.add(rows, g(i, j, IZ), -2 * f_etas[i, j] / dxc / dzi + offd_vz)
    T.add(rows, g(i, j - 1, IZ), 2 * f_etas[i, j] / dxc / dzi)
    T.add(rows, g(i, j, IP), -2 * Kc / dxc)
    T.add(rows, g(i, j - 1, IP), 2 * Kc / dxc)
    rhs[rows] = -0.5 * (f_rho[i, j] + f_rho[i + 1, j]) * G[IX]

    # pressure anchor, :525-551: cell (3,2), or mid-height on a flow-through x-wall (the last such wall in the
    # reference's loop order wins; for the x = L wall that is the GHOST pressure column -- no real anchor)
    if bz0 & BC_TYPE_FLOWTHRU or bz1 & BC_TYPE_FLOWTHRU:
        raise Exception("flow bnd condition in IZ dir no implemented")       # :546
    anchor = g(3, 2, IP)
    if bx1 & BC_TYPE_FLOWTHRU:
        anchor = g(int(nz / 2), nxx - 1, IP)
    elif bx0 & BC_TYPE_FLOWTHRU:
        anchor = g(int(nz / 2), 0, IP)
    A = T.tocsr(dof, drop_rows=[anchor])
    A = A + scipy.sparse.csr_matrix(([Kc], ([anchor], [anchor])), shape=(dof, dof))
    return A.tocsr(), rhs


# --------------------------------------------------------------------------------------
# energy system    (pylamp_diff.py:15-28, 78-83, 85-183)
# --------------------------------------------------------------------------------------
def x2t(x, nx):
    """pylamp_diff.py:78-83"""
    return x.reshape(tuple(nx))


def makeDiffusionMatrix(nx, grid, gridmp, f_T, f_k, f_Cp, f_rho, f_H, bc, bcvalue, tstep):
    """Implicit-Euler heat conduction system (N x N).  pylamp_diff.py:85-183."""
    nz, nxx = int(nx[IZ]), int(nx[IX])
    gz, gx = np.asarray(grid[IZ]), np.asarray(grid[IX])
    mz, mx = np.asarray(gridmp[IZ]), np.asarray(gridmp[IX])
    kz, kx = f_k[IZ], f_k[IX]
    dof = nz * nxx
    rhs = np.zeros(dof)
    T = _Triplets()

    def g(i, j):
        return i * nxx + j

    jj = np.arange(nxx)
    for wall, i in ((0, 0), (1, nz - 1)):               # z-walls own the corners, :99-124
        b, val = bc[DIM * wall + IZ], bcvalue[DIM * wall + IZ]
        if b == BC_TYPE_FIXTEMP:
            T.add(g(i, jj), g(i, jj), 1.0)
        elif b == BC_TYPE_FIXFLOW:
            if wall == 0:
                T.add(g(i, jj), g(i + 1, jj), kz[i, jj] / (gz[i + 1] - gz[i]))
                T.add(g(i, jj), g(i, jj), -kz[i, jj] / (gz[i + 1] - gz[i]))
            else:
                T.add(g(i, jj), g(i, jj), kz[i - 1, jj] / (gz[i] - gz[i - 1]))
                T.add(g(i, jj), g(i - 1, jj), -kz[i - 1, jj] / (gz[i] - gz[i - 1]))
        rhs[g(i, jj)] = val
    ii = np.arange(1, nz - 1)
    for wall, j in ((0, 0), (1, nxx - 1)):              # x-walls, :126-152
        b, val = bc[DIM * wall + IX], bcvalue[DIM * wall + IX]
        if b == BC_TYPE_FIXTEMP:
            T.add(g(ii, j), g(ii, j), 1.0)
        elif b == BC_TYPE_FIXFLOW:
            if wall == 0:
                T.add(g(ii, j), g(ii, j + 1), kx[ii, j] / (gx[j + 1] - gx[j]))
                T.add(g(ii, j), g(ii, j), -kx[ii, j] / (gx[j + 1] - gx[j]))
            else:
                T.add(g(ii, j), g(ii, j), kx[ii, j - 1] / (gx[j] - gx[j - 1]))
                T.add(g(ii, j), g(ii, j - 1), -kx[ii, j - 1] / (gx[j] - gx[j - 1]))
        rhs[g(ii, j)] = val

    i, j = [a.ravel() for a in np.meshgrid(np.arange(1, nz - 1), np.arange(1, nxx - 1), indexing="ij")]
    rows = g(i, j)
    pre = tstep / (f_rho[i, j] * f_Cp[i, j])            # :165
    dxe, dxw, dxm_ = gx[j + 1] - gx[j], gx[j] - gx[j - 1], mx[j] - mx[j - 1]
    dzs, dzn, dzm_ = gz[i + 1] - gz[i], gz[i] - gz[i - 1], mz[i] - mz[i - 1]
    T.add(rows, g(i, j + 1), pre * kx[i, j] / dxe / dxm_)
    T.add(rows, g(i, j - 1), pre * kx[i, j - 1] / dxw / dxm_)
    T.add(rows, g(i + 1, j), pre * kz[i, j] / dzs / dzm_)
    T.add(rows, g(i - 1, j), pre * kz[i - 1, j] / dzn / dzm_)
    T.add(rows, g(i, j), pre * (-kx[i, j] / dxe / dxm_ + -kx[i, j - 1] / dxw / dxm_
                                + -kz[i, j] / dzs / dzm_ + -kz[i - 1, j] / dzn / dzm_) - 1)
    rhs[rows] = -f_T[i, j] - tstep * f_H[i, j] / (f_rho[i, j] * f_Cp[i, j])      # :179
    return T.tocsr(dof), rhs


# --------------------------------------------------------------------------------------
# the solve: scipy.sparse.linalg.spsolve at pylamp2.py:360, 394, 419
# --------------------------------------------------------------------------------------
def spsolve(A, rhs):
    """What the reference driver calls: SuperLU via scipy (pylamp2.py:360)."""
    return scipy.sparse.linalg.spsolve(scipy.sparse.csc_matrix(A), rhs)


def solve_refined(A, rhs, steps=1):
    """Ground truth for solver parity: SuperLU + ``steps`` fp64 refinement steps.
    Raw spsolve is only reproducible to 1e-6..1e-8 on stiff systems."""
    A = scipy.sparse.csc_matrix(A)
    lu = scipy.sparse.linalg.splu(A)
    x = lu.solve(rhs)
    for _ in range(steps):
        x = x + lu.solve(rhs - A @ x)
    return x


# --------------------------------------------------------------------------------------
# driver-inline steps of the loop body, pylamp2.py:273-594
# --------------------------------------------------------------------------------------
def make_grids(nx, L):
    """grid, mesh, gridmp, meshmp of pylamp2.py:87-97."""
    grid = [np.linspace(0, L[i], nx[i]) for i in range(DIM)]
    mesh = np.meshgrid(*grid, indexing="ij")
    gridmp = [(grid[i][1:nx[i]] + grid[i][0:(nx[i] - 1)]) / 2 for i in range(DIM)]
    for i in range(DIM):
        gridmp[i] = np.append(gridmp[i], gridmp[i][-1] + (gridmp[i][-1] - gridmp[i][-2]))
    meshmp = np.meshgrid(*gridmp, indexing="ij")
    return grid, mesh, gridmp, meshmp


def update_properties(tr_f, tdep_rho, tdep_eta, Tref, etamin, etamax):
    """pylamp2.py:291-303"""
    if tdep_rho:
        tr_f[:, TR_RHO] = ((tr_f[:, TR_ALP] * (tr_f[:, TR_TMP] - Tref) + 1) / tr_f[:, TR_RH0]) ** (-1)
    else:
        tr_f[:, TR_RHO] = tr_f[:, TR_RH0]
    if tdep_eta:
        tr_f[:, TR_ETA] = tr_f[:, TR_ET0] * np.exp(tr_f[:, TR_ACE] / (GASR * tr_f[:, TR_TMP])
                                                 - tr_f[:, TR_ACE] / (GASR * Tref))
        tr_f[tr_f[:, TR_ETA] < etamin, TR_ETA] = etamin
        tr_f[tr_f[:, TR_ETA] > etamax, TR_ETA] = etamax
    else:
        tr_f[:, TR_ETA] = tr_f[:, TR_ET0]


def clamp(v, lo, hi):
    return max(min(v, hi), lo)


def heat_timestep(f_kz, f_rho, f_Cp, dx, modifier, lo, hi):
    """pylamp2.py:339-343"""
    return clamp(modifier * np.min(dx) ** 2 / np.max(2 * (f_kz / (f_rho * f_Cp))), lo, hi)


def stokes_timestep(newvel, dx, modifier, lo, hi):
    """pylamp2.py:364-366 -- np.max over both components is *signed* (quirk 6)."""
    return clamp(modifier * np.min(dx) / np.max(newvel), lo, hi)


def centre_velocities(newvel, gridmp, nx, bcstokes):
    """Cell-centre velocities with a BC ghost ring, pylamp2.py:491-545.
    Returns ([newgridz, newgridx], [Vz, Vx]) with fields of shape (nz+1, nxx+1)."""
    nz, nxx = nx[IZ], nx[IX]
    vels = [np.zeros((nz + 1, nxx + 1)), np.zeros((nz + 1, nxx + 1))]
    vels[IZ][1:nz, 1:nxx] = 0.5 * (newvel[IZ][1:, :-1] + newvel[IZ][:-1, :-1])
    vels[IX][1:nz, 1:nxx] = 0.5 * (newvel[IX][:-1, 1:] + newvel[IX][:-1, :-1])
    pre = [gridmp[d][0] - (gridmp[d][1] - gridmp[d][0]) for d in range(DIM)]
    newgrid = [np.insert(gridmp[IZ], 0, pre[IZ]), np.insert(gridmp[IX], 0, pre[IX])]

    def ring(wall, axis):
        b = bcstokes[DIM * wall + axis]
        gh, inn, wrap = (0, 1, -2) if wall == 0 else (-1, -2, 1)
        sl = (lambda k: (k, slice(None))) if axis == IZ else (lambda k: (slice(None), k))
        tang, norm = (IX, IZ) if axis == IZ else (IZ, IX)
        if b & BC_TYPE_FREESLIP:
            vels[tang][sl(gh)] = vels[tang][sl(inn)]
            vels[norm][sl(gh)] = -vels[norm][sl(inn)]
        elif b & BC_TYPE_CYCLIC:
            vels[tang][sl(gh)] = vels[tang][sl(wrap)]
            vels[norm][sl(gh)] = vels[norm][sl(wrap)]
        # NOSLIP == 0 never matches '&' in the reference (:503, :513, :525, :535): ring stays 0
        if axis == IX and b & BC_TYPE_FLOWTHRU:
            vels[IX][sl(gh)] = vels[IX][sl(inn)]

    ring(0, IZ), ring(0, IX), ring(1, IZ), ring(1, IX)     # order of :503-545
    return newgrid, vels


def fence(tr_x, tr_f, L, bcstokes, enabled=True):
    """pylamp2.py:558-572 (cyclic test reproduces the d*0+IX indexing bug, quirk a-14)."""
    for d in range(DIM):
        if bcstokes[d * 0 + IX] & BC_TYPE_CYCLIC:
            tr_x[tr_x[:, d] <= 0, d] += L[d]
            tr_x[tr_x[:, d] >= L[d], d] -= L[d]
        else:
            idx = tr_x[:, d] <= 0
            if enabled and not (bcstokes[DIM * 0 + d] & BC_TYPE_FLOWTHRU):
                tr_x[idx, d] = EPS
            else:
                tr_f[idx, TR__ID] = -1
            idx = tr_x[:, d] >= L[d]
            if enabled and not (bcstokes[DIM * 1 + d] & BC_TYPE_FLOWTHRU):
                tr_x[idx, d] = L[d] - EPS
            else:
                tr_f[idx, TR__ID] = -1


def delete_outside(tr_x, tr_f, trac_vel):
    """pylamp2.py:573-581: markers that `fence` flagged (TR__ID = -1: beyond a wall with the fence
    disabled, or beyond a FLOWTHRU wall) are removed from tr_x, tr_f and trac_vel.
    Returns (tr_x, tr_f, trac_vel, number removed)."""
    outside = tr_f[:, TR__ID] < 0
    idx = np.where(outside)[0]
    return (np.delete(tr_x, idx, axis=0), np.delete(tr_f, idx, axis=0),
            np.delete(trac_vel, idx, axis=0) if trac_vel is not None else None, int(np.sum(outside)))


def cell_index_count(tr_x, nx, L):
    """Marker cell index and per-cell count, pylamp2.py:588-593.  BIT-EXACT parity item."""
    ielem = np.floor((nx[IZ] - 1) * tr_x[:, IZ] / L[IZ]).astype(np.int64)
    jelem = np.floor((nx[IX] - 1) * tr_x[:, IX] / L[IX]).astype(np.int64)
    kelem = ielem * (nx[IX] - 1) + jelem
    ncell = (nx[IZ] - 1) * (nx[IX] - 1)
    count = np.bincount(np.append(kelem, np.arange(ncell))) - 1
    return kelem, count


def subgrid_diffusion(tr_x, tr_f, old_T, mesh, grid, nx, dx, tstep):
    """pylamp2.py:471-480; returns the node correction field f_sgc as well."""
    d = 0.5
    dt0 = tr_f[:, TR_HCP] * tr_f[:, TR_RHO] / (tr_f[:, TR_HCD] * ((2 / dx[IX]) ** 2 + (2 / dx[IZ]) ** 2))
    Tsg = old_T - (old_T - tr_f[:, TR_TMP]) * np.exp(-d * tstep / dt0)
    dT = Tsg - tr_f[:, TR_TMP]
    back = np.zeros_like(dT)
    f_sgc = np.zeros(tuple(nx))
    trac2grid(tr_x, dT[:, None], mesh, grid, [f_sgc], nx, avgscheme=[INTERP_AVG_ARITHW])
    grid2trac(tr_x, back[:, None], grid, [f_sgc], nx, method=INTERP_METHOD_LINEAR, stopOnError=True)
    tr_f[:, TR_TMP] = Tsg - back
    return f_sgc


class Options:
    """The configurable locals of pylamp2.py:37-77 (defaults as shipped)."""
    def __init__(self, **kw):
        self.do_stokes = True
        self.do_advect = True
        self.do_heatdiff = True
        self.do_subgrid_heatdiff = True
        self.tstep_adv_max = 50e9 * SECINYR
        self.tstep_adv_min = 50e-9 * SECINYR
        self.tstep_dif_max = 50e9 * SECINYR
        self.tstep_dif_min = 50e-9 * SECINYR
        self.tstep_modifier = 0.67
        self.tdep_rho = True
        self.tdep_eta = True
        self.etamin = 1e17
        self.etamax = 1e23
        self.Tref = 1623
        self.tracs_fence_enabled = True
        self.surface_stabilization = False      # pylamp2.py:71-73
        self.surfstab_theta = 0.5
        self.surfstab_tstep = -1                # negative: the dynamic time step is used (re-solve loop)
        self.bcstokes = [BC_TYPE_FREESLIP] * 4
        self.bcheat = [BC_TYPE_FIXTEMP, BC_TYPE_FIXFLOW, BC_TYPE_FIXTEMP, BC_TYPE_FIXFLOW]
        self.bcheatvals = [273, 0, 1623, 0]
        self.solve = spsolve                     # what the reference calls
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class State:
    """All arrays the reference driver keeps as locals (pylamp2.py:100-127)."""
    def __init__(self, nx, L, tr_x, tr_f):
        self.nx, self.L = list(nx), list(L)
        self.dx = [L[i] / (nx[i] - 1) for i in range(DIM)]
        self.grid, self.mesh, self.gridmp, self.meshmp = make_grids(nx, L)
        z = lambda: np.zeros(tuple(nx))
        self.f_etas, self.f_T, self.f_rho, self.f_Cp, self.f_etan = z(), z(), z(), z(), z()
        self.f_k = [z(), z()]
        self.f_H, self.f_mat, self.f_sgc = z(), z(), z()
        self.tr_x, self.tr_f = tr_x, tr_f
        self.it, self.totaltime = 0, 0.0
        self.newvel, self.newpres, self.newtemp = None, None, None
        self.trac_vel, self.tstep, self.limiter = None, None, ""
        self.kelem, self.count = None, None


def timestep(s, o, timers=None):
    """One pass of the loop body pylamp2.py:273-594 (no injection, no output, NPROC=1).  ``timers`` (dict) accumulates per-phase seconds."""
    import time as _time
    t0 = [_time.perf_counter()]

    def lap(name):
        if timers is not None:
            t = _time.perf_counter()
            timers[name] = timers.get(name, 0.0) + (t - t0[0])
            t0[0] = t

    s.it += 1
    nx, grid, gridmp, mesh, meshmp = s.nx, s.grid, s.gridmp, s.mesh, s.meshmp
    tr_x, tr_f = s.tr_x, s.tr_f
    update_properties(tr_f, o.tdep_rho, o.tdep_eta, o.Tref, o.etamin, o.etamax)
    lap("properties")
    if o.do_advect and o.do_heatdiff:                                               # :307-313
        trac2grid(tr_x, tr_f[:, [TR_RHO, TR_ETA, TR_HCP, TR_TMP, TR_IHT, TR_MAT]], mesh, grid,
                  [s.f_rho, s.f_etas, s.f_Cp, s.f_T, s.f_H, s.f_mat], nx,
                  avgscheme=[INTERP_AVG_ARITHW, INTERP_AVG_GEOMW] + [INTERP_AVG_ARITHW] * 4)
        trac2grid(tr_x, tr_f[:, [TR_ETA]], meshmp, gridmp, [s.f_etan], nx, avgscheme=[INTERP_AVG_GEOMW])
        trac2grid(tr_x, tr_f[:, [TR_HCD]], [meshmp[IZ], mesh[IX]], [gridmp[IZ], grid[IX]], [s.f_k[IZ]],
                  nx, avgscheme=[INTERP_AVG_ARITHW])
        trac2grid(tr_x, tr_f[:, [TR_HCD]], [mesh[IZ], meshmp[IX]], [grid[IZ], gridmp[IX]], [s.f_k[IX]],
                  nx, avgscheme=[INTERP_AVG_ARITHW])
    elif o.do_advect:                                                               # :316-319
        trac2grid(tr_x, tr_f[:, [TR_RHO, TR_ETA]], mesh, grid, [s.f_rho, s.f_etas], nx,
                  avgscheme=[INTERP_AVG_ARITHW, INTERP_AVG_GEOMW])
        trac2grid(tr_x, tr_f[:, [TR_ETA]], meshmp, gridmp, [s.f_etan], nx,
                  avgscheme=[INTERP_AVG_GEOMETRIC])
    else:
        raise NotImplementedError("heat-only mode (pylamp2.py:321-331) is outside the hot path")
    lap("trac2grid")
    if o.do_heatdiff and s.it > 1:                                                  # :333-337
        s.f_T[:, 0], s.f_T[:, -1] = s.newtemp[:, 0], s.newtemp[:, -1]
        s.f_T[0, :], s.f_T[-1, :] = s.newtemp[0, :], s.newtemp[-1, :]
    if o.do_heatdiff:
        tstep_temp = heat_timestep(s.f_k[IZ], s.f_rho, s.f_Cp, s.dx, o.tstep_modifier,
                                   o.tstep_dif_min, o.tstep_dif_max)
    if not o.surface_stabilization or o.surfstab_tstep < 0:                         # :352-355
        A, rhs = makeStokesMatrix(nx, grid, s.f_etas, s.f_etan, s.f_rho, o.bcstokes)
    else:
        A, rhs = makeStokesMatrix(nx, grid, s.f_etas, s.f_etan, s.f_rho, o.bcstokes, surfstab=True,
                                  tstep=o.surfstab_tstep, surfstab_theta=o.surfstab_theta)
    lap("stokes_assembly")
    x = o.solve(A, rhs)                                                             # :360
    lap("stokes_solve")
    s.newvel, s.newpres = x2vp(x, nx)
    tstep_stokes = stokes_timestep(s.newvel, s.dx, o.tstep_modifier, o.tstep_adv_min, o.tstep_adv_max)
    if o.surfstab_tstep > 0:                                                        # :368-372
        tstep_stokes = o.surfstab_tstep
    if o.do_heatdiff:
        s.limiter = "H" if tstep_temp < tstep_stokes else "S"                      # :374-379
        tstep = min(tstep_temp, tstep_stokes)
    else:
        tstep, s.limiter = tstep_stokes, "S"
    s.stab_solves = 0
    if o.surface_stabilization and o.surfstab_tstep < 0:                            # :387-405
        while True:      # redo the solve with the stabilisation terms of the step actually taken
            A, rhs = makeStokesMatrix(nx, grid, s.f_etas, s.f_etan, s.f_rho, o.bcstokes, surfstab=True,
                                      tstep=tstep, surfstab_theta=o.surfstab_theta)
            x = o.solve(A, rhs)
            s.stab_solves += 1
            s.newvel, s.newpres = x2vp(x, nx)
            check = o.tstep_modifier * np.min(s.dx) / np.max(s.newvel)               # :399 (not clamped)
            if check < tstep:
                tstep, s.limiter = check, "Ss"
            else:
                break
        lap("stokes_solve")
    s.tstep = tstep
    s.totaltime += tstep
    lap("dt")
    if o.do_heatdiff:
        A, rhs = makeDiffusionMatrix(nx, grid, gridmp, s.f_T, s.f_k, s.f_Cp, s.f_rho, s.f_H,
                                     o.bcheat, o.bcheatvals, tstep)                 # :415
        lap("heat_assembly")
        newtemp = x2t(o.solve(A, rhs), nx)                                          # :419-421
        lap("heat_solve")
        old_T = tr_f[:, TR_TMP].copy()                                              # :436
        interp = np.zeros((tr_f.shape[0], 1))
        if s.it == 1:                                                               # :441-447
            grid2trac(tr_x, interp, grid, [newtemp], nx, method=INTERP_METHOD_LINEAR, stopOnError=True)
            tr_f[:, TR_TMP] = interp[:, 0]
        else:                                                                       # :448-480
            grid2trac(tr_x, interp, grid, [newtemp - s.f_T], nx, method=INTERP_METHOD_LINEAR,
                      stopOnError=True)
            tr_f[:, TR_TMP] = tr_f[:, TR_TMP] + interp[:, 0]
            if o.do_subgrid_heatdiff:
                s.f_sgc = subgrid_diffusion(tr_x, tr_f, old_T, mesh, grid, nx, s.dx, tstep)
        s.newtemp = newtemp
        lap("grid2trac_T")
    newgrid, vels = centre_velocities(s.newvel, gridmp, nx, o.bcstokes)             # :491-545
    s.trac_vel, s.tr_x = RK(tr_x, newgrid, vels, nx, tstep)                         # :550
    lap("advect")
    fence(s.tr_x, tr_f, s.L, o.bcstokes, o.tracs_fence_enabled)                      # :558-572
    s.tr_x, s.tr_f, s.trac_vel, s.removed = delete_outside(s.tr_x, tr_f, s.trac_vel)  # :573-581
    s.kelem, s.count = cell_index_count(s.tr_x, nx, s.L)                            # :588-593
    lap("fence_count")
    return s


def inject_markers(s, tracdens, tracdens_min):
    """Marker injection into under-populated cells, pylamp2.py:594-633 (host-side; uses the
    global NumPy Mersenne-Twister stream exactly like the reference).  ``s.kelem/count`` must
    be current.  Returns the number of injected markers."""
    nx, grid = s.nx, s.grid
    few = s.count < tracdens_min
    if np.sum(few) == 0:
        return 0
    ielem = s.kelem // (nx[IX] - 1)
    jelem = s.kelem % (nx[IX] - 1)
    kmiss = np.where(few)[0]
    imiss = np.floor(kmiss / (nx[IX] - 1)).astype(int)
    jmiss = (kmiss % (nx[IX] - 1)).astype(int)
    nmiss = tracdens - s.count[few]
    prev_tr_f = np.copy(s.tr_f)
    tr_x, tr_f = s.tr_x, s.tr_f
    for c in range(nmiss.size):
        xt = np.random.rand(nmiss[c], DIM)
        xt[:, IX] = xt[:, IX] * (grid[IX][jmiss[c] + 1] - grid[IX][jmiss[c]]) + grid[IX][jmiss[c]]
        xt[:, IZ] = xt[:, IZ] * (grid[IZ][imiss[c] + 1] - grid[IZ][imiss[c]]) + grid[IZ][imiss[c]]
        ft = np.zeros((nmiss[c], NFTRAC))
        maxid = np.max(tr_f[:, TR__ID])
        ft[:, TR__ID] = np.arange(maxid, maxid + nmiss[c])
        incell = (ielem == imiss[c]) & (jelem == jmiss[c])
        for k in range(NFTRAC):
            if k != TR__ID:
                with np.errstate(invalid="ignore", divide="ignore"):
                    ft[:, k] = np.sum(prev_tr_f[incell, k]) / np.sum(incell)
        tr_f = np.append(tr_f, ft, axis=0)
        tr_x = np.append(tr_x, xt, axis=0)
    s.tr_x, s.tr_f = tr_x, tr_f
    return int(np.sum(nmiss))
