"""Synthetic setups for the configs in BASELINE.json (host-side NumPy generators).

The reference has no config files: setups are hard-coded blocks in its driver
(/root/reference/pylamp2.py:116-253).  These generators restate the two that matter for
parity (model 5 as shipped = C1; model 1 = thermo-mechanical variant) and add the
benchmark setups of SURVEY.md §8d (SolCx C2, Rayleigh-Taylor C3, convection C4).
"""
import numpy as np

from .pylamp_const import *  # noqa: F401,F403

BC_FREESLIP = 1


def make_grids(nx, L):
    """grid / gridmp of pylamp2.py:87-95 (gridmp = midpoints + one point beyond the end)."""
    grid = [np.linspace(0, L[i], nx[i]) for i in range(DIM)]
    gridmp = [(grid[i][1:nx[i]] + grid[i][0:(nx[i] - 1)]) / 2 for i in range(DIM)]
    for i in range(DIM):
        gridmp[i] = np.append(gridmp[i], gridmp[i][-1] + (gridmp[i][-1] - gridmp[i][-2]))
    return grid, gridmp


def _blank_markers(nx, L, tracdens, seed):
    """pylamp2.py:116-127: ntrac = prod(nx)*tracdens counts NODES (quirk 10); legacy global
    RandomState so that the same seed reproduces the reference's marker cloud."""
    ntrac = int(np.prod(nx)) * tracdens
    rs = np.random.RandomState(seed)
    tr_x = np.multiply(rs.rand(ntrac, DIM), L)
    tr_f = np.zeros((ntrac, NFTRAC))
    tr_f[:, TR__ID] = np.arange(0, ntrac)
    return tr_x, tr_f


def c1_shipped(seed=1234):
    """BASELINE.json configs[0]: pylamp2.py as shipped (model 5, :37-39, :218-242)."""
    nx, L = [201, 41], [1, 0.2]
    tr_x, tr_f = _blank_markers(nx, L, 45, seed)
    tr_f[:, TR_RH0] = 1420
    tr_f[:, TR_MAT] = 1
    tr_f[:, TR_ET0] = 1e2
    idx = (tr_x[:, IX] - 0.1) ** 2 + (tr_x[:, IZ] - 0.2) ** 2 < 0.01 ** 2
    tr_f[idx, TR_RH0] = 1470
    tr_f[idx, TR_MAT] = 2
    tr_f[idx, TR_ET0] = 1e12
    opts = dict(do_heatdiff=False, tdep_rho=False, tdep_eta=False, bcstokes=[BC_FREESLIP] * 4)
    return nx, L, tr_x, tr_f, opts


def thermo_variant(seed=4321, nx=(33, 49), L=(660e3, 1000e3), tracdens=20):
    """Model 1 of pylamp2.py:136-170 (stagnant lid) with free-slip walls added: the
    thermo-mechanical path (Stokes + energy + subgrid diffusion + advection)."""
    nx, L = list(nx), list(L)
    tr_x, tr_f = _blank_markers(nx, L, tracdens, seed)
    zcrust = tr_x[:, IZ] < 50e3
    tr_f[:, TR_RH0] = 3300
    tr_f[:, TR_ALP] = 3.5e-5
    tr_f[:, TR_MAT] = 2
    tr_f[:, TR_ET0] = 1e20
    tr_f[:, TR_HCD] = 4.0
    tr_f[:, TR_HCP] = 1250
    tr_f[:, TR_TMP] = 1623
    tr_f[:, TR_ACE] = 120e3
    tr_f[:, TR_IHT] = 0.02e-6 / 3300
    tr_f[zcrust, TR_IHT] = 2.5e-6 / 2900
    tr_f[zcrust, TR_RH0] = 2900
    tr_f[zcrust, TR_MAT] = 1
    tr_f[zcrust, TR_ET0] = 1e22
    tr_f[zcrust, TR_HCD] = 2.5
    tr_f[zcrust, TR_HCP] = 1000
    tr_f[zcrust, TR_TMP] = 273
    opts = dict(do_heatdiff=True, tdep_rho=True, tdep_eta=True, bcstokes=[BC_FREESLIP] * 4)
    return nx, L, tr_x, tr_f, opts


def lattice_markers(ncz, ncx, L, per_side, seed, jitter=0.5):
    """``per_side``^2 markers per cell on a jittered sub-lattice, generated in cell-major
    order: no empty nodes, and the cloud starts cell-sorted."""
    rng = np.random.default_rng(seed)
    nsz, nsx = ncz * per_side, ncx * per_side
    z = (np.arange(nsz) + 0.5) / nsz
    x = (np.arange(nsx) + 0.5) / nsx
    # cell-major ordering: (cell i, cell j, sub i, sub j)
    zz = z.reshape(ncz, 1, per_side, 1)
    xx = x.reshape(1, ncx, 1, per_side)
    Z = np.broadcast_to(zz, (ncz, ncx, per_side, per_side)).reshape(-1).copy()
    X = np.broadcast_to(xx, (ncz, ncx, per_side, per_side)).reshape(-1).copy()
    Z += (rng.random(Z.shape) - 0.5) * jitter / nsz
    X += (rng.random(X.shape) - 0.5) * jitter / nsx
    tr_x = np.stack([Z * L[IZ], X * L[IX]], axis=1)
    tr_f = np.zeros((tr_x.shape[0], NFTRAC))
    tr_f[:, TR__ID] = np.arange(tr_x.shape[0])
    return tr_x, tr_f


def rayleigh_taylor(ncell=64, per_side=4, seed=7):
    """BASELINE.json configs[2] (C3): two isoviscous-contrast layers, cosine interface."""
    nx, L = [ncell + 1, ncell + 1], [1.0, 1.0]
    tr_x, tr_f = lattice_markers(ncell, ncell, L, per_side, seed)
    upper = tr_x[:, IZ] < 0.5 + 0.02 * np.cos(np.pi * tr_x[:, IX])
    tr_f[:, TR_RH0] = 3200
    tr_f[:, TR_ET0] = 1e20
    tr_f[:, TR_MAT] = 1
    tr_f[upper, TR_RH0] = 3300
    tr_f[upper, TR_ET0] = 1e21
    tr_f[upper, TR_MAT] = 2
    opts = dict(do_heatdiff=False, tdep_rho=False, tdep_eta=False, bcstokes=[BC_FREESLIP] * 4)
    return nx, L, tr_x, tr_f, opts


def convection(ncell=64, per_side=4, seed=11, Ra=1e6, Lbox=1e6):
    """BASELINE.json configs[3] (C4): Ra=1e6 convection, Arrhenius viscosity clipped to
    [1e17, 1e23] (pylamp2.py:62-63), T 273..1623 K (pylamp2.py:250-251), conductive profile +
    1 % sinusoidal perturbation."""
    nx, L = [ncell + 1, ncell + 1], [Lbox, Lbox]
    tr_x, tr_f = lattice_markers(ncell, ncell, L, per_side, seed)
    rho0, alpha, k, cp, Ea = 3300.0, 3.5e-5, 4.0, 1250.0, 120e3
    dT = 1623.0 - 273.0
    eta0 = rho0 * G[IZ] * alpha * dT * Lbox ** 3 / ((k / (rho0 * cp)) * Ra)
    zn, xn = tr_x[:, IZ] / Lbox, tr_x[:, IX] / Lbox
    tr_f[:, TR_TMP] = 273.0 + dT * zn + 0.01 * dT * np.sin(np.pi * zn) * np.cos(np.pi * xn)
    tr_f[:, TR_RH0] = rho0
    tr_f[:, TR_ALP] = alpha
    tr_f[:, TR_HCD] = k
    tr_f[:, TR_HCP] = cp
    tr_f[:, TR_ACE] = Ea
    tr_f[:, TR_ET0] = eta0
    tr_f[:, TR_MAT] = 1
    opts = dict(do_heatdiff=True, tdep_rho=True, tdep_eta=True, bcstokes=[BC_FREESLIP] * 4)
    return nx, L, tr_x, tr_f, opts


def solcx_fields(n, eta_right=1e6):
    """BASELINE.json configs[1] (C2, Duretz et al. 2011 SolCx-type): unit square, viscosity
    1 (x<0.5) / eta_right (x>=0.5) set directly on the node and centre grids, density
    rho = -sin(pi z) cos(pi x) / g so that rho*g is the SolCx body force."""
    nx, L = [n, n], [1.0, 1.0]
    grid, gridmp = make_grids(nx, L)
    zs, xs = np.meshgrid(grid[IZ], grid[IX], indexing="ij")
    zc, xc = np.meshgrid(gridmp[IZ], gridmp[IX], indexing="ij")
    etas = np.where(xs < 0.5, 1.0, eta_right)
    etan = np.where(xc < 0.5, 1.0, eta_right)
    rho = -np.sin(np.pi * zs) * np.cos(np.pi * xs) / G[IZ]
    return nx, L, grid, gridmp, etas, etan, rho


def convection_fields(ncell, t=0.0, Ra=1e6, Lbox=1e6, shift_cells=0.3):
    """C4-type coefficient fields given analytically on the node / centre grids (no markers), as a
    function of a step counter `t`: the temperature field of `convection` with a 10 % lateral
    perturbation plus a plume-like anomaly, translated by `shift_cells` cells per unit of t --
    i.e. what consecutive time steps of the loop hand to the Stokes solve.  Used for solver parity at
    sizes where only ONE direct solve is affordable (tests/golden/large_*.npz hold the reference's
    solution at t = 0; the GPU solver is stepped through t = -n..0 with its time-loop settings).
    Returns (nx, L, grid, gridmp, f_etas, f_etan, f_rho)."""
    nx, L = [ncell + 1, ncell + 1], [Lbox, Lbox]
    grid, gridmp = make_grids(nx, L)
    rho0, alpha, k, cp, Ea, Tref = 3300.0, 3.5e-5, 4.0, 1250.0, 120e3, 1623.0
    dT = 1623.0 - 273.0
    eta0 = rho0 * G[IZ] * alpha * dT * Lbox ** 3 / ((k / (rho0 * cp)) * Ra)
    s = shift_cells * t / ncell

    def temp(z, x):
        zn, xn = z / Lbox, x / Lbox - s
        T = 273.0 + dT * zn + 0.10 * dT * np.sin(np.pi * zn) * np.cos(2 * np.pi * xn)
        T = T + 0.25 * dT * np.exp(-((xn - 0.37) ** 2 + (zn - 0.55) ** 2) / 0.01) * np.sin(np.pi * zn)
        return np.clip(T, 273.0, 1900.0)

    def eta(T):
        return np.clip(eta0 * np.exp(Ea / (GASR * T) - Ea / (GASR * Tref)), 1e17, 1e23)

    zs, xs = np.meshgrid(grid[IZ], grid[IX], indexing="ij")
    zc, xc = np.meshgrid(gridmp[IZ], gridmp[IX], indexing="ij")
    Ts = temp(zs, xs)
    f_etas, f_etan = eta(Ts), eta(temp(zc, xc))
    f_rho = rho0 / (alpha * (Ts - Tref) + 1)
    return nx, L, grid, gridmp, f_etas, f_etan, f_rho


def convection_device(ncell=4096, per_side=4, seed=11, Ra=1e6, Lbox=1e6, device="cuda", rank=0, world=1):
    """`convection` generated directly in HBM with torch (the 4096^2 case has 2.7e8 markers: ~32 GB
    of host arrays otherwise).  Same lattice/ordering/physics; the jitter comes from torch's RNG,
    so positions differ from the NumPy generator's (benchmarks only -- parity tests use `convection`).
    With world > 1 each rank generates the markers of its share of the cell rows (marker-parallel
    ranks: the cell-major order makes every share a contiguous z-slab of the cloud) -- the SAME markers the
    one-rank call generates: every rank draws the whole cloud's jitter from the same seed and keeps its rows, so
    that N ranks hold one problem, not another realisation of it (the Stokes iteration counts of the benchmark
    depend on the realisation: profiles/r02_SUMMARY.md section 8).
    Returns (nx, L, tr_x (M,2) cuda, cols list of 13 (M,) cuda tensors, opts)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    nx, L = [ncell + 1, ncell + 1], [Lbox, Lbox]
    ns = ncell * per_side
    dev = torch.device(device)
    r0, r1 = (rank * ncell) // world, ((rank + 1) * ncell) // world
    nrow = r1 - r0
    ci = torch.arange(ncell, device=dev, dtype=torch.float64)
    cz = torch.arange(r0, r1, device=dev, dtype=torch.float64)
    si = torch.arange(per_side, device=dev, dtype=torch.float64)
    # cell-major ordering (cell i, cell j, sub i, sub j)
    z = ((cz.view(-1, 1, 1, 1) * per_side + si.view(1, 1, -1, 1) + 0.5) / ns).expand(nrow, ncell, per_side, per_side)
    x = ((ci.view(1, -1, 1, 1) * per_side + si.view(1, 1, 1, -1) + 0.5) / ns).expand(nrow, ncell, per_side, per_side)
    M = nrow * ncell * per_side * per_side
    tr_x = torch.empty((M, 2), dtype=torch.float64, device=dev)
    tr_x[:, 0] = z.reshape(-1)
    tr_x[:, 1] = x.reshape(-1)
    if world == 1:
        jitter = torch.rand((M, 2), generator=g, dtype=torch.float64, device=dev)
    else:
        per_row = ncell * per_side * per_side
        whole = torch.rand((ncell * per_row, 2), generator=g, dtype=torch.float64, device=dev)
        jitter = whole[r0 * per_row:r1 * per_row].clone()
        del whole
    tr_x += (jitter - 0.5) * (0.5 / ns)
    del jitter
    rho0, alpha, k, cp, Ea = 3300.0, 3.5e-5, 4.0, 1250.0, 120e3
    dT = 1623.0 - 273.0
    eta0 = rho0 * G[IZ] * alpha * dT * Lbox ** 3 / ((k / (rho0 * cp)) * Ra)
    zn, xn = tr_x[:, 0], tr_x[:, 1]
    T = 273.0 + dT * zn + 0.01 * dT * torch.sin(np.pi * zn) * torch.cos(np.pi * xn)
    tr_x *= Lbox
    const = lambda v: torch.full((M,), float(v), dtype=torch.float64, device=dev)
    cols = [None] * NFTRAC
    cols[TR_TMP] = T
    cols[TR_RH0], cols[TR_ALP], cols[TR_HCD], cols[TR_HCP] = const(rho0), const(alpha), const(k), const(cp)
    cols[TR_ACE], cols[TR_ET0], cols[TR_MAT], cols[TR_IHT] = const(Ea), const(eta0), const(1), const(0)
    cols[TR_RHO], cols[TR_ETA] = torch.empty_like(T), torch.empty_like(T)
    # TR_MRK / TR__ID are never read on the hot path: share one zero column
    cols[TR_MRK] = cols[TR__ID] = cols[TR_IHT]
    opts = dict(do_heatdiff=True, tdep_rho=True, tdep_eta=True, bcstokes=[BC_FREESLIP] * 4)
    return nx, L, tr_x, cols, opts
