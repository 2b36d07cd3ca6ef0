"""GPU parity: marker kernels (trac2grid / grid2trac / RK / fence / cell index+count / property
update / centre velocities) through the C-ABI versus the oracle and the committed golden vectors
of the unmodified reference.

Tolerances: cell indices and per-cell counts BIT-EXACT; grid2trac / RK positions 1e-10 relative
(north_star) -- in practice ~1e-15; trac2grid sums differ from np.add.at only by fp64 summation
order (atomics), asserted to 1e-12 relative.
"""
import numpy as np
import pytest
import torch

from oracle import pylamp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    from pylamp_b200 import pylamp_trac
    return pylamp_trac


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def test_trac2grid_golden(T, golden_kernels):
    g = golden_kernels
    nx, L = list(g["nx"]), list(g["L"])
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    gf = [np.zeros(nx) for _ in range(6)]
    T.trac2grid(g["tr_x"], g["tr_vals"], mesh, grid, gf, nx, avgscheme=list(g["t2g_nodes_scheme"]))
    ref = g["t2g_nodes"]
    assert np.array_equal(np.isnan(np.array(gf)), np.isnan(ref))
    assert np.allclose(np.array(gf), ref, rtol=1e-12, atol=0, equal_nan=True)
    targets = {"cc": ([gridmp[0], gridmp[1]], meshmp),
               "zmid": ([gridmp[0], grid[1]], [meshmp[0], mesh[1]]),
               "xmid": ([grid[0], gridmp[1]], [mesh[0], meshmp[1]])}
    for name, (gr, m) in targets.items():
        gf = [np.zeros(nx) for _ in range(2)]
        T.trac2grid(g["tr_x"], g["tr_vals"][:, :2], m, gr, gf, nx, avgscheme=[6, 2])
        assert np.allclose(np.array(gf), g["t2g_" + name], rtol=1e-12, atol=0, equal_nan=True), name


def test_grid2trac_rk_golden(T, golden_kernels, capsys):
    g = golden_kernels
    nx, L = list(g["nx"]), list(g["L"])
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    F = list(g["g2t_fields"])
    M = g["tr_x"].shape[0]
    for m, name in ((16, "linear"), (8, "nearest")):
        a = np.zeros((M, 2))
        T.grid2trac(g["tr_x"], a, grid, F, nx, method=m)
        assert np.allclose(a, g["g2t_" + name], rtol=1e-13, atol=1e-300)
        a = np.zeros((M, 2))
        T.grid2trac(g["g2t_x_outside"], a, grid, F, nx, method=m, defval=-7.5)
        assert np.allclose(a, g["g2t_" + name + "_outside"], rtol=1e-13, atol=1e-300)
    with pytest.raises(Exception, match="stopOnError"):
        T.grid2trac(g["g2t_x_outside"], np.zeros((M, 2)), grid, F, nx, stopOnError=True)
    with pytest.raises(Exception, match="VELDIV"):
        T.grid2trac(g["tr_x"], np.zeros((M, 1)), grid, F[:1], nx, method=32)
    a = np.zeros((M, 2))
    T.grid2trac(g["tr_x"], a, [g["rk_grid_z"], g["rk_grid_x"]], list(g["rk_vels"]),
                [nx[0] + 1, nx[1] + 1], defval=0, method=32)
    assert np.allclose(a, g["g2t_veldiv"], rtol=1e-12, atol=1e-300)
    v, x = T.RK(g["tr_x"], [g["rk_grid_z"], g["rk_grid_x"]], list(g["rk_vels"]), nx, float(g["rk_tstep"]))
    assert _rel(x, g["rk_x"]) < 1e-10 and np.allclose(x, g["rk_x"], rtol=1e-10, atol=0)
    assert _rel(v, g["rk_vel"]) < 1e-9
    with pytest.raises(Exception, match="don't know"):
        T.RK(g["tr_x"], None, None, nx, 1.0, order=3)


def test_strided_view_written_in_place(T, golden_kernels):
    """pylamp2.py:445 passes tr_f[IPROC::NPROC] style views: the result must land in the view."""
    g = golden_kernels
    nx, L = list(g["nx"]), list(g["L"])
    grid = O.make_grids(nx, L)[0]
    M = g["tr_x"].shape[0]
    big = np.full((M, 3), -1.0)
    T.grid2trac(g["tr_x"][::2], big[::2, 1:3], grid, list(g["g2t_fields"]), nx, method=16)
    assert np.allclose(big[::2, 1:3], g["g2t_linear"][::2], rtol=1e-13)
    assert np.all(big[1::2] == -1.0) and np.all(big[:, 0] == -1.0)


@pytest.mark.parametrize("nxL", [([17, 9], [2.0, 0.5]), ([201, 41], [1.0, 0.2]), ([1025, 1025], [1e6, 1e6])])
def test_cell_index_count_bit_exact(nxL):
    from pylamp_b200 import markers
    nx, L = nxL
    rng = np.random.default_rng(5)
    x = rng.random((300000, 2)) * L
    # adversarial positions: exactly on nodes and one ulp either side
    nodes = np.linspace(0, L[0], nx[0])[1:-1]
    k = min(len(nodes), 1000)
    x[:k, 0] = nodes[:k]
    x[k:2 * k, 0] = np.nextafter(nodes[:k], 0)
    x[2 * k:3 * k, 0] = np.nextafter(nodes[:k], np.inf)
    kelem, count = O.cell_index_count(x, nx, L)
    kd, cd = markers.cell_index_count(torch.as_tensor(x).cuda(), nx, L)
    assert np.array_equal(kd.cpu().numpy(), kelem)
    assert np.array_equal(cd.cpu().numpy(), count)


def test_fence_properties_centre(golden_kernels):
    from pylamp_b200 import markers
    g = golden_kernels
    rng = np.random.default_rng(2)
    L = [1.0, 0.2]
    x = (rng.random((50000, 2)) * 1.2 - 0.1) * L
    x[:10] = 0.0
    x[10:20] = L
    ref = x.copy()
    O.fence(ref, np.zeros((x.shape[0], O.NFTRAC)), L, [1, 1, 1, 1])
    xd = torch.as_tensor(x).cuda()
    markers.fence(xd, L)
    assert np.array_equal(xd.cpu().numpy(), ref)
    # property update, pylamp2.py:291-303
    M = 20000
    tr_f = np.zeros((M, O.NFTRAC))
    tr_f[:, O.TR_TMP] = rng.uniform(273, 1900, M)
    tr_f[:, O.TR_RH0] = rng.uniform(2500, 3400, M)
    tr_f[:, O.TR_ALP] = 3.5e-5
    tr_f[:, O.TR_ACE] = rng.uniform(0, 300e3, M)
    tr_f[:, O.TR_ET0] = 10 ** rng.uniform(18, 22, M)
    for tr, te in ((True, True), (False, False)):
        ref = tr_f.copy()
        O.update_properties(ref, tr, te, 1623, 1e17, 1e23)
        cols = {k: torch.as_tensor(np.ascontiguousarray(tr_f[:, k])).cuda()
                for k in (O.TR_TMP, O.TR_RH0, O.TR_ALP, O.TR_ACE, O.TR_ET0)}
        rho, eta = markers.update_properties(cols[O.TR_TMP], cols[O.TR_RH0], cols[O.TR_ALP],
                                             cols[O.TR_ACE], cols[O.TR_ET0], tr, te, 1623, 1e17, 1e23)
        assert np.allclose(rho.cpu().numpy(), ref[:, O.TR_RHO], rtol=1e-14)
        assert np.allclose(eta.cpu().numpy(), ref[:, O.TR_ETA], rtol=1e-12)
    # centre velocities + ring
    nx = list(g["nx"])
    gridmp = O.make_grids(nx, list(g["L"]))[2]
    for bc in ([1, 1, 1, 1], [0, 1, 0, 1], [1, 1, 0, 1]):
        ng, vels = O.centre_velocities(list(g["rk_newvel"]), gridmp, nx, bc)
        vz, vx = markers.centre_velocities(torch.as_tensor(g["rk_newvel"][0]).cuda(),
                                           torch.as_tensor(g["rk_newvel"][1]).cuda(), bc)
        assert np.array_equal(vz.cpu().numpy(), vels[0]) and np.array_equal(vx.cpu().numpy(), vels[1])


def test_large_random_cloud_vs_oracle(T):
    """1025^2-node grid, 4 M markers: trac2grid (all four staggered targets) + RK vs the oracle."""
    rng = np.random.default_rng(3)
    nx, L = [513, 385], [1.0, 0.75]
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    M = 2_000_000
    x = rng.random((M, 2)) * L
    f = np.stack([rng.uniform(1, 2, M), 10 ** rng.uniform(18, 24, M)], axis=1)
    for gr, sch in (([grid[0], grid[1]], [5, 6]), ([gridmp[0], gridmp[1]], [1, 2]),
                    ([gridmp[0], grid[1]], [5, 6]), ([grid[0], gridmp[1]], [5, 6])):
        ref = [np.zeros(nx), np.zeros(nx)]
        O.trac2grid(x, f, None, gr, ref, nx, avgscheme=sch)
        out = [np.zeros(nx), np.zeros(nx)]
        T.trac2grid(x, f, None, gr, out, nx, avgscheme=sch)
        for a, b in zip(out, ref):
            assert np.array_equal(np.isnan(a), np.isnan(b))
            assert np.allclose(a, b, rtol=1e-11, atol=0, equal_nan=True)
    # divergence-free-ish velocity field, RK
    zc, xc = np.meshgrid(np.arange(nx[0]) / (nx[0] - 1), np.arange(nx[1]) / (nx[1] - 1), indexing="ij")
    newvel = [np.sin(np.pi * zc) * np.cos(np.pi * xc), -np.cos(np.pi * zc) * np.sin(np.pi * xc)]
    ng, vels = O.centre_velocities(newvel, gridmp, nx, [1, 1, 1, 1])
    dt = 0.67 * (L[0] / (nx[0] - 1))
    vr, xr = O.RK(x, ng, vels, nx, dt)
    vg, xg = T.RK(x, ng, vels, nx, dt)
    assert np.allclose(xg, xr, rtol=1e-10, atol=0)
    assert _rel(vg, vr) < 1e-9


@pytest.mark.parametrize("cloud", ["sorted", "random", "faces"])
def test_rk4_cell_lookup_on_faces_and_fused_fence_count(T, cloud):
    """RK4 against the oracle's RK -- cell-ordered and unordered clouds, markers on and within a few ulps of cell
    faces of the centre grid (the Meyer-Jenny term is discontinuous there: the reference's exact multiply-then-divide
    cell lookup decides), a time step twice the CFL one (some stage positions leave the grid) -- and
    plb_rk4_fence_count against the oracle's RK + fence + cell_index_count, bit for bit in the indices."""
    from pylamp_b200 import _lib, setups
    ctx = _lib.default_context()
    rng = np.random.default_rng(31)
    ncz, ncx, L = 40, 56, [1.0, 1.4]
    nx = [ncz + 1, ncx + 1]
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    if cloud == "random":
        x = rng.random((50000, 2)) * L
    else:
        x = setups.lattice_markers(ncz, ncx, L, 4, seed=8)[0]
        if cloud == "faces":
            m = x.shape[0] // 3
            x[:m, 0] = np.round(x[:m, 0] * ncz / L[0] - 0.5) * L[0] / ncz + 0.5 * L[0] / ncz    # on centre-grid faces
            x[:m:2, 0] = np.nextafter(x[:m:2, 0], 0)
            x[m:2 * m, 1] = np.nextafter(np.round(x[m:2 * m, 1] * ncx / L[1]) * L[1] / ncx, 10)
            x = np.clip(x, 1e-6, np.array(L) - 1e-6)
    z, xx = np.meshgrid(grid[0], grid[1], indexing="ij")
    newvel = [np.sin(np.pi * z / L[0]) * np.cos(np.pi * xx / L[1]) + 0.05 * rng.normal(size=nx),
              -np.cos(np.pi * z / L[0]) * np.sin(np.pi * xx / L[1]) + 0.05 * rng.normal(size=nx)]
    ng, vels = O.centre_velocities(newvel, gridmp, nx, [1, 1, 1, 1])
    for cfl in (0.67, 1.5):
        dt = cfl * (L[0] / ncz) / max(np.abs(newvel[0]).max(), np.abs(newvel[1]).max())
        vr, xr = O.RK(x, ng, vels, nx, dt)
        vg, xg = T.RK(x, ng, vels, nx, dt)
        assert np.allclose(xg, xr, rtol=1e-12, atol=1e-15), (cloud, cfl, np.abs(xg - xr).max())
        assert np.allclose(vg, vr, rtol=1e-9, atol=1e-12 * np.abs(vr).max())
        # fused RK4 + fence + count
        xd = torch.as_tensor(x).cuda()
        v, xn, kelem, count = T.rk4_fence_count_device(ctx, xd, ng, torch.as_tensor(vels[0]).cuda(), torch.as_tensor(vels[1]).cuda(),
                                                       [nx[0] + 1, nx[1] + 1], dt, nx, L, 2.0 ** -10)
        xf = xr.copy()
        O.fence(xf, np.zeros((xf.shape[0], O.NFTRAC)), L, [1, 1, 1, 1])
        kr, cr = O.cell_index_count(xf, nx, L)
        assert np.allclose(xn.cpu().numpy(), xf, rtol=1e-12, atol=1e-15)
        assert np.allclose(v.cpu().numpy(), vr, rtol=1e-9, atol=1e-12 * np.abs(vr).max())
        same = kelem.cpu().numpy() == kr
        assert same.mean() > 0.9999          # (a position that differs in the last bit may sit on the other side of a face)
        if same.all():
            assert np.array_equal(count.cpu().numpy(), cr)
        kg, cg = O.cell_index_count(xn.cpu().numpy(), nx, L)       # ... but the indices of the positions returned are exact
        assert np.array_equal(kelem.cpu().numpy(), kg) and np.array_equal(count.cpu().numpy(), cg)


def test_sort_by_cell_keeps_results(T):
    """A shuffled cloud and its cell-sorted version give the same grids (order-independent sums)."""
    from pylamp_b200 import markers
    rng = np.random.default_rng(9)
    nx, L = [65, 49], [1.0, 0.75]
    grid = O.make_grids(nx, L)[0]
    M = 200000
    x = rng.random((M, 2)) * L
    f = rng.uniform(1, 2, M)
    xd, fd = torch.as_tensor(x).cuda(), torch.as_tensor(f).cuda()
    xs, (fs,), _ = markers.sort_by_cell(xd, [fd], nx, L)
    k, _ = markers.cell_index_count(xs, nx, L)
    assert bool((k[1:] >= k[:-1]).all())
    out_a, out_b = [np.zeros(nx)], [np.zeros(nx)]
    T.trac2grid(x, f[:, None], None, grid, out_a, nx, avgscheme=[5])
    T.trac2grid(xs.cpu().numpy(), fs.cpu().numpy()[:, None], None, grid, out_b, nx, avgscheme=[5])
    ref = [np.zeros(nx)]
    O.trac2grid(x, f[:, None], None, grid, ref, nx, avgscheme=[5])
    assert np.allclose(out_a[0], ref[0], rtol=1e-12) and np.allclose(out_b[0], ref[0], rtol=1e-12)


def test_device_sort_is_a_cell_major_permutation(T):
    """csrc/sort.cu: the slot table is a permutation, the result is cell-major (keys as pylamp2.py:588-589,
    markers outside the box clamped into the nearest cell), the first-slot table equals the exclusive scan of
    np.bincount, every carried array (incl. (M,2) extras and aliased columns) moves with its marker."""
    from pylamp_b200 import markers
    rng = np.random.default_rng(10)
    for nx, L, M in (([65, 49], [1.0, 0.75], 300001), ([9, 1030], [0.3, 2.0], 70000), ([6, 6], [1.0, 1.0], 37)):
        x = rng.random((M, 2)) * L
        x[:7] = [-0.01, 0.5 * L[1]]              # outside (fence disabled / FLOWTHRU): clamped keys
        x[7:11] = [0.5 * L[0], 1.2 * L[1]]
        ident = np.arange(M, dtype=np.float64)
        vel = rng.normal(size=(M, 2))
        xd, idd, vd = torch.as_tensor(x).cuda(), torch.as_tensor(ident).cuda(), torch.as_tensor(vel).cuda()
        for consume in (False, True):
            xin, idin, vin = (xd.clone(), idd.clone(), vd.clone()) if consume else (xd, idd, vd)
            xs, cols, (vs,), start = markers.sort_by_cell(xin, [idin, idin], nx, L, extra=[vin], want_cell_start=True,
                                                          consume=consume)
            assert cols[0].data_ptr() == cols[1].data_ptr()          # aliased columns stay aliased
            order = cols[0].cpu().numpy().astype(np.int64)
            assert np.array_equal(np.sort(order), np.arange(M))       # a permutation
            assert np.array_equal(xs.cpu().numpy(), x[order]) and np.array_equal(vs.cpu().numpy(), vel[order])
            ie = np.clip(np.floor((nx[0] - 1) * x[:, 0] / L[0]).astype(np.int64), 0, nx[0] - 2)
            je = np.clip(np.floor((nx[1] - 1) * x[:, 1] / L[1]).astype(np.int64), 0, nx[1] - 2)
            key = ie * (nx[1] - 1) + je
            assert bool(np.all(np.diff(key[order]) >= 0))             # cell-major
            ncell = (nx[0] - 1) * (nx[1] - 1)
            want = np.concatenate([[0], np.cumsum(np.bincount(key, minlength=ncell))])
            assert np.array_equal(start.cpu().numpy().astype(np.int64), want)
        if not consume:
            assert np.array_equal(xd.cpu().numpy(), x)                # inputs untouched without `consume`


def _t2g_dev(T, x, cols, schemes, grid, nx, variant, view_offset=0):
    """trac2grid_device on CUDA tensors with the given scatter-kernel variant; `view_offset` > 0
    passes views that start `view_offset` markers into larger allocations (not 32-byte aligned)."""
    from pylamp_b200 import _lib
    ctx = _lib.default_context()
    ctx.set_param("t2g_variant", variant)
    try:
        pad = np.zeros((view_offset, 2))
        xd = torch.as_tensor(np.concatenate([pad, x])).cuda()[view_offset:]
        cd = [torch.as_tensor(np.concatenate([np.ones(view_offset), c])).cuda()[view_offset:] for c in cols]
        out = [torch.zeros(tuple(nx), dtype=torch.float64, device="cuda") for _ in cols]
        T.trac2grid_device(ctx, xd, cd, schemes, grid, out)
        return [o.cpu().numpy() for o in out]
    finally:
        ctx.set_param("t2g_variant", 1)


@pytest.mark.parametrize("cloud", ["random", "sorted", "drifted", "outside"])
@pytest.mark.parametrize("ragged", [0, 1, 2, 3])
def test_trac2grid_chunk_kernel_matches_generic_and_oracle(T, cloud, ragged):
    """The wide-load chunk kernel (t2g_variant 1: two aggregates per 4-marker chunk, one-marker path
    for the rest; t2g_variant 2: the first runs are combined across lanes as well) and the generic
    scatter kernel against the oracle, on clouds that exercise every
    path: unordered (mostly one-marker path), cell-ordered (single runs), cell-ordered then displaced
    (two runs per chunk), markers beyond the grid (ghost extension), marker counts not divisible by
    4 (tail launch), all four staggered targets, arithmetic and geometric weighted means."""
    from pylamp_b200 import setups
    rng = np.random.default_rng(11)
    ncz, ncx, L = 48, 40, [1.0, 0.75]
    nx = [ncz + 1, ncx + 1]
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    if cloud == "random":
        x = rng.random((60000, 2)) * L
    else:
        x = setups.lattice_markers(ncz, ncx, L, 4, seed=3)[0]
        if cloud == "drifted":
            x = x + np.array([0.37 * L[0] / ncz, 0.61 * L[1] / ncx])
            x = np.minimum(np.maximum(x, 1e-9), np.array(L) - 1e-9)
        if cloud == "outside":
            x = x + np.array([-0.4 * L[0] / ncz, 0.3 * L[1] / ncx])      # beyond z=0 and x=L
    if ragged:
        x = x[:x.shape[0] - 4 + ragged]
    M = x.shape[0]
    cols = [rng.uniform(1, 2, M), 10 ** rng.uniform(18, 24, M), rng.uniform(-1, 1, M)]
    schemes = [5, 6, 5]
    f = np.stack(cols, axis=1)
    for gr in ([grid[0], grid[1]], [gridmp[0], gridmp[1]], [gridmp[0], grid[1]], [grid[0], gridmp[1]]):
        ref = [np.zeros(nx) for _ in cols]
        O.trac2grid(x, f, None, gr, ref, nx, avgscheme=schemes)
        for variant in (1, 2, 0):          # chunk kernel, chunk kernel with merged first runs, generic kernel
            for a, r in zip(_t2g_dev(T, x, cols, schemes, gr, nx, variant), ref):
                assert np.array_equal(np.isnan(a), np.isnan(r)), variant
                # the arithmetic mean of values in (-1,1) can cancel: absolute floor of a few ulps of 1
                assert np.allclose(a, r, rtol=1e-12, atol=1e-14, equal_nan=True), variant


def _fused_case(T, x, nx, L, node_k, centre_k, view_offset=0, poison=False, seed=13):
    """plb_trac2grid_fused (all targets of a step in one pass) against the oracle's trac2grid per target."""
    from pylamp_b200 import _lib
    ctx = _lib.default_context()
    rng = np.random.default_rng(seed)
    grid, mesh, gridmp, meshmp = O.make_grids(nx, L)
    M = x.shape[0]
    ncols = node_k + 2
    cols = [rng.uniform(-1, 1, M) if c % 3 == 2 else 10 ** rng.uniform(18, 24, M) if c % 3 == 1 else rng.uniform(1, 2, M)
            for c in range(ncols)]
    if poison:
        cols[1][[3, M // 2, M - 1]] = 0.0            # log -> -inf: nodes zeroed before exp (pylamp_trac.py:301)
        cols[0][[7, M // 3]] = np.nan                 # NaN property (marker injected into an empty cell)
    sch = lambda c: 6 if c % 3 == 1 else 5
    pad = lambda a: torch.as_tensor(np.concatenate([np.ones((view_offset,) + a.shape[1:]), a])).cuda()[view_offset:]
    xd, cd = pad(x), [pad(c) for c in cols]
    new = lambda: torch.full(tuple(nx), -7.0, dtype=torch.float64, device="cuda")
    node_ids = list(range(node_k))
    centre_ids = [1, 0][:centre_k]                     # the geometric column first (shared with the node target)
    targets = [(0, node_ids), (1, centre_ids), (2, [ncols - 1]), (3, [ncols - 1])]
    grids = {0: [grid[0], grid[1]], 1: [gridmp[0], gridmp[1]], 2: [gridmp[0], grid[1]], 3: [grid[0], gridmp[1]]}
    outs = {kind: [new() for _ in ids] for kind, ids in targets}
    mm = T.marker_minmax(xd, ctx)
    ok = T.trac2grid_fused_device(ctx, xd, [(kind, [cd[c] for c in ids], [sch(c) for c in ids], outs[kind])
                                            for kind, ids in targets], grid, gridmp, mm)
    inside = x[:, 0].min() >= 0 and x[:, 0].max() <= L[0] and x[:, 1].min() >= 0 and x[:, 1].max() <= L[1]
    assert ok == bool(inside)
    if not ok:
        assert all(bool((o == -7.0).all()) for os_ in outs.values() for o in os_)      # nothing written
        return
    for kind, ids in targets:
        ref = [np.zeros(nx) for _ in ids]
        with np.errstate(divide="ignore", invalid="ignore"):
            O.trac2grid(x, np.stack([cols[c] for c in ids], axis=1), None, grids[kind], ref, nx,
                        avgscheme=[sch(c) for c in ids])
        for o, r in zip(outs[kind], ref):
            a = o.cpu().numpy()
            assert np.array_equal(np.isnan(a), np.isnan(r)), (kind, int(np.isnan(a).sum()), int(np.isnan(r).sum()))
            assert np.allclose(a, r, rtol=1e-12, atol=1e-14, equal_nan=True), kind


@pytest.mark.parametrize("cloud", ["random", "sorted", "drifted", "outside", "on_nodes"])
@pytest.mark.parametrize("ragged", [0, 3])
def test_trac2grid_fused_matches_oracle(T, cloud, ragged):
    """The fused step kernel on clouds that exercise every path: unordered (runs of one marker), cell-ordered
    (one run per cell), cell-ordered then displaced, markers exactly on nodes / cell faces / midpoints (weights
    0 and 1, half-cell ties), markers beyond the grid (request refused -> caller falls back), chunks that are
    not full (plain loads instead of bulk copies)."""
    from pylamp_b200 import setups
    rng = np.random.default_rng(11)
    ncz, ncx, L = 48, 40, [1.0, 0.75]
    nx = [ncz + 1, ncx + 1]
    if cloud == "random":
        x = rng.random((60000, 2)) * L
    elif cloud == "on_nodes":
        g = O.make_grids(nx, L)
        # (not the last node: floor((n-1)(x-Lmin)/L) = n-1 there, which the reference itself cannot index)
        zz, xx = np.meshgrid(np.concatenate([g[0][0][:-1], g[2][0][:-1]]), np.concatenate([g[0][1][:-1], g[2][1][:-1]]), indexing="ij")
        x = np.stack([zz.ravel(), xx.ravel()], axis=1)
        x = np.concatenate([x, rng.random((5000, 2)) * L])
    else:
        x = setups.lattice_markers(ncz, ncx, L, 4, seed=3)[0]
        if cloud == "drifted":
            x = x + np.array([0.37 * L[0] / ncz, 0.61 * L[1] / ncx])
            x = np.minimum(np.maximum(x, 1e-9), np.array(L) - 1e-9)
        if cloud == "outside":
            x = x + np.array([-0.4 * L[0] / ncz, 0.3 * L[1] / ncx])      # beyond z=0 and x=L
    if ragged:
        x = x[:x.shape[0] - 4 + ragged]
    _fused_case(T, x, nx, L, node_k=6, centre_k=1)


def test_trac2grid_fused_shapes_alignment_and_nonfinite_values(T):
    from pylamp_b200 import setups
    ncz, ncx, L = 24, 36, [0.5, 1.0]
    nx = [ncz + 1, ncx + 1]
    x = setups.lattice_markers(ncz, ncx, L, 4, seed=4)[0]
    for node_k, centre_k in ((1, 1), (2, 2), (4, 1), (6, 2)):             # 1..2 node tasks, 1..2 centre tasks
        _fused_case(T, x, nx, L, node_k, centre_k)
    from pylamp_b200 import _lib
    ctx = _lib.default_context()
    for parts, nm, nfmax in ((1, 1024, 3), (2, 960, 6), (4, 1024, 2), (2, 1024, 6)):   # tuning knobs: same results
        ctx.set_param("t2g_parts", parts), ctx.set_param("t2g_nm", nm), ctx.set_param("t2g_nfmax", nfmax)
        try:
            _fused_case(T, x, nx, L, 6, 1)
            _fused_case(T, x[5:3000], nx, L, 1, 1, poison=True)
        finally:
            ctx.set_param("t2g_parts", 0), ctx.set_param("t2g_nm", 0), ctx.set_param("t2g_nfmax", 0)
    _fused_case(T, x, nx, L, 6, 1, view_offset=1)                          # 8-byte aligned views: no bulk copies
    _fused_case(T, x[:1], nx, L, 3, 1)                                     # a single marker
    _fused_case(T, x[:1024 * 3], nx, L, 6, 1)                              # whole chunks only
    _fused_case(T, x, nx, L, 6, 1, poison=True)                            # log(0) = -inf and NaN properties
    rng = np.random.default_rng(5)
    _fused_case(T, rng.random((20000, 2)) * L, nx, L, 6, 1, poison=True)   # the same, unordered


def test_trac2grid_chunk_kernel_falls_back_on_unaligned_views_and_counts(T):
    """Views that do not start on a 32-byte boundary and unweighted schemes (per-node marker
    counts) must take the generic kernel whatever the variant: same results either way."""
    rng = np.random.default_rng(12)
    nx, L = [33, 25], [1.0, 0.75]
    grid = O.make_grids(nx, L)[0]
    M = 30001
    x = rng.random((M, 2)) * L
    cols = [rng.uniform(1, 2, M), 10 ** rng.uniform(18, 24, M)]
    f = np.stack(cols, axis=1)
    for schemes, off in (([5, 6], 1), ([5, 6], 3), ([1, 2], 0), ([5, 2], 0)):
        ref = [np.zeros(nx) for _ in cols]
        O.trac2grid(x, f, None, grid, ref, nx, avgscheme=schemes)
        for variant in (0, 1, 2):
            out = _t2g_dev(T, x, cols, schemes, grid, nx, variant, view_offset=off)
            for a, r in zip(out, ref):
                assert np.allclose(a, r, rtol=1e-12, atol=0, equal_nan=True)


def test_trac2grid_chunk_kernel_log_of_zero(T):
    """A marker with eta = 0 gives ln = -inf: its nodes' sums are zeroed before exp (pylamp_trac.py:301)
    and no other node is affected (the run masks select values, they never multiply by them)."""
    from pylamp_b200 import setups
    ncz, ncx, L = 16, 12, [1.0, 0.75]
    nx = [ncz + 1, ncx + 1]
    grid = O.make_grids(nx, L)[0]
    x = setups.lattice_markers(ncz, ncx, L, 4, seed=5)[0]
    M = x.shape[0]
    eta = np.full(M, 1e20)
    eta[[5, 1234, M - 2]] = 0.0
    ref = [np.zeros(nx)]
    with np.errstate(divide="ignore", invalid="ignore"):
        O.trac2grid(x, eta[:, None], None, grid, ref, nx, avgscheme=[6])
    for variant in (0, 1, 2):
        out = _t2g_dev(T, x, [eta], [6], grid, nx, variant)
        assert np.allclose(out[0], ref[0], rtol=1e-12, atol=0, equal_nan=True)


def test_fence_count_fused_bit_exact():
    """plb_fence_count == plb_fence followed by plb_cell_index_count == the oracle, bit for bit."""
    from pylamp_b200 import markers
    rng = np.random.default_rng(21)
    nx, L = [201, 41], [1.0, 0.2]
    x = (rng.random((400000, 2)) * 1.1 - 0.05) * L          # some markers beyond every wall
    x[:5] = 0.0
    x[5:10] = L
    ref = x.copy()
    O.fence(ref, np.zeros((x.shape[0], O.NFTRAC)), L, [1, 1, 1, 1])
    kelem, count = O.cell_index_count(ref, nx, L)
    xd = torch.as_tensor(x).cuda()
    kd, cd = markers.fence_count(xd, nx, L)
    assert np.array_equal(xd.cpu().numpy(), ref)
    assert np.array_equal(kd.cpu().numpy(), kelem) and np.array_equal(cd.cpu().numpy(), count)
    xd2 = torch.as_tensor(x).cuda()
    _, cd2 = markers.fence_count(xd2, nx, L, want_kelem=False)
    assert np.array_equal(cd2.cpu().numpy(), count)


def test_delete_outside_on_device_matches_reference_block():
    """markers.delete_outside on CUDA tensors keeps exactly the markers the reference's own block keeps
    (tests/golden/fence_delete.npz, fence disabled), and the driver's cell count then sees them all."""
    import os
    import types
    from conftest import GOLDEN
    from pylamp_b200 import markers
    g = np.load(os.path.join(GOLDEN, "fence_delete.npz"))
    x0, f0, v0 = g["fd_tr_x"], g["fd_tr_f"], g["fd_vel"]
    cols = [torch.as_tensor(np.ascontiguousarray(f0[:, k])).cuda() for k in range(O.NFTRAC)]
    s = types.SimpleNamespace(L=list(g["fd_L"]), tr_x=torch.as_tensor(x0.copy()).cuda(), cols=cols,
                              trac_vel=torch.as_tensor(v0.copy()).cuda())
    n = markers.delete_outside(s)
    want_x, want_f, want_v = g["fd_off_tr_x"], g["fd_off_tr_f"], g["fd_off_vel"]
    assert n == x0.shape[0] - want_x.shape[0]
    got_f = np.stack([c.cpu().numpy() for c in s.cols], axis=1)
    order = np.argsort(got_f[:, O.TR__ID])
    assert np.array_equal(got_f[order], want_f) and np.array_equal(s.tr_x.cpu().numpy()[order], want_x)
    assert np.array_equal(s.trac_vel.cpu().numpy()[order], want_v)
    nx = [9, 7]
    _, count = markers.cell_index_count(s.tr_x, nx, s.L, want_kelem=False)
    assert int(count.sum().item()) == want_x.shape[0]


@pytest.mark.parametrize("frac_outside", [0.0, 0.02, 0.6, 1.0])
def test_delete_outside_kernels_any_fraction(frac_outside):
    """plb_delete_outside (list / tail survivors / fill) against NumPy's boolean mask for clouds with none, a few,
    most (the second listing pass: more than an eighth of the cloud) and all markers beyond the box
    (pylamp2.py:563-581: x <= 0 or x >= L); shared zero column, (M,2) velocities."""
    import types
    from pylamp_b200 import markers
    rng = np.random.default_rng(int(100 * frac_outside) + 3)
    M, L = 300_001, [2.0, 3.0]
    x = rng.uniform(1e-3, 1.0, (M, 2)) * (np.array(L) - 2e-3)
    out = rng.random(M) < frac_outside
    side = rng.integers(0, 4, M)
    x[out & (side == 0), 0] = -rng.random(int((out & (side == 0)).sum()))
    x[out & (side == 1), 0] = L[0] + rng.random(int((out & (side == 1)).sum()))
    x[out & (side == 2), 1] = 0.0                                     # exactly on the wall: outside
    x[out & (side == 3), 1] = L[1]
    ident = np.arange(M, dtype=np.float64)
    temp = rng.random(M)
    vel = rng.standard_normal((M, 2))
    zero = torch.zeros(M, dtype=torch.float64, device="cuda")
    s = types.SimpleNamespace(L=L, tr_x=torch.as_tensor(x.copy()).cuda(),
                              cols=[torch.as_tensor(ident).cuda(), zero, torch.as_tensor(temp).cuda(), zero],
                              trac_vel=torch.as_tensor(vel.copy()).cuda())
    n = markers.delete_outside(s)
    keep = ~((x[:, 0] <= 0) | (x[:, 0] >= L[0]) | (x[:, 1] <= 0) | (x[:, 1] >= L[1]))
    assert n == M - int(keep.sum()) and s.tr_x.shape[0] == int(keep.sum())
    got_id = s.cols[0].cpu().numpy()
    order = np.argsort(got_id)
    assert np.array_equal(got_id[order], ident[keep])
    assert np.array_equal(s.tr_x.cpu().numpy()[order], x[keep])
    assert np.array_equal(s.cols[2].cpu().numpy()[order], temp[keep])
    assert np.array_equal(s.trac_vel.cpu().numpy()[order], vel[keep])
    assert s.cols[1].data_ptr() == s.cols[3].data_ptr() and s.cols[1].shape[0] == int(keep.sum())
    assert markers.delete_outside(s) == 0                             # idempotent

