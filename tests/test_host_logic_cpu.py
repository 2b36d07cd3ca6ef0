"""CPU tests of the host-side logic of the drop-in modules (no GPU, no library calls):
ghost-axis extension, option/constant tables, setups, the spsolve dispatcher."""
import numpy as np
import pytest

from oracle import pylamp_oracle as O


def test_extended_axis_matches_oracle():
    from pylamp_b200 import pylamp_trac as T
    ax = np.linspace(0.0, 1.0, 11)
    mid = (ax[1:] + ax[:-1]) / 2
    mid = np.append(mid, mid[-1] + (mid[-1] - mid[-2]))
    for axis, lo, hi in ((ax, 0.01, 0.99), (mid, 0.001, 0.999), (mid, -0.2, 1.3), (ax, 0.0, 1.0)):
        a, l, r = T._extended_axis(axis, lo, hi)
        b, l2, r2 = O._extended_axis(axis, lo, hi)
        assert np.array_equal(a, b) and (l, r) == (l2, r2)
        assert a[0] <= lo and a[-1] >= hi


def test_constants_and_flags_match_oracle():
    from pylamp_b200 import pylamp_const as Cn, pylamp_trac as T, pylamp_stokes as S, pylamp_diff as D
    for name in ("DIM", "IZ", "IX", "IP", "G", "SECINYR", "GASR", "NFTRAC", "EPS", "TR_RHO", "TR_ETA", "TR_MRK",
                 "TR_TMP", "TR_HCD", "TR_HCP", "TR_RH0", "TR_ALP", "TR_MAT", "TR_ACE", "TR_ET0", "TR_IHT", "TR__ID"):
        assert getattr(Cn, name) == getattr(O, name), name
    for name in ("INTERP_AVG_ARITHMETIC", "INTERP_AVG_GEOMETRIC", "INTERP_AVG_WEIGHTED", "INTERP_AVG_ARITHW",
                 "INTERP_AVG_GEOMW", "INTERP_METHOD_ELEM", "INTERP_METHOD_NEAREST", "INTERP_METHOD_LINEAR",
                 "INTERP_METHOD_VELDIV"):
        assert getattr(T, name) == getattr(O, name), name
    assert (S.BC_TYPE_NOSLIP, S.BC_TYPE_FREESLIP, S.BC_TYPE_CYCLIC, S.BC_TYPE_FLOWTHRU) == (0, 1, 2, 4)
    assert (D.BC_TYPE_FIXTEMP, D.BC_TYPE_FIXFLOW) == (0, 1)
    x = np.arange(3 * 4 * 5, dtype=float)
    (vz, vx), p = S.x2vp(x, [4, 5])
    (oz, ox), op = O.x2vp(x, [4, 5])
    assert np.array_equal(vz, oz) and np.array_equal(vx, ox) and np.array_equal(p, op)
    assert np.array_equal(D.x2t(np.arange(20.0), [4, 5]), O.x2t(np.arange(20.0), [4, 5]))


def test_driver_options_mirror_the_reference_defaults():
    from pylamp_b200 import driver
    a, b = driver.Options(), O.Options()
    for k in ("do_stokes", "do_advect", "do_heatdiff", "do_subgrid_heatdiff", "tstep_adv_max", "tstep_adv_min",
              "tstep_dif_max", "tstep_dif_min", "tstep_modifier", "tdep_rho", "tdep_eta", "etamin", "etamax", "Tref",
              "tracs_fence_enabled", "bcstokes", "bcheat", "bcheatvals"):
        assert getattr(a, k) == getattr(b, k), k
    with pytest.raises(AttributeError):
        driver.Options(no_such_option=1)


def test_setups_are_deterministic_and_consistent():
    from pylamp_b200 import setups
    nx, L, x1, f1, o1 = setups.convection(ncell=8)
    _, _, x2, f2, _ = setups.convection(ncell=8)
    assert np.array_equal(x1, x2) and np.array_equal(f1, f2)
    assert x1.shape == (8 * 8 * 16, 2) and nx == [9, 9]
    # cell-major order: the cloud starts cell-sorted
    kelem, count = O.cell_index_count(x1, nx, L)
    assert np.all(np.diff(kelem) >= 0) and np.all(count == 16)
    nx, L, x, cols, o = setups.convection_device(ncell=8, device="cpu")
    assert x.shape == (8 * 8 * 16, 2) and len(cols) == 13
    k2, c2 = O.cell_index_count(x.numpy(), nx, L)
    assert np.all(np.diff(k2) >= 0) and np.all(c2 == 16)
    # a rank's share is a z-slab of the cloud
    _, _, xs, _, _ = setups.convection_device(ncell=8, device="cpu", rank=1, world=2)
    assert xs.shape[0] == x.shape[0] // 2 and float(xs[:, 0].min()) >= 0.5 * L[0] - 1e-9 * L[0]
    nx, L, tr_x, tr_f, opts = setups.c1_shipped(1234)
    assert nx == [201, 41] and tr_x.shape[0] == 201 * 41 * 45 and opts["do_heatdiff"] is False


def test_spsolve_dispatch_rejects_foreign_matrices():
    import scipy.sparse
    from pylamp_b200 import solve
    with pytest.raises(TypeError):
        solve.spsolve(scipy.sparse.identity(4, format="csc"), np.ones(4))
    class Fake:
        def solve(self, rhs, **kw):
            return rhs * 2
    assert np.array_equal(solve.spsolve(solve.csc_matrix(Fake()), np.ones(3)), 2 * np.ones(3))


def test_inject_markers_logic_on_cpu_tensors():
    """markers.inject_markers is pure tensor logic: run it on CPU tensors against the oracle's
    restatement of pylamp2.py:594-633 (counts, cells, ids and cell-mean properties must agree)."""
    import types
    import torch
    from pylamp_b200 import markers
    rng = np.random.default_rng(3)
    nx, L = [9, 7], [1.0, 0.5]
    M = 600
    tr_x = rng.random((M, 2)) * L
    tr_x[:, 0] = tr_x[:, 0] ** 2            # uneven density: some cells under-populated, some empty
    tr_f = rng.random((M, O.NFTRAC))
    tr_f[:, O.TR__ID] = np.arange(M)
    so = O.State(nx, L, tr_x.copy(), tr_f.copy())
    so.kelem, so.count = O.cell_index_count(so.tr_x, nx, L)
    s = types.SimpleNamespace(nx=nx, L=L, grid=so.grid, tr_x=torch.as_tensor(tr_x.copy()),
                              cols=[torch.as_tensor(np.ascontiguousarray(tr_f[:, k])) for k in range(O.NFTRAC)],
                              kelem=torch.as_tensor(so.kelem), count=torch.as_tensor(so.count))
    np.random.seed(0)
    n_ref = O.inject_markers(so, 12, 6)
    n_new = markers.inject_markers(s, 12, 6)
    assert n_new == n_ref > 0
    new_f = np.stack([c.numpy()[M:] for c in s.cols], axis=1)
    assert np.allclose(new_f, so.tr_f[M:], rtol=1e-12, atol=0, equal_nan=True)
    k_ref, c_ref = O.cell_index_count(so.tr_x, nx, L)
    k_new, c_new = O.cell_index_count(s.tr_x.numpy(), nx, L)
    assert np.array_equal(c_new, c_ref) and np.array_equal(k_new[M:], k_ref[M:])


def test_delete_outside_logic_on_cpu_tensors():
    """markers.delete_outside (fence disabled: markers beyond a wall are removed, pylamp2.py:563-581) is
    tensor logic: on CPU tensors it must keep exactly the markers the reference's block keeps (golden
    vectors of that block), whatever their order, with aliased columns still aliased."""
    import os
    import types
    import torch
    from conftest import GOLDEN
    from pylamp_b200 import markers
    g = np.load(os.path.join(GOLDEN, "fence_delete.npz"))
    x0, f0, v0 = g["fd_tr_x"], g["fd_tr_f"], g["fd_vel"]
    cols = [torch.as_tensor(np.ascontiguousarray(f0[:, k])) for k in range(O.NFTRAC)]
    cols[O.TR_MRK] = cols[O.TR_IHT]                                   # aliased pair
    s = types.SimpleNamespace(L=list(g["fd_L"]), tr_x=torch.as_tensor(x0.copy()), cols=cols,
                              trac_vel=torch.as_tensor(v0.copy()))
    n = markers.delete_outside(s)
    want_x, want_f, want_v = g["fd_off_tr_x"], g["fd_off_tr_f"].copy(), g["fd_off_vel"]
    want_f[:, O.TR_MRK] = want_f[:, O.TR_IHT]
    assert n == x0.shape[0] - want_x.shape[0] and s.tr_x.shape[0] == want_x.shape[0]
    got_f = np.stack([c.numpy() for c in s.cols], axis=1)
    order = np.argsort(got_f[:, O.TR__ID])
    assert np.array_equal(got_f[order], want_f) and np.array_equal(s.tr_x.numpy()[order], want_x)
    assert np.array_equal(s.trac_vel.numpy()[order], want_v)
    assert s.cols[O.TR_MRK].data_ptr() == s.cols[O.TR_IHT].data_ptr()
    assert markers.delete_outside(s) == 0                             # idempotent


def test_benchmark_cloud_is_one_realisation_for_any_rank_count():
    """setups.convection_device: the slabs N ranks generate are the rows of the cloud one rank generates (same jitter
    stream), so `bench.py --gpus N` times ONE problem -- the Stokes iteration counts depend on the realisation."""
    import torch
    from pylamp_b200 import setups
    from pylamp_b200.pylamp_const import TR_TMP
    nx, L, x1, c1, _ = setups.convection_device(ncell=24, per_side=4, device="cpu")
    for world in (2, 3, 4):
        parts = [setups.convection_device(ncell=24, per_side=4, device="cpu", rank=r, world=world) for r in range(world)]
        assert torch.equal(torch.cat([p[2] for p in parts]), x1)
        assert torch.equal(torch.cat([p[3][TR_TMP] for p in parts]), c1[TR_TMP])
    _, _, x2, _, _ = setups.convection_device(ncell=24, per_side=4, seed=12, device="cpu")
    assert not torch.equal(x1, x2)

