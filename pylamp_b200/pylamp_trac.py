"""Drop-in for the reference's pylamp_trac module: trac2grid / grid2trac / RK on the GPU.

Same names, argument meaning and error behaviour as /root/reference/pylamp_trac.py
(trac2grid :161-318, grid2trac :30-158, RK :321-388).  Arrays may be NumPy arrays (host: they are
copied to the device, processed by the CUDA kernels and copied back -- the reference-facing,
"end to end" path) or torch CUDA float64 tensors (device-resident: no copies, results written in
place).  There is no CPU implementation here: without the CUDA library every call raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .pylamp_const import *  # noqa: F401,F403
from .pylamp_const import DIM, IX, IZ

# pylamp_trac.py:11-22
INTERP_AVG_ARITHMETIC = 1
INTERP_AVG_GEOMETRIC = 2
INTERP_AVG_WEIGHTED = 4
INTERP_AVG_ARITHW = 5
INTERP_AVG_GEOMW = 6
INTERP_METHOD_IDW = 1
INTERP_METHOD_GRIDDATA = 2
INTERP_METHOD_ELEM = 4
INTERP_METHOD_NEAREST = 8
INTERP_METHOD_LINEAR = 16
INTERP_METHOD_VELDIV = 32


def _ctx(t=None):
    dev = t.device.index if isinstance(t, torch.Tensor) and t.is_cuda else None
    return _lib.default_context(dev)


def _is_dev(a):
    return isinstance(a, torch.Tensor) and a.is_cuda


def _to_dev(a, ctx):
    """float64 contiguous CUDA tensor for `a` (NumPy array or tensor)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=ctx.torch_device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(ctx.torch_device)


def _axis_np(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy().astype(np.float64)
    return np.asarray(a, dtype=np.float64)


def _extended_axis(coords, lo, hi, min_left=0):
    """Ghost-node extension of one axis (pylamp_trac.py:207-217): while markers lie outside the
    grid, prepend/append one node at the edge spacing.  `min_left`: prepend at least that many (an
    unused ghost node changes nothing: the result is cropped to the real nodes, :313-316)."""
    ax = np.array(coords, dtype=np.float64, copy=True)
    nleft = nright = 0
    while lo < ax[0] or nleft < min_left:
        ax = np.concatenate([[ax[0] - (ax[1] - ax[0])], ax])
        nleft += 1
    while hi > ax[-1]:
        ax = np.concatenate([ax, [ax[-1] + (ax[-1] - ax[-2])]])
        nright += 1
    return ax, nleft, nright


def marker_minmax(tr_x_d, ctx=None):
    ctx = ctx or _ctx(tr_x_d)
    out = (C.c_double * 4)()
    ctx.call("plb_marker_minmax", tr_x_d.shape[0], tr_x_d.data_ptr(), out)
    return list(out)


def trac2grid_device(ctx, tr_x_d, cols_d, schemes, grid, out_d, minmax=None):
    """Device-resident core of trac2grid: `cols_d` is a list of k contiguous (M,) CUDA tensors,
    `out_d` a list of k contiguous (nz,nxx) CUDA tensors (overwritten)."""
    k = len(cols_d)
    M = tr_x_d.shape[0]
    if minmax is None:
        # (several ranks: the extent is all-reduced -- a rank without markers must take part)
        minmax = marker_minmax(tr_x_d, ctx) if (M > 0 or ctx.comm_info()[1] > 1) else [0, 0, 0, 0]
        if not all(np.isfinite(minmax)):
            minmax = [0, 0, 0, 0]
    gz, gx = _axis_np(grid[IZ]), _axis_np(grid[IX])
    axz, lz, rz = _extended_axis(gz, minmax[0], minmax[1])
    axx, lx, rx = _extended_axis(gx, minmax[2], minmax[3])
    axz_d, axx_d = _to_dev(axz, ctx), _to_dev(axx, ctx)
    nz, nxx = out_d[0].shape
    assert nz == gz.shape[0] and nxx == gx.shape[0]
    ctx.call("plb_trac2grid", M, tr_x_d.data_ptr(), k, _lib.ptr_array(cols_d),
             _lib.int_array(schemes), axz_d.data_ptr(), axz.shape[0], axx_d.data_ptr(),
             axx.shape[0], float(axz[0]), float(axz[-1] - axz[0]), float(axx[0]),
             float(axx[-1] - axx[0]), lz, lx, nz, nxx, nxx, _lib.ptr_array(out_d))
    return minmax


def trac2grid_fused_device(ctx, tr_x_d, targets, grid, gridmp, minmax):
    """All marker->grid targets of a time step in one pass over the markers (plb_trac2grid_fused:
    coordinates and every distinct column read once).  `targets`: list of (kind, cols_d, schemes, outs_d)
    with kind 0 = nodes (grid, grid), 1 = centres (gridmp, gridmp), 2 = (gridmp_z, grid_x),
    3 = (grid_z, gridmp_x) -- the four targets of pylamp2.py:309-313.  Returns False (nothing written)
    when the request does not fit the fused kernel (unweighted scheme, markers outside the node grid,
    more than 8 distinct columns): the caller then falls back to `trac2grid_device` per target."""
    gz, gx = _axis_np(grid[IZ]), _axis_np(grid[IX])
    if minmax[0] < gz[0] or minmax[1] > gz[-1] or minmax[2] < gx[0] or minmax[3] > gx[-1]:
        return False
    gmz, gmx = _axis_np(gridmp[IZ]), _axis_np(gridmp[IX])
    nz, nxx = gz.shape[0], gx.shape[0]
    keep, arr = [], (_lib.T2GTarget * len(targets))()
    ext = {}

    def axis(stag, d):
        if (stag, d) not in ext:
            ax, nl, _ = _extended_axis((gmz, gmx)[d] if stag else (gz, gx)[d], minmax[2 * d], minmax[2 * d + 1],
                                       min_left=1 if stag else 0)
            ext[(stag, d)] = (_to_dev(ax, ctx), ax.shape[0], nl)
        return ext[(stag, d)]

    for t, (kind, cols_d, schemes, outs_d) in zip(arr, targets):
        if any((int(s) & INTERP_AVG_WEIGHTED) == 0 for s in schemes) or len(cols_d) > 8:
            return False
        az, ax = axis(kind in (1, 2), IZ), axis(kind in (1, 3), IX)
        t.kind, t.k = int(kind), len(cols_d)
        for f, (c, s, o) in enumerate(zip(cols_d, schemes, outs_d)):
            assert tuple(o.shape) == (nz, nxx) and o.is_contiguous() and c.is_contiguous()
            t.fields[f], t.scheme[f], t.out[f] = c.data_ptr(), int(s), o.data_ptr()
        t.axis_z, t.nze, t.crop_z0 = az[0].data_ptr(), az[1], az[2]
        t.axis_x, t.nxe, t.crop_x0 = ax[0].data_ptr(), ax[1], ax[2]
        keep.append((cols_d, outs_d))
    rc = ctx.lib.plb_trac2grid_fused(ctx.h, tr_x_d.shape[0], tr_x_d.data_ptr(), nz, nxx, nxx, float(gz[0]),
                                     float(gz[-1] - gz[0]), float(gx[0]), float(gx[-1] - gx[0]), len(targets), arr)
    if rc == 3:
        return False
    ctx.check(rc)
    return True


def trac2grid_slab(ctx, tr_x_d, cols_d, schemes, grid, out_d, minmax, bounds, group=None, check=False):
    """`trac2grid_device` for slab-owned markers on several ranks (migrate.py): raw sums of the own
    markers (plb_trac2grid_scatter), boundary rows combined with the two neighbouring slabs, own rows
    finalised (plb_trac2grid_finalise), finished rows all-gathered -- instead of the all-reduce of
    every raw plane inside plb_trac2grid (slabgrid.py).  `minmax` must be the GLOBAL marker extent
    (`marker_minmax` all-reduces it), `bounds` the cell-row bounds of the slabs."""
    import torch.distributed as dist
    from . import slabgrid
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    k = len(cols_d)
    M = tr_x_d.shape[0]
    gz, gx = _axis_np(grid[IZ]), _axis_np(grid[IX])
    axz, lz, rz = _extended_axis(gz, minmax[0], minmax[1])
    axx, lx, rx = _extended_axis(gx, minmax[2], minmax[3])
    axz_d, axx_d = _to_dev(axz, ctx), _to_dev(axx, ctx)
    nz, nxx = out_d[0].shape
    assert nz == gz.shape[0] and nxx == gx.shape[0]
    nze, nxe = axz.shape[0], axx.shape[0]
    planes = torch.empty((k + 2, nze, nxe), dtype=torch.float64, device=tr_x_d.device)
    npl = C.c_int(0)
    ctx.call("plb_trac2grid_scatter", M, tr_x_d.data_ptr(), k, _lib.ptr_array(cols_d), _lib.int_array(schemes),
             axz_d.data_ptr(), nze, axx_d.data_ptr(), nxe, float(axz[0]), float(axz[-1] - axz[0]), float(axx[0]),
             float(axx[-1] - axx[0]), planes.data_ptr(), C.byref(npl))
    p = slabgrid.row_partition(bounds, lz, nze)
    slabgrid.exchange_boundary_rows(planes[:npl.value], p, rank, world, group, check=check)
    q = [0] + [int(b) for b in bounds[1:-1]] + [int(nz)]
    ctx.call("plb_trac2grid_finalise", k, _lib.int_array(schemes), planes.data_ptr(), nze, nxe, lz, lx, nz, nxx,
             nxx, q[rank], q[rank + 1], _lib.ptr_array(out_d))
    slabgrid.gather_rows(out_d, q, rank, world, group)
    return minmax


def trac2grid(tr_x, tr_f, mesh, grid, gridfield, nx, distweight=None, avgscheme=None,
              method=INTERP_METHOD_ELEM, debug=False):
    """Marker-to-node averaging; writes ``gridfield[k][:, :]`` in place and returns None.
    Reference: pylamp_trac.py:161-318 (INTERP_METHOD_ELEM)."""
    assert len(gridfield) == tr_f.shape[1]                                   # :164
    if avgscheme is None:
        avgscheme = [INTERP_AVG_ARITHW for _ in range(len(gridfield))]       # :172-176
    else:
        assert type(avgscheme) == list                                       # :178
        assert len(avgscheme) == len(gridfield)                              # :179
    if not (method & INTERP_METHOD_ELEM):
        raise NotImplementedError("only INTERP_METHOD_ELEM is on the hot path (SURVEY.md §8a-1)")
    ctx = _ctx(tr_x)
    nfield = len(gridfield)
    schemes, live = [], []
    for k, s in enumerate(avgscheme):
        if s & (INTERP_AVG_ARITHMETIC | INTERP_AVG_GEOMETRIC):
            live.append(k), schemes.append(int(s))
        else:
            print("!!! ERROR INVALID AVERAGING SCHEME")                      # :309
    tr_x_d = _to_dev(tr_x, ctx)
    tr_f_d = tr_f if _is_dev(tr_f) else _to_dev(tr_f, ctx)
    minmax = None
    for c0 in range(0, len(live), _lib_max_fields()):
        chunk = live[c0:c0 + _lib_max_fields()]
        cols = [tr_f_d[:, k].contiguous() for k in chunk]
        outs = [gridfield[k] if _is_dev(gridfield[k]) and gridfield[k].is_contiguous()
                else torch.empty(tuple(int(v) for v in gridfield[k].shape), dtype=torch.float64,
                                 device=ctx.torch_device) for k in chunk]
        minmax = trac2grid_device(ctx, tr_x_d, cols, schemes[c0:c0 + len(chunk)], grid, outs, minmax)
        for k, o in zip(chunk, outs):
            if o is gridfield[k]:
                continue
            if isinstance(gridfield[k], torch.Tensor):
                gridfield[k][:, :] = o
            else:
                gridfield[k][:, :] = o.cpu().numpy()                         # :313-316
    return


def _lib_max_fields():
    return 8


def grid2trac_device(ctx, tr_x_d, grid, fields_d, nx, method, defval, outs_d, want_count=True):
    """Device-resident core of grid2trac; fields (nz,nxx) contiguous, outs k x (M,)."""
    gz, gx = _axis_np(grid[IZ]), _axis_np(grid[IX])
    gz_d, gx_d = _to_dev(gz, ctx), _to_dev(gx, ctx)
    nbad = C.c_longlong(0)
    nz, nxx = int(nx[IZ]), int(nx[IX])
    ctx.call("plb_grid2trac", tr_x_d.shape[0], tr_x_d.data_ptr(), int(method), len(fields_d),
             _lib.ptr_array(fields_d), gz_d.data_ptr(), nz, gx_d.data_ptr(), nxx,
             int(fields_d[0].shape[1]), float(gz[0]), float(gz[-1] - gz[0]), float(gx[0]),
             float(gx[-1] - gx[0]), float(defval), _lib.ptr_array(outs_d),
             C.byref(nbad) if want_count else None)
    return int(nbad.value)


def grid2trac(tr_x, tr_f, grid, gridfield, nx, defval=np.nan, method=INTERP_METHOD_LINEAR,
              stopOnError=False):
    """Grid-to-marker interpolation (NEAREST / LINEAR / VELDIV); writes ``tr_f[:, k]`` in place.
    Reference: pylamp_trac.py:30-158."""
    assert len(gridfield) == tr_f.shape[1]                                   # :36
    assert method & (INTERP_METHOD_LINEAR | INTERP_METHOD_NEAREST | INTERP_METHOD_VELDIV)
    nfield = len(gridfield)
    if method & INTERP_METHOD_VELDIV and not method & (INTERP_METHOD_LINEAR | INTERP_METHOD_NEAREST):
        if nfield != 2:                                                      # :122-123
            raise Exception("grid2trac(): method INTERP_METHOD_VELDIV only works in 2D and "
                            "expects field to be (vz,vx)")
    ctx = _ctx(tr_x)
    tr_x_d = _to_dev(tr_x, ctx)
    fields_d = [_to_dev(f, ctx) for f in gridfield]
    M = tr_x_d.shape[0]
    outs = [torch.empty(M, dtype=torch.float64, device=ctx.torch_device) for _ in range(nfield)]
    nbad = grid2trac_device(ctx, tr_x_d, grid, fields_d, nx, method, defval, outs)
    if nbad > 0:
        if stopOnError:
            raise Exception("stopOnError in grid2trac")                      # :54
        print("!!! Warning, grid2trac(): Using default value for extrapolation in ", nbad,
              "tracers")
    for k in range(nfield):
        if isinstance(tr_f, torch.Tensor):
            tr_f[:, k] = outs[k]
        else:
            tr_f[:, k] = outs[k].cpu().numpy()
    return


def rk4_device(ctx, tr_x_d, grids, vz_d, vx_d, nx1, tstep, want_vel=True, spare=False):
    """Device-resident RK4 (plb_rk4).  `spare`: allocate the outputs with a little spare capacity
    (slab-owned marker clouds grow and shrink by migration, see migrate.py)."""
    gz, gx = _axis_np(grids[IZ]), _axis_np(grids[IX])
    gz_d, gx_d = _to_dev(gz, ctx), _to_dev(gx, ctx)
    M = tr_x_d.shape[0]
    if spare:
        from .migrate import empty_rows
        new = lambda: empty_rows(M, (2,), torch.float64, tr_x_d.device)
    else:
        new = lambda: torch.empty_like(tr_x_d)
    x_out = new()
    v_out = new() if want_vel else None
    ctx.call("plb_rk4", M, tr_x_d.data_ptr(), vz_d.data_ptr(), vx_d.data_ptr(), gz_d.data_ptr(),
             int(nx1[IZ]), gx_d.data_ptr(), int(nx1[IX]), int(vz_d.shape[1]), float(gz[0]),
             float(gz[-1] - gz[0]), float(gx[0]), float(gx[-1] - gx[0]), float(tstep),
             x_out.data_ptr(), v_out.data_ptr() if want_vel else None)
    return v_out, x_out


def rk4_fence_count_device(ctx, tr_x_d, grids, vz_d, vx_d, nx1, tstep, nx, L, eps, want_kelem=True):
    """`rk4_device` + fence + cell index + per-cell count of the new positions in one pass over the markers
    (plb_rk4_fence_count; pylamp2.py:550, :558-572, :588-593).  Returns (trac_vel, tr_x_new, kelem, count)."""
    gz, gx = _axis_np(grids[IZ]), _axis_np(grids[IX])
    gz_d, gx_d = _to_dev(gz, ctx), _to_dev(gx, ctx)
    M = tr_x_d.shape[0]
    x_out, v_out = torch.empty_like(tr_x_d), torch.empty_like(tr_x_d)
    nz, nxx = int(nx[IZ]), int(nx[IX])
    kelem = torch.empty(M, dtype=torch.int64, device=tr_x_d.device) if want_kelem else None
    count = torch.empty((nz - 1) * (nxx - 1), dtype=torch.int64, device=tr_x_d.device)
    ctx.call("plb_rk4_fence_count", M, tr_x_d.data_ptr(), vz_d.data_ptr(), vx_d.data_ptr(), gz_d.data_ptr(),
             int(nx1[IZ]), gx_d.data_ptr(), int(nx1[IX]), int(vz_d.shape[1]), float(gz[0]), float(gz[-1] - gz[0]),
             float(gx[0]), float(gx[-1] - gx[0]), float(tstep), x_out.data_ptr(), v_out.data_ptr(), float(L[IZ]),
             float(L[IX]), float(eps), nz, nxx, kelem.data_ptr() if want_kelem else None, count.data_ptr())
    return v_out, x_out, kelem, count


def RK(tr_x, grids, vels, nx, tstep, order=4):
    """Runge-Kutta marker advection with Meyer-Jenny velocity interpolation; returns
    (trac_vel, tr_x_final) as new arrays.  Reference: pylamp_trac.py:321-388 (the reference's
    order-2 branch uses undefined names and cannot run; order 4 uses 1/6 weights, :385)."""
    if order != 2 and order != 4:
        raise Exception("Sorry, don't know how to do that")                  # :326-327
    if len(nx) != 2:
        raise Exception("Sorry, only 2D supported at the moment")            # :329-330
    if order == 2:
        raise NameError("name 'grid' is not defined")     # what the reference does, :336
    ctx = _ctx(tr_x)
    host = not isinstance(tr_x, torch.Tensor)
    tr_x_d = _to_dev(tr_x, ctx)
    vz_d, vx_d = _to_dev(vels[IZ], ctx), _to_dev(vels[IX], ctx)
    nx1 = [int(nx[IZ]) + 1, int(nx[IX]) + 1]
    assert tuple(vz_d.shape) == tuple(nx1)
    v, x = rk4_device(ctx, tr_x_d, grids, vz_d, vx_d, nx1, tstep)
    if host:
        return v.cpu().numpy(), x.cpu().numpy()
    return v, x
