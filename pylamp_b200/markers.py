"""Device-resident wrappers for the driver-inline steps of the reference loop body
(pylamp2.py:291-303, :339-366, :471-480, :491-545, :558-572, :588-593).  All arguments are torch
CUDA float64 tensors; everything runs in the CUDA library (no CPU path)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .pylamp_const import EPS, GASR, IX, IZ


def _ctx(t):
    return _lib.default_context(t.device.index)


def cell_index_count(tr_x, nx, L, want_kelem=True):
    """kelem (M,) int64 and per-cell count ((nz-1)*(nxx-1),) int64 -- pylamp2.py:588-593."""
    ctx = _ctx(tr_x)
    M = tr_x.shape[0]
    nz, nxx = int(nx[IZ]), int(nx[IX])
    kelem = torch.empty(M, dtype=torch.int64, device=tr_x.device) if want_kelem else None
    count = torch.empty((nz - 1) * (nxx - 1), dtype=torch.int64, device=tr_x.device)
    ctx.call("plb_cell_index_count", M, tr_x.data_ptr(), nz, nxx, float(L[IZ]), float(L[IX]),
             kelem.data_ptr() if want_kelem else None, count.data_ptr())
    return kelem, count


def fence(tr_x, L, eps=EPS, bc=None):
    """In-place fence to [eps, L-eps] -- pylamp2.py:558-572.  `bc`: the Stokes wall types [z=0, x=0, z=L, x=L];
    markers beyond a BC_TYPE_FLOWTHRU wall are not fenced (they are removed afterwards, :573-581)."""
    walls = 15
    if bc is not None:
        walls = sum(1 << w for w in range(4) if not (int(bc[w]) & 4))
    _ctx(tr_x).call("plb_fence_walls", tr_x.shape[0], tr_x.data_ptr(), float(L[IZ]), float(L[IX]), float(eps), walls)


def fence_count(tr_x, nx, L, eps=EPS, want_kelem=True):
    """`fence` then `cell_index_count` in one pass over the coordinates (pylamp2.py:558-572, :588-593)."""
    ctx = _ctx(tr_x)
    M = tr_x.shape[0]
    nz, nxx = int(nx[IZ]), int(nx[IX])
    kelem = torch.empty(M, dtype=torch.int64, device=tr_x.device) if want_kelem else None
    count = torch.empty((nz - 1) * (nxx - 1), dtype=torch.int64, device=tr_x.device)
    ctx.call("plb_fence_count", M, tr_x.data_ptr(), float(L[IZ]), float(L[IX]), float(eps), nz, nxx,
             kelem.data_ptr() if want_kelem else None, count.data_ptr())
    return kelem, count


def delete_outside(s):
    """With the fence disabled (or beyond a flow-through wall) the reference removes every marker that left the box
    (pylamp2.py:563-581: x <= 0 or x >= L in either direction -> TR__ID = -1 -> np.delete on tr_x, tr_f and
    trac_vel).  Same set of survivors here; the survivors of the tail are moved into the holes instead of shifting
    every array (marker order is free: every kernel is order-independent).  CUDA state: the library's kernels
    (plb_delete_outside, csrc/migrate.cu); CPU tensors (host-logic tests): torch indexing.
    Returns the number of markers removed."""
    from .migrate import _distinct, compaction_plan
    x, L = s.tr_x, s.L
    M = int(x.shape[0])
    vel = getattr(s, "trac_vel", None)
    if vel is not None and vel.shape[0] != M:
        vel = None
    uniq, where = _distinct(s.cols)
    if x.is_cuda:
        arrays = [x if x.is_contiguous() else x.contiguous()] + [c if c.is_contiguous() else c.contiguous() for c in uniq]
        if vel is not None:
            arrays.append(vel if vel.is_contiguous() else vel.contiguous())
        widths = [1 if t.dim() == 1 else int(t.shape[1]) for t in arrays]
        M_new = C.c_longlong(0)
        _ctx(x).call("plb_delete_outside", M, len(arrays), _lib.ptr_array(arrays), _lib.int_array(widths), float(L[IZ]),
                     float(L[IX]), C.byref(M_new))
        n = int(M_new.value)
        arrays = [t[:n] for t in arrays]
        s.tr_x = arrays[0]
        new_uniq = arrays[1:1 + len(uniq)]
        s.cols = [new_uniq[w] for w in where]
        if vel is not None:
            s.trac_vel = arrays[-1]
        return M - n
    outside = (x[:, 0] <= 0) | (x[:, 0] >= L[IZ]) | (x[:, 1] <= 0) | (x[:, 1] >= L[IX])
    holes = torch.nonzero(outside).flatten()
    n = int(holes.numel())
    if n == 0:
        return 0
    M_new, _, src, dst = compaction_plan(M, holes, outside, 0)

    def cut(t):
        if src.numel():
            t.index_copy_(0, dst, t.index_select(0, src))
        return t[:M_new]

    s.tr_x = cut(s.tr_x)
    new_uniq = [cut(c) for c in uniq]
    s.cols = [new_uniq[w] for w in where]
    if vel is not None:
        s.trac_vel = cut(vel)
    return n


def update_properties(T, rho0, alpha, Ea, eta0, tdep_rho, tdep_eta, Tref, etamin, etamax,
                      rho_out=None, eta_out=None):
    """rho(T), eta(T) on markers -- pylamp2.py:291-303."""
    ctx = _ctx(T)
    rho = rho_out if rho_out is not None else torch.empty_like(T)
    eta = eta_out if eta_out is not None else torch.empty_like(T)
    ctx.call("plb_update_properties", T.shape[0], int(bool(tdep_rho)), int(bool(tdep_eta)),
             float(Tref), float(etamin), float(etamax), float(GASR), T.data_ptr(), rho0.data_ptr(),
             alpha.data_ptr(), Ea.data_ptr(), eta0.data_ptr(), rho.data_ptr(), eta.data_ptr())
    return rho, eta


def centre_velocities(vz, vx, bc):
    """(nz+1, nxx+1) cell-centre velocities with the BC ghost ring -- pylamp2.py:491-545."""
    ctx = _ctx(vz)
    nz, nxx = vz.shape
    vzc = torch.empty((nz + 1, nxx + 1), dtype=torch.float64, device=vz.device)
    vxc = torch.empty_like(vzc)
    ctx.call("plb_centre_velocities", nz, nxx, vz.stride(0), vz.data_ptr(), vx.data_ptr(),
             _lib.int_array(bc), nxx + 1, vzc.data_ptr(), vxc.data_ptr())
    return vzc, vxc


def field_max(f):
    """Signed maximum of a 2-D field (np.max of pylamp2.py:364, quirk 6)."""
    out = C.c_double(0)
    _ctx(f).call("plb_field_max", f.shape[0], f.shape[1], f.stride(0), f.data_ptr(), C.byref(out))
    return out.value


def max_diffusivity2(kz, rho, cp):
    """max(2*kz/(rho*cp)) -- pylamp2.py:340."""
    out = C.c_double(0)
    _ctx(kz).call("plb_max_diffusivity2", kz.shape[0], kz.shape[1], kz.stride(0), kz.data_ptr(),
                  rho.data_ptr(), cp.data_ptr(), C.byref(out))
    return out.value


def subgrid_stage1(dt, dz, dx, Told, T, cp, rho, k):
    """Tsg, dT of pylamp2.py:472-475."""
    Tsg, dT = torch.empty_like(T), torch.empty_like(T)
    _ctx(T).call("plb_subgrid_stage1", T.shape[0], float(dt), float(dz), float(dx), Told.data_ptr(),
                 T.data_ptr(), cp.data_ptr(), rho.data_ptr(), k.data_ptr(), Tsg.data_ptr(),
                 dT.data_ptr())
    return Tsg, dT


def subgrid_stage2(Tsg, back, T_out):
    """T = Tsg - back, pylamp2.py:480."""
    _ctx(Tsg).call("plb_subgrid_stage2", Tsg.shape[0], Tsg.data_ptr(), back.data_ptr(), T_out.data_ptr())


def sort_by_cell(tr_x, cols, nx, L, extra=(), want_cell_start=False, consume=False, ws=None):
    """Physically re-order the marker arrays by cell index (cell-major, like the setups generate
    them) with the device counting sort of csrc/sort.cu (plb_sort_plan + one plb_permute per distinct
    array).  The results of every kernel are order-independent (up to fp64 summation order in
    trac2grid); a cell-ordered cloud keeps trac2grid's run aggregation effective -- one reduction per
    cell run instead of one per marker -- and the grid gathers of grid2trac/RK4 cache-local.
    Returns (tr_x, cols, extra) as new tensors (+ the (ncell+1,) int32 first-slot table if asked).
    `extra`: further (M,) or (M,2) float64 tensors to carry along (e.g. marker velocities).
    `consume=True`: the input tensors may be overwritten (each permuted array's old storage becomes the
    next array's destination: one spare array instead of a second copy of the whole cloud).
    `ws` (a dict the caller keeps between calls, with consume=True): the slot table and the spare arrays
    stay allocated from one sort to the next -- nothing is allocated in a time loop's later sorts."""
    ctx = _ctx(tr_x)
    M = tr_x.shape[0]
    nz, nxx = int(nx[IZ]), int(nx[IX])
    keep = ws if (ws is not None and consume) else None
    dest = keep.get("dest") if keep is not None else None
    if dest is None or dest.shape[0] != M or dest.device != tr_x.device:
        dest = torch.empty(M, dtype=torch.int32, device=tr_x.device)
        if keep is not None:
            keep["dest"] = dest
    start = torch.empty((nz - 1) * (nxx - 1) + 1, dtype=torch.int32, device=tr_x.device) if want_cell_start else None
    ctx.call("plb_sort_plan", M, tr_x.data_ptr(), nz, nxx, float(L[IZ]), float(L[IX]), dest.data_ptr(),
             start.data_ptr() if want_cell_start else None)
    # the array permuted last becomes the next one's destination (one spare per shape)
    spare = keep.setdefault("spare", {}) if keep is not None else {}

    def perm(t):
        width = 1 if t.dim() == 1 else int(t.shape[1])
        out = spare.pop(width, None) if consume else None
        if out is None or out.shape != t.shape or out.device != t.device or out.data_ptr() == t.data_ptr():
            out = torch.empty_like(t)
        ctx.call("plb_permute", M, dest.data_ptr(), t.data_ptr(), out.data_ptr(), width)
        if consume:
            spare[width] = t
        return out

    tr_x = perm(tr_x.contiguous())
    seen, out = {}, []
    for c in cols:                      # columns may alias each other (shared zero column)
        key = c.data_ptr()
        if key not in seen:
            seen[key] = perm(c)
        out.append(seen[key])
    extra = [perm(e.contiguous()) for e in extra]
    if want_cell_start:
        return tr_x, out, extra, start
    return tr_x, out, extra


def subgrid_fused(stage, tr_x, grid, field, T, dt=0.0, dz=1.0, dx=1.0, cp=None, rho=None, k=None, Tsg=None,
                  dT=None):
    """Fused marker temperature update + subgrid diffusion (pylamp2.py:448-480), one pass per stage.
    stage 1 returns (Tsg, dT) from T (left untouched) and field = T_new - T_grid;
    stage 2 writes T = Tsg - interp(field = f_sgc).  Raises like grid2trac(stopOnError=True)."""
    import numpy as np
    ctx = _ctx(tr_x)
    gz, gx = np.asarray(grid[IZ], dtype=np.float64), np.asarray(grid[IX], dtype=np.float64)
    gz_d = torch.as_tensor(gz).to(tr_x.device)
    gx_d = torch.as_tensor(gx).to(tr_x.device)
    if stage == 1:
        Tsg, dT = torch.empty_like(T), torch.empty_like(T)
    nbad = C.c_longlong(0)
    p = lambda t: t.data_ptr() if t is not None else None
    ctx.call("plb_subgrid_fused", stage, tr_x.shape[0], tr_x.data_ptr(), field.data_ptr(), gz_d.data_ptr(),
             field.shape[0], gx_d.data_ptr(), field.shape[1], field.stride(0), float(gz[0]),
             float(gz[-1] - gz[0]), float(gx[0]), float(gx[-1] - gx[0]), float(dt), float(dz), float(dx),
             T.data_ptr(), p(cp), p(rho), p(k), p(Tsg), p(dT), C.byref(nbad))
    if nbad.value:
        raise Exception("stopOnError in grid2trac")
    return Tsg, dT


def inject_markers(s, tracdens, tracdens_min, generator=None, cell_rows=None, group=None, seed=None):
    """Marker injection into under-populated cells -- pylamp2.py:594-633.  CUDA state: the library's kernels
    (csrc/inject.cu: plan by scans over the cells, cell means by one pass over the markers, positions from a
    counter-based Philox stream); CPU tensors (gloo tests of the multi-rank host logic): `inject_markers_torch`.
    Same counts, cells, cell-mean properties and ids as the reference's loop; positions are random inside the same
    cells.  Returns the number of markers injected on this rank."""
    if not s.tr_x.is_cuda:
        return inject_markers_torch(s, tracdens, tracdens_min, generator=generator, cell_rows=cell_rows, group=group)
    from .migrate import _distinct, resize_rows
    from .pylamp_const import TR__ID
    ctx = s.ctx
    nz, nxx = int(s.nx[IZ]), int(s.nx[IX])
    ncx, ncell = nxx - 1, (nz - 1) * (nxx - 1)
    dev = s.tr_x.device
    c0, c1 = (0, ncell) if cell_rows is None else (cell_rows[0] * ncx, cell_rows[1] * ncx)
    plan = (C.c_longlong * 3)()
    ctx.call("plb_inject_plan", ncell, s.count.data_ptr(), int(tracdens), int(tracdens_min), int(c0), int(c1), plan)
    n_def, n_new, id_adv = int(plan[0]), int(plan[1]), int(plan[2])
    M0 = int(s.tr_x.shape[0])
    max0 = s.cols[TR__ID].max() if M0 else torch.tensor(-1.0, dtype=torch.float64, device=dev)
    id_offset = 0.0
    if cell_rows is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        max0 = max0.clone()
        dist.all_reduce(max0, op=dist.ReduceOp.MAX, group=group)
        adv = torch.zeros(world, dtype=torch.float64, device=dev)
        adv[rank] = float(id_adv)
        dist.all_reduce(adv, group=group)
        id_offset = float(adv[:rank].sum().item())      # ids continue over the ranks in cell order
    if n_new == 0:
        return 0
    # the id column must not share storage with another column once new ids are written
    id_aliased = any(k != TR__ID and s.cols[k].data_ptr() == s.cols[TR__ID].data_ptr() for k in range(len(s.cols)))
    if id_aliased:
        s.cols = list(s.cols)
        s.cols[TR__ID] = s.cols[TR__ID].clone()
    uniq, where = _distinct(s.cols)
    uniq = [resize_rows(c if c.is_contiguous() else c.contiguous(), M0 + n_new) for c in uniq]
    s.tr_x = resize_rows(s.tr_x, M0 + n_new)
    gz = torch.as_tensor(np.asarray(s.grid[IZ], dtype=np.float64)).to(dev)
    gx = torch.as_tensor(np.asarray(s.grid[IX], dtype=np.float64)).to(dev)
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device="cpu" if generator is None or generator.device.type == "cpu" else dev).item())
    s._inject_stream = getattr(s, "_inject_stream", 0)
    ctx.call("plb_inject_apply", M0, s.kelem.data_ptr(), s.count.data_ptr(), s.tr_x.data_ptr(), len(uniq),
             _lib.ptr_array(uniq), int(where[TR__ID]), float(max0.item()) + id_offset, gz.data_ptr(), gx.data_ptr(), nxx,
             C.c_ulonglong(seed), C.c_ulonglong(s._inject_stream))
    s._inject_stream += n_new
    s.cols = [uniq[w] for w in where]
    return n_new


def inject_markers_torch(s, tracdens, tracdens_min, generator=None, cell_rows=None, group=None):
    """Marker injection into under-populated cells on the device -- pylamp2.py:594-633.

    Every cell with fewer than `tracdens_min` markers receives `tracdens - count` new markers at
    uniformly random positions inside the cell; their properties are the mean of the cell's existing
    markers (NaN for an empty cell, like the reference's 0/0), their TR__ID continues from the
    current maximum exactly as the reference's per-cell loop does (each cell's first new id repeats
    the running maximum, :614-615).  `s.kelem` / `s.count` must be current (end of `driver.timestep`).
    The reference draws positions from NumPy's global Mersenne-Twister stream, so positions (not
    counts, cells or properties) differ: parity is at the count/property level.
    Implemented with torch tensor operations (device-side plumbing; no new kernel).

    Several ranks with slab-owned markers (migrate.py): `cell_rows` = (first, last+1) cell row of
    this rank's slab -- only those cells are served (their local counts are complete), and the ids
    continue over the ranks in cell order like the reference's single loop (collective: every rank
    must call).  Returns the number of markers injected on this rank."""
    from .migrate import resize_rows
    from .pylamp_const import NFTRAC, TR__ID
    nz, nxx = int(s.nx[IZ]), int(s.nx[IX])
    ncx = nxx - 1
    dev = s.tr_x.device
    count, kelem = s.count, s.kelem
    few = count < tracdens_min
    if cell_rows is not None:
        own = torch.zeros_like(few)
        own[cell_rows[0] * ncx:cell_rows[1] * ncx] = True
        few &= own
    kmiss = torch.nonzero(few).flatten()
    nmiss = (tracdens - count[kmiss]).to(torch.int64)
    M0 = s.tr_x.shape[0]
    max0 = s.cols[TR__ID].max() if M0 else torch.tensor(-1.0, dtype=torch.float64, device=dev)
    id_offset = 0.0
    if cell_rows is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        max0 = max0.clone()
        dist.all_reduce(max0, op=dist.ReduceOp.MAX, group=group)
        sums = torch.zeros(world, dtype=torch.float64, device=dev)
        sums[rank] = (nmiss - 1).sum().to(torch.float64)
        dist.all_reduce(sums, group=group)
        id_offset = float(sums[:rank].sum().item())
    if kmiss.numel() == 0:
        return 0
    total = int(nmiss.sum().item())
    ncell_m = kmiss.numel()
    # existing markers of the deficient cells -> per-cell property means
    lut = torch.full((count.numel(),), -1, dtype=torch.int64, device=dev)
    lut[kmiss] = torch.arange(ncell_m, device=dev)
    slot = lut[kelem]
    idx = torch.nonzero(slot >= 0).flatten()
    slot = slot[idx]
    cnt = count[kmiss].to(torch.float64)
    rep = torch.repeat_interleave(torch.arange(ncell_m, device=dev), nmiss)
    # ids: cell c starts at the running maximum: max0 + sum_{c'<c} (n_c' - 1)
    start = max0 + id_offset + torch.cumsum(nmiss - 1, 0).to(torch.float64) - (nmiss - 1).to(torch.float64)
    within = torch.arange(total, device=dev, dtype=torch.float64) - \
        torch.repeat_interleave((torch.cumsum(nmiss, 0) - nmiss).to(torch.float64), nmiss)
    new_cols, seen = [], {}
    id_aliased = any(k != TR__ID and s.cols[k].data_ptr() == s.cols[TR__ID].data_ptr() for k in range(NFTRAC))
    for k in range(NFTRAC):
        col = s.cols[k]
        if k == TR__ID:
            new = start[rep] + within
            if id_aliased:                               # the id column shares storage with another one: un-alias
                new_cols.append(torch.cat([col, new]))
            else:
                out = resize_rows(col, M0 + total)
                out[M0:] = new
                new_cols.append(out)
            continue
        key = col.data_ptr()
        if key in seen:                                  # aliased (shared) columns stay aliased
            new_cols.append(seen[key])
            continue
        sums = torch.zeros(ncell_m, dtype=torch.float64, device=dev).index_add_(0, slot, col[idx])
        new = (sums / cnt)[rep]                          # 0/0 = NaN for empty cells, like the reference
        out = resize_rows(col, M0 + total)               # appended in place while the spare capacity lasts
        out[M0:] = new
        seen[key] = out
        new_cols.append(out)
    gz = torch.as_tensor(np.asarray(s.grid[IZ], dtype=np.float64)).to(dev)
    gx = torch.as_tensor(np.asarray(s.grid[IX], dtype=np.float64)).to(dev)
    im, jm = (kmiss // ncx)[rep], (kmiss % ncx)[rep]
    u = torch.rand((total, 2), dtype=torch.float64, device=dev, generator=generator)
    xt = torch.empty((total, 2), dtype=torch.float64, device=dev)
    xt[:, IX] = u[:, IX] * (gx[jm + 1] - gx[jm]) + gx[jm]
    xt[:, IZ] = u[:, IZ] * (gz[im + 1] - gz[im]) + gz[im]
    s.tr_x = resize_rows(s.tr_x, M0 + total)
    s.tr_x[M0:] = xt
    s.cols = new_cols
    return total
