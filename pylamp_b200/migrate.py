"""Spatial (z-slab) marker ownership with per-step migration between the owning slabs
(BASELINE.json north_star: "markers migrate between owning slabs each step").

The reference shares markers among MPI ranks by index (`tr_x[IPROC::NPROC]`, pylamp2.py:445-555)
and all-reduces marker-sized arrays.  Here rank r owns the markers whose cell row
`ielem = floor((nz-1)*z/Lz)` (the cell lookup of pylamp2.py:588) lies in its slab
`[bounds[r], bounds[r+1])` -- the same rows the z-slab Stokes solver gives it.  After advection
and fence a marker has moved at most 0.45 cells (CFL 0.67, pylamp_trac.py:385 weights), so the few
markers that crossed a slab boundary are exchanged:

  1. owner rank per marker from its z coordinate; leavers = markers whose owner is another rank;
  2. leavers are packed into one (n, W) payload ordered by destination (W = 2 coordinates + the
     distinct property columns + 2 velocity components), counts travel with one small
     all-to-all, payloads with one uneven all-to-all (NCCL over NVLink on GPUs, gloo in the CPU tests);
  3. arrivals are written into the leavers' slots, surplus arrivals are appended, surplus holes are
     filled from the tail: only O(movers) entries of the marker arrays are touched, the arrays
     keep a little spare capacity so that growing does not reallocate every step.

Everything here is torch tensor plumbing (index arithmetic, packing, torch.distributed): it runs
unchanged on CUDA tensors in the product and on CPU tensors in the gloo tests.  The marker kernels
(trac2grid, grid2trac, RK4, ...) give order-independent results, so moving markers between slots
and ranks changes nothing but the summation order of trac2grid.
"""
import torch
import torch.distributed as dist

SLACK = 0.02      # spare capacity kept when a marker array has to grow
GRANULE = 1 << 20  # capacities are rounded up to this many rows: a slab's marker count changes by a few rows every step,
                   # and allocations of ever-changing sizes defeat the caching allocator (measured on 4 GPUs: 12 ms per
                   # step of cudaMalloc/cudaFree in the RK4 phase)


def _capacity(n, slack=SLACK):
    cap = int(n) + int(n * slack) + 1024
    return ((cap + GRANULE - 1) // GRANULE) * GRANULE if cap > GRANULE else cap


def slab_bounds(ncell_z, world):
    """Cell-row range [bounds[r], bounds[r+1]) owned by rank r (even split, like the slab solver)."""
    return [(r * int(ncell_z)) // int(world) for r in range(int(world) + 1)]


def owner_of(z, nz, Lz, bounds):
    """Owner rank of every marker from its z coordinate: cell row as in pylamp2.py:588
    (multiply, divide, floor -- same IEEE operations as the per-cell count), clamped to the
    grid, then located in `bounds`."""
    ncell = int(nz) - 1
    ie = torch.floor((ncell * z) / float(Lz)).to(torch.int64).clamp_(0, ncell - 1)
    inner = torch.as_tensor(bounds[1:-1], dtype=torch.int64, device=z.device)
    return torch.bucketize(ie, inner, right=True)


def resize_rows(t, n, slack=SLACK):
    """`t` with its first dimension changed to `n`, keeping the leading rows.  Shrinking narrows
    the view (the storage keeps its room), growing reuses spare room in the storage if there is
    any and otherwise reallocates with `slack` spare capacity."""
    n = int(n)
    if n <= t.shape[0]:
        return t[:n]
    row = 1
    for d in t.shape[1:]:
        row *= int(d)
    st = t.untyped_storage()
    need = (t.storage_offset() + n * row) * t.element_size()
    if t.is_contiguous() and st.nbytes() >= need:
        return t.new_empty(0).set_(st, t.storage_offset(), (n,) + tuple(t.shape[1:]))
    cap = _capacity(n, slack)
    new = torch.empty((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    new[:t.shape[0]] = t
    return new[:n]


def empty_rows(n, tail_shape, dtype, device, slack=SLACK):
    """Uninitialised (n, *tail_shape) view of an allocation with spare capacity."""
    cap = _capacity(n, slack)
    return torch.empty((cap,) + tuple(tail_shape), dtype=dtype, device=device)[:int(n)]


def compaction_plan(M_old, holes, leave_mask, n_arr):
    """Where arrivals go and which tail entries move, for `holes` (ascending positions of the
    leavers) and `n_arr` arrivals.  Returns (M_new, fill_pos, move_src, move_dst):
    arrival i is written to fill_pos[i]; then entries move_src are copied to move_dst; the arrays
    are finally cut to M_new rows."""
    n_leave = int(holes.numel())
    M_new = M_old - n_leave + n_arr
    empty = holes.new_empty(0)
    if n_arr >= n_leave:
        extra = torch.arange(M_old, M_new, dtype=torch.int64, device=holes.device)
        return M_new, torch.cat([holes, extra]), empty, empty
    rest = holes[n_arr:]
    move_dst = rest[rest < M_new]
    move_src = torch.nonzero(~leave_mask[M_new:M_old]).flatten() + M_new
    return M_new, holes[:n_arr], move_src, move_dst


def _distinct(cols):
    """Distinct tensors among the property columns (columns may alias one another) and the
    index of each column's tensor in that list."""
    uniq, where, seen = [], [], {}
    for c in cols:
        key = (c.data_ptr(), c.shape[0])
        if key not in seen:
            seen[key] = len(uniq)
            uniq.append(c)
        where.append(seen[key])
    return uniq, where


def migrate(s, group=None, bounds=None):
    """Send every marker of `s` that left this rank's slab to its new owner and take in the
    arrivals.  `s` needs `tr_x` (M,2), `cols` (list of (M,) tensors), `nx`, `L` and optionally
    `trac_vel` (M,2 or None).  Returns {"sent": n, "received": n, "markers": M_new}."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if bounds is None:
        bounds = slab_bounds(s.nx[0] - 1, world)
    dev = s.tr_x.device
    M_old = int(s.tr_x.shape[0])
    owner = owner_of(s.tr_x[:, 0], s.nx[0], s.L[0], bounds)
    leave_mask = owner != rank
    holes = torch.nonzero(leave_mask).flatten()              # ascending
    dest = owner[holes]
    order = torch.argsort(dest, stable=True)
    send_idx = holes[order]
    uniq, where = _distinct(s.cols)
    vel = getattr(s, "trac_vel", None)
    if vel is not None and vel.shape[0] != M_old:
        vel = None
    W = 2 + len(uniq) + (2 if vel is not None else 0)
    # counts (and the payload width, which every rank must agree on) with one small all-to-all
    meta = torch.empty((world, 2), dtype=torch.int64, device=dev)
    meta[:, 0] = torch.bincount(dest, minlength=world)
    meta[:, 1] = W
    meta_in = torch.empty_like(meta)
    dist.all_to_all_single(meta_in.view(-1), meta.view(-1), group=group)
    send_counts = [int(v) for v in meta[:, 0].tolist()]
    recv_counts = [int(v) for v in meta_in[:, 0].tolist()]
    if any(int(w) != W for w in meta_in[:, 1].tolist()):
        raise RuntimeError("migrate: ranks disagree on the marker payload layout")
    n_leave, n_arr = sum(send_counts), sum(recv_counts)
    # payload: one row per leaver, ordered by destination
    send = torch.empty((n_leave, W), dtype=torch.float64, device=dev)
    send[:, 0:2] = s.tr_x.index_select(0, send_idx)
    for j, c in enumerate(uniq):
        send[:, 2 + j] = c.index_select(0, send_idx)
    if vel is not None:
        send[:, 2 + len(uniq):] = vel.index_select(0, send_idx)
    recv = torch.empty((n_arr, W), dtype=torch.float64, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts,
                           group=group)
    M_new, fill_pos, move_src, move_dst = compaction_plan(M_old, holes, leave_mask, n_arr)

    def apply(t, arrivals):
        t = resize_rows(t, max(M_old, M_new))
        if n_arr:
            t.index_copy_(0, fill_pos, arrivals)
        if move_src.numel():
            t.index_copy_(0, move_dst, t.index_select(0, move_src))
        return t[:M_new]

    s.tr_x = apply(s.tr_x, recv[:, 0:2])
    new_uniq = [apply(c, recv[:, 2 + j]) for j, c in enumerate(uniq)]
    s.cols = [new_uniq[w] for w in where]
    if vel is not None:
        s.trac_vel = apply(vel, recv[:, 2 + len(uniq):])
    return {"sent": n_leave, "received": n_arr, "markers": M_new}


def migrate_native(s, bounds=None):
    """`migrate` on CUDA state with the library's kernels (csrc/migrate.cu: plb_migrate_plan + plb_migrate_apply):
    the leavers are listed and packed by kernels and travel to the slab below / above with one grouped
    ncclSend/ncclRecv pair per neighbour; arrivals fill the leavers' slots.  Same result as `migrate` up to the
    order of the markers.  Returns {"sent": n, "received": n, "markers": M_new}."""
    import ctypes as C
    from . import _lib
    ctx = s.ctx
    rank, world = ctx.comm_info()
    if bounds is None:
        bounds = slab_bounds(s.nx[0] - 1, world)
    M = int(s.tr_x.shape[0])
    counts = (C.c_longlong * 4)()
    ctx.call("plb_migrate_plan", M, s.tr_x.data_ptr(), int(s.nx[0]), float(s.L[0]), int(bounds[rank]),
             int(bounds[rank + 1]), counts)
    n_leave, n_arr = counts[0] + counts[1], counts[2] + counts[3]
    uniq, where = _distinct(s.cols)
    vel = getattr(s, "trac_vel", None)
    if vel is not None and vel.shape[0] != M:
        vel = None
    arrays = [s.tr_x] + uniq + ([vel] if vel is not None else [])
    need = max(M, M - n_leave + n_arr, M + max(0, n_arr - n_leave))
    arrays = [resize_rows(t.contiguous() if not t.is_contiguous() else t, need) for t in arrays]   # capacity; keeps the data
    widths = [1 if t.dim() == 1 else int(t.shape[1]) for t in arrays]
    M_new = C.c_longlong(0)
    ctx.call("plb_migrate_apply", M, len(arrays), _lib.ptr_array(arrays), _lib.int_array(widths), need, int(s.nx[0]),
             float(s.L[0]), int(bounds[rank]), int(bounds[rank + 1]), C.byref(M_new))
    n = int(M_new.value)
    arrays = [t[:n] for t in arrays]
    s.tr_x = arrays[0]
    new_uniq = arrays[1:1 + len(uniq)]
    s.cols = [new_uniq[w] for w in where]
    if vel is not None:
        s.trac_vel = arrays[-1]
    return {"sent": int(n_leave), "received": int(n_arr), "markers": n}


def check_ownership(s, group=None, bounds=None):
    """Number of local markers that lie outside this rank's slab (0 after `migrate`)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if bounds is None:
        bounds = slab_bounds(s.nx[0] - 1, world)
    return int((owner_of(s.tr_x[:, 0], s.nx[0], s.L[0], bounds) != rank).sum().item())
