"""Node sums of slab-owned markers without the full-plane all-reduce ("trac2grid
halo-row accumulate").

With `marker_ownership="slab"` (migrate.py) the markers of rank r lie in its cell rows
[b[r], b[r+1]), so their raw node sums (sum of weights, sum of weight*value; pylamp_trac.py:252-298)
only touch the target rows of that slab plus at most one row on either side.  Instead of
all-reducing every raw plane over all ranks (what `plb_trac2grid` does for index-owned markers):

  1. every rank adds the few rows its markers wrote into a neighbour's part to that neighbour's
     sums (`exchange_boundary_rows`: one send/recv pair per neighbour, H rows of every plane);
  2. every rank divides / exponentiates its own rows only (`plb_trac2grid_finalise` on a row range);
  3. the finished rows are all-gathered so that every rank holds the whole field again
     (`gather_rows`: the solver and the marker kernels read replicated grids).

Raw planes are (nplanes, nze, nxe) tensors over the extended target axes (ghost nodes included,
pylamp_trac.py:207-220); `lz` is the number of ghost rows prepended in z.  Pure torch +
torch.distributed plumbing: runs on CUDA tensors over NCCL and on CPU tensors over gloo (tests).
"""
import torch
import torch.distributed as dist

HALO = 2      # rows exchanged with each neighbour (one is needed when ownership is exact)


def row_partition(bounds, lz, nze):
    """Rows of the extended planes that rank r finalises: [p[r], p[r+1]).  Interior cuts follow the
    cell-row bounds of the slabs (shifted by the prepended ghost rows); the first rank also takes
    the ghost rows below, the last one everything above."""
    world = len(bounds) - 1
    p = [0] + [int(bounds[r]) + int(lz) for r in range(1, world)] + [int(nze)]
    return p


def exchange_boundary_rows(planes, p, rank, world, group=None, halo=HALO, check=False):
    """Add the rows this rank's markers wrote into the neighbours' parts to the neighbours' sums and
    take in theirs.  `planes`: (nplanes, nze, nxe), modified in place: on return the rows
    [p[rank], p[rank+1]) hold the complete sums (other rows are meaningless)."""
    lo, hi = p[rank], p[rank + 1]
    nze = planes.shape[1]
    if check:
        far = (planes[:, :max(lo - halo, 0)].abs().sum() + planes[:, min(hi + halo, nze):].abs().sum()).item()
        if far != 0:
            raise RuntimeError("exchange_boundary_rows: sums beyond the halo rows -- markers are not slab-owned")
    ops, recv_dn, recv_up = [], None, None
    if rank > 0:
        send_dn = planes[:, max(lo - halo, 0):lo].contiguous()                 # rows of rank-1's part
        recv_dn = torch.empty_like(planes[:, lo:min(lo + halo, hi)])           # its sums for my first rows
        ops += [dist.P2POp(dist.isend, send_dn, _peer(rank - 1, group), group),
                dist.P2POp(dist.irecv, recv_dn, _peer(rank - 1, group), group)]
    if rank < world - 1:
        send_up = planes[:, hi:min(hi + halo, nze)].contiguous()
        recv_up = torch.empty_like(planes[:, max(hi - halo, lo):hi])
        ops += [dist.P2POp(dist.isend, send_up, _peer(rank + 1, group), group),
                dist.P2POp(dist.irecv, recv_up, _peer(rank + 1, group), group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if recv_dn is not None:
        planes[:, lo:lo + recv_dn.shape[1]] += recv_dn
    if recv_up is not None:
        planes[:, hi - recv_up.shape[1]:hi] += recv_up
    return planes


def _peer(r, group):
    return r if group is None else dist.get_global_rank(group, r)


def gather_rows(fields, q, rank, world, group=None):
    """All-gather of row blocks: on entry rank r holds rows [q[r], q[r+1]) of every (nz, nxx) tensor in
    `fields`, on return every rank holds all rows.  One all-gather for the whole list (blocks padded
    to the tallest one)."""
    k = len(fields)
    nxx = fields[0].shape[1]
    hmax = max(q[r + 1] - q[r] for r in range(world))
    send = torch.zeros((k, hmax, nxx), dtype=fields[0].dtype, device=fields[0].device)
    for f, t in enumerate(fields):
        send[f, :q[rank + 1] - q[rank]] = t[q[rank]:q[rank + 1]]
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    for r in range(world):
        if r == rank:
            continue
        for f, t in enumerate(fields):
            t[q[r]:q[r + 1]] = recv[r][f, :q[r + 1] - q[r]]
    return fields
