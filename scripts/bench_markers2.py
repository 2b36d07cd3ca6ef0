"""Round-2 marker micro-benchmark (CUDA events on the launching stream, every array larger than L2):
the fused step trac2grid (plb_trac2grid_fused: 4 targets, 7 distinct columns, 72 B/marker) with 1 / 2 / 4
lanes per run, against the five separate calls of round 1; on the cell-ordered cloud, on the same cloud after
an advection-like displacement of a few cells, and on the displaced cloud after the device sort; the sort
itself; the subgrid call; RK4.  Prints one JSON object.
  python scripts/bench_markers2.py [ncell=2048] [reps=5]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylamp_b200 import _lib, markers, pylamp_trac as T, setups  # noqa: E402
from pylamp_b200.pylamp_const import IX, IZ, TR_ETA, TR_HCD, TR_HCP, TR_IHT, TR_MAT, TR_RHO, TR_TMP  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = _lib.default_context(0)
nx, L, tr_x, cols, opts = setups.convection_device(ncell=ncell, per_side=4, device="cuda:0")
grid, gridmp = setups.make_grids(nx, L)
M = tr_x.shape[0]
cols[TR_RHO].fill_(3300.0)
cols[TR_ETA].copy_(1e20 * (1.0 + tr_x[:, 0] / L[0]))
new = lambda: torch.empty(tuple(nx), dtype=torch.float64, device="cuda")
out6, o_n, o_kz, o_kx, o_sg = [new() for _ in range(6)], new(), new(), new(), new()
dz = L[0] / ncell


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


node_cols = [cols[k] for k in (TR_RHO, TR_ETA, TR_HCP, TR_TMP, TR_IHT, TR_MAT)]
node_sch = [5, 6, 5, 5, 5, 5]


def fused(x, mm):
    assert T.trac2grid_fused_device(ctx, x, [(0, node_cols, node_sch, out6), (1, [cols[TR_ETA]], [6], [o_n]),
                                             (2, [cols[TR_HCD]], [5], [o_kz]), (3, [cols[TR_HCD]], [5], [o_kx])],
                                    grid, gridmp, mm)


def subgrid(x, mm):
    assert T.trac2grid_fused_device(ctx, x, [(0, [cols[TR_TMP]], [5], [o_sg])], grid, gridmp, mm)


def separate(x, mm):
    T.trac2grid_device(ctx, x, node_cols, node_sch, grid, out6, mm)
    T.trac2grid_device(ctx, x, [cols[TR_ETA]], [6], gridmp, [o_n], mm)
    T.trac2grid_device(ctx, x, [cols[TR_HCD]], [5], [gridmp[IZ], grid[IX]], [o_kz], mm)
    T.trac2grid_device(ctx, x, [cols[TR_HCD]], [5], [grid[IZ], gridmp[IX]], [o_kx], mm)


res = {"ncell": ncell, "markers": M, "reps": reps}
zn, xn = tr_x[:, 0] / L[0], tr_x[:, 1] / L[1]
disp = torch.stack([torch.sin(np.pi * zn) * torch.cos(np.pi * xn), -torch.cos(np.pi * zn) * torch.sin(np.pi * xn)], 1)
moved = (tr_x + 3.3 * dz * disp).clamp_(2.0 ** -10, L[0] - 2.0 ** -10).contiguous()
del disp, zn, xn
ms_sort = timed(lambda: markers.sort_by_cell(moved, [cols[TR_TMP]], nx, L))
res["sort_x_plus_1_column_ms"] = ms_sort
resorted = markers.sort_by_cell(moved, [], nx, L)[0]
for name, x in (("ordered", tr_x), ("displaced", moved), ("displaced_resorted", resorted)):
    mm = T.marker_minmax(x, ctx)
    r = {}
    for nm, nfmax, parts, minb in ((960, 6, 1, 2), (960, 6, 1, 3), (960, 3, 1, 3), (1024, 6, 1, 2), (960, 2, 1, 3)):
        ctx.set_param("t2g_parts", parts), ctx.set_param("t2g_nm", nm), ctx.set_param("t2g_nfmax", nfmax)
        ctx.set_param("t2g_minb", minb)
        tag = "nm%d_nf%d_p%d_b%d" % (nm, nfmax, parts, minb)
        ms = timed(lambda: fused(x, mm))
        r["fused4_%s_ms" % tag] = round(ms, 3)
        r["fused4_%s_GBps(72B/marker)" % tag] = round(72 * M / ms / 1e6)
        r["subgrid_%s_ms" % tag] = round(timed(lambda: subgrid(x, mm)), 3)
    ctx.set_param("t2g_nm", 0), ctx.set_param("t2g_nfmax", 0), ctx.set_param("t2g_minb", 0)
    ctx.set_param("t2g_parts", 0)
    r["separate4_ms"] = timed(lambda: separate(x, mm))
    res["t2g_" + name] = r
# RK4 in a convection-roll velocity field
zc = torch.linspace(-0.5 * dz, L[0] + 0.5 * dz, nx[0] + 1, dtype=torch.float64, device="cuda").view(-1, 1) / L[0]
xc = torch.linspace(-0.5 * dz, L[1] + 0.5 * dz, nx[1] + 1, dtype=torch.float64, device="cuda").view(1, -1) / L[1]
vz = (1e-9 * torch.sin(np.pi * zc) * torch.cos(np.pi * xc)).contiguous()
vx = (-1e-9 * torch.cos(np.pi * zc) * torch.sin(np.pi * xc)).contiguous()
pre = [gridmp[d][0] - (gridmp[d][1] - gridmp[d][0]) for d in range(2)]
newgrid = [np.insert(gridmp[IZ], 0, pre[IZ]), np.insert(gridmp[IX], 0, pre[IX])]
dt = 0.67 * dz / 1e-9
for name, x in (("ordered", tr_x), ("displaced", moved)):
    ms = timed(lambda: T.rk4_device(ctx, x, newgrid, vz, vx, [nx[0] + 1, nx[1] + 1], dt))
    res["rk4_" + name] = {"ms": ms, "GBps": 48 * M / ms / 1e6}
    ms_f = timed(lambda: T.rk4_fence_count_device(ctx, x, newgrid, vz, vx, [nx[0] + 1, nx[1] + 1], dt, nx, L, 2.0 ** -10,
                                                   want_kelem=False))
    xq = x.clone()
    ms_s = timed(lambda: markers.fence_count(xq, nx, L, want_kelem=False))
    res["rk4_fence_count_" + name] = {"fused_ms": ms_f, "separate_ms": ms + ms_s}
print(json.dumps(res))
