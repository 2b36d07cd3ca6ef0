"""Marker-kernel micro-benchmark (CUDA events on the launching stream, inputs larger than L2):
trac2grid with the generic scatter kernel (t2g_variant 0) and the wide-load chunk kernel (1), on the
cell-ordered cloud and on the same cloud after advection-like displacement; RK4; fence + count
fused and separate.  Prints one JSON object.
  python scripts/bench_markers.py [ncell=2048] [reps=10]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylamp_b200 import _lib, markers, pylamp_trac as T, setups  # noqa: E402
from pylamp_b200.pylamp_const import IX, IZ, TR_ETA, TR_HCD, TR_HCP, TR_IHT, TR_MAT, TR_RHO, TR_TMP  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = _lib.default_context(0)
nx, L, tr_x, cols, opts = setups.convection_device(ncell=ncell, per_side=4, device="cuda:0")
grid, gridmp = setups.make_grids(nx, L)
M = tr_x.shape[0]
cols[TR_RHO].fill_(3300.0)
cols[TR_ETA].copy_(1e20 * (1.0 + tr_x[:, 0] / L[0]))
out6 = [torch.empty(tuple(nx), dtype=torch.float64, device="cuda") for _ in range(6)]
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


res = {"ncell": ncell, "markers": M, "reps": reps}
# a smooth, divergence-free-ish displacement of a few cells: what the cloud looks like after some steps
zn, xn = tr_x[:, 0] / L[0], tr_x[:, 1] / L[1]
disp = torch.stack([torch.sin(np.pi * zn) * torch.cos(np.pi * xn), -torch.cos(np.pi * zn) * torch.sin(np.pi * xn)], 1)
moved = (tr_x + 3.3 * dz * disp).clamp_(2.0 ** -10, L[0] - 2.0 ** -10).contiguous()
del disp, zn, xn
for name, x in (("ordered", tr_x), ("displaced", moved)):
    mm = T.marker_minmax(x, ctx)
    for variant in (0, 1, 2):
        ctx.set_param("t2g_variant", variant)
        nodes = lambda: T.trac2grid_device(ctx, x, [cols[k] for k in (TR_RHO, TR_ETA, TR_HCP, TR_TMP, TR_IHT, TR_MAT)],
                                           [5, 6, 5, 5, 5, 5], grid, out6, mm)
        centre = lambda: T.trac2grid_device(ctx, x, [cols[TR_ETA]], [6], gridmp, out6[:1], mm)
        stag = lambda: T.trac2grid_device(ctx, x, [cols[TR_HCD]], [5], [gridmp[IZ], grid[IX]], out6[:1], mm)
        ms6, ms1g, ms1a = timed(nodes), timed(centre), timed(stag)
        res["t2g_%s_variant%d" % (name, variant)] = {
            "k6_nodes_ms": ms6, "k6_GBps": (16 + 48) * M / ms6 / 1e6, "k1_geom_centres_ms": ms1g,
            "k1_geom_GBps": 24 * M / ms1g / 1e6, "k1_arith_staggered_ms": ms1a, "k1_arith_GBps": 24 * M / ms1a / 1e6,
            "per_step_ms(k6 + k1g + 3 k1a)": ms6 + ms1g + 3 * ms1a}
    ctx.set_param("t2g_variant", 1)
# RK4 in a convection-roll velocity field
zc = torch.linspace(-0.5 * dz, L[0] + 0.5 * dz, nx[0] + 1, dtype=torch.float64, device="cuda").view(-1, 1) / L[0]
xc = torch.linspace(-0.5 * dz, L[1] + 0.5 * dz, nx[1] + 1, dtype=torch.float64, device="cuda").view(1, -1) / L[1]
vz = (1e-9 * torch.sin(np.pi * zc) * torch.cos(np.pi * xc)).contiguous()
vx = (-1e-9 * torch.cos(np.pi * zc) * torch.sin(np.pi * xc)).contiguous()
pre = [gridmp[d][0] - (gridmp[d][1] - gridmp[d][0]) for d in range(2)]
newgrid = [np.insert(gridmp[IZ], 0, pre[IZ]), np.insert(gridmp[IX], 0, pre[IX])]
dt = 0.67 * dz / 1e-9
ms = timed(lambda: T.rk4_device(ctx, tr_x, newgrid, vz, vx, [nx[0] + 1, nx[1] + 1], dt))
res["rk4"] = {"ms": ms, "GBps": 48 * M / ms / 1e6}
x2 = moved.clone()
ms_sep = timed(lambda: (markers.fence(x2, L), markers.cell_index_count(x2, nx, L, want_kelem=False)))
ms_fus = timed(lambda: markers.fence_count(x2, nx, L, want_kelem=False))
res["fence_count"] = {"separate_ms": ms_sep, "fused_ms": ms_fus}
print(json.dumps(res))
