"""The five trac2grid calls of one timestep (pylamp2.py:309-313 + the subgrid call :478) at the
benchmark size, once, for `ncu --set full -k regex:k_t2g_chunk`: DRAM bytes per launch of the
dominant kernel class (profiles/traffic.json, bench.py's roofline.traffic).
  python scripts/prof_t2g.py [ncell=4096]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylamp_b200 import _lib, pylamp_trac as T, setups  # noqa: E402
from pylamp_b200.pylamp_const import IX, IZ, TR_ETA, TR_HCD, TR_HCP, TR_IHT, TR_MAT, TR_RHO, TR_TMP  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = _lib.default_context(0)
nx, L, tr_x, cols, opts = setups.convection_device(ncell=ncell, per_side=4, device="cuda:0")
grid, gridmp = setups.make_grids(nx, L)
cols[TR_RHO].fill_(3300.0)
cols[TR_ETA].copy_(1e20 * (1.0 + tr_x[:, 0] / L[0]))
out = [torch.empty(tuple(nx), dtype=torch.float64, device="cuda") for _ in range(6)]
mm = T.marker_minmax(tr_x, ctx)
T.trac2grid_device(ctx, tr_x, [cols[k] for k in (TR_RHO, TR_ETA, TR_HCP, TR_TMP, TR_IHT, TR_MAT)],
                   [5, 6, 5, 5, 5, 5], grid, out, mm)
T.trac2grid_device(ctx, tr_x, [cols[TR_ETA]], [6], gridmp, out[:1], mm)
T.trac2grid_device(ctx, tr_x, [cols[TR_HCD]], [5], [gridmp[IZ], grid[IX]], out[:1], mm)
T.trac2grid_device(ctx, tr_x, [cols[TR_HCD]], [5], [grid[IZ], gridmp[IX]], out[:1], mm)
T.trac2grid_device(ctx, tr_x, [cols[TR_TMP]], [5], grid, out[:1], mm)
torch.cuda.synchronize()
print("done", tr_x.shape[0])
