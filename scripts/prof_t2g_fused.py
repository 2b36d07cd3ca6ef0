"""The fused trac2grid launches of one timestep at the benchmark size (main call: 4 targets / 7 columns;
subgrid call: nodes / 1 column), once each, for `ncu --set full -k regex:k_t2g_fused`.
  python scripts/prof_t2g_fused.py [ncell=4096]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylamp_b200 import _lib, pylamp_trac as T, setups  # noqa: E402
from pylamp_b200.pylamp_const import TR_ETA, TR_HCD, TR_HCP, TR_IHT, TR_MAT, TR_RHO, TR_TMP  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = _lib.default_context(0)
nx, L, tr_x, cols, opts = setups.convection_device(ncell=ncell, per_side=4, device="cuda:0")
grid, gridmp = setups.make_grids(nx, L)
cols[TR_RHO].fill_(3300.0)
cols[TR_ETA].copy_(1e20 * (1.0 + tr_x[:, 0] / L[0]))
new = lambda: torch.empty(tuple(nx), dtype=torch.float64, device="cuda")
out = [new() for _ in range(10)]
mm = T.marker_minmax(tr_x, ctx)
for _ in range(2):
    assert T.trac2grid_fused_device(ctx, tr_x, [(0, [cols[k] for k in (TR_RHO, TR_ETA, TR_HCP, TR_TMP, TR_IHT, TR_MAT)],
                                                 [5, 6, 5, 5, 5, 5], out[:6]), (1, [cols[TR_ETA]], [6], [out[6]]),
                                                (2, [cols[TR_HCD]], [5], [out[7]]), (3, [cols[TR_HCD]], [5], [out[8]])],
                                    grid, gridmp, mm)
    assert T.trac2grid_fused_device(ctx, tr_x, [(0, [cols[TR_TMP]], [5], [out[9]])], grid, gridmp, mm)
torch.cuda.synchronize()
print("done", tr_x.shape[0])
