"""RK4 (+ fence + count) and the two marker-temperature kernels of one time step at the benchmark size, twice each, for
`ncu --set full -k regex:"k_rk4|k_subgrid_fused"`.
  python scripts/prof_markers.py [ncell=4096]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylamp_b200 import _lib, markers, pylamp_trac as T, setups  # noqa: E402
from pylamp_b200.pylamp_const import IX, IZ, TR_HCD, TR_HCP, TR_RHO, TR_TMP  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = _lib.default_context(0)
nx, L, tr_x, cols, opts = setups.convection_device(ncell=ncell, per_side=4, device="cuda:0")
grid, gridmp = setups.make_grids(nx, L)
cols[TR_RHO].fill_(3300.0)
dz = L[0] / ncell
zc = torch.linspace(-0.5 * dz, L[0] + 0.5 * dz, nx[0] + 1, dtype=torch.float64, device="cuda").view(-1, 1) / L[0]
xc = torch.linspace(-0.5 * dz, L[1] + 0.5 * dz, nx[1] + 1, dtype=torch.float64, device="cuda").view(1, -1) / L[1]
vz = (1e-9 * torch.sin(np.pi * zc) * torch.cos(np.pi * xc)).contiguous()
vx = (-1e-9 * torch.cos(np.pi * zc) * torch.sin(np.pi * xc)).contiguous()
pre = [gridmp[d][0] - (gridmp[d][1] - gridmp[d][0]) for d in range(2)]
newgrid = [np.insert(gridmp[IZ], 0, pre[IZ]), np.insert(gridmp[IX], 0, pre[IX])]
dt = 0.67 * dz / 1e-9
dTg = torch.rand(tuple(nx), dtype=torch.float64, device="cuda")
for _ in range(2):
    T.rk4_fence_count_device(ctx, tr_x, newgrid, vz, vx, [nx[0] + 1, nx[1] + 1], dt, nx, L, 2.0 ** -10, want_kelem=False)
    Tsg, dT = markers.subgrid_fused(1, tr_x, grid, dTg, cols[TR_TMP], 1e10, dz, dz, cols[TR_HCP], cols[TR_RHO], cols[TR_HCD])
    markers.subgrid_fused(2, tr_x, grid, dTg, cols[TR_TMP], Tsg=Tsg)
torch.cuda.synchronize()
print("done", tr_x.shape[0])
