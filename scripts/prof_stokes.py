"""One Stokes solve on the analytic C4-type fields at the benchmark size with bench.py's solver settings, for
  ncu --set full -k regex:'k_stokes_op_tile|k_precond_rhs|k_vel_op|k_restrict|k_prolong_add' --launch-skip 120 --launch-count 60
(the finest-level launches are the ones with the largest grids).
  python scripts/prof_stokes.py [ncell=4096]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pylamp_b200 import pylamp_stokes as S, setups  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nx, L, grid, gridmp, es, en, rho = setups.convection_fields(ncell, 0)
f = [torch.as_tensor(a).to("cuda") for a in (es, en, rho)]
A = S.StokesOperator(nx, grid, *f, [1, 1, 1, 1])
A.warn_unconverged = False
for k, v in bench.stokes_params().items():
    A.set_param(k, v)
A.solve(None, rtol=1e-9, maxit=60, raise_on_fail=False)
torch.cuda.synchronize()
print("done", A.stats)
