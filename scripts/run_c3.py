"""BASELINE.json configs[2] (C3) as named: 2-D Rayleigh-Taylor instability, temperature-independent Newtonian
viscosity (1e21 over 1e20, density 3300 over 3200, cosine-perturbed interface), 1024^2 cells, 16 markers/cell,
100 time steps (Stokes + RK4 advection: the reference's Stokes-only mode, pylamp2.py:315-319).  Logs per-step
phase times (CUDA events) so that the decay of the marker order between re-sorts is visible, and the Stokes
iteration counts.  Prints one JSON object.
  python scripts/run_c3.py [ncell=1024] [nsteps=100] [resort_every=16]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pylamp_b200 import driver, setups  # noqa: E402

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
resort = int(sys.argv[3]) if len(sys.argv) > 3 else 16
nx, L, tr_x, tr_f, opts = setups.rayleigh_taylor(ncell=ncell, per_side=4)
s = driver.State(nx, L, tr_x, tr_f, device=0)
del tr_x, tr_f
o = driver.Options(**opts)
o.stokes_rtol = bench.DEFAULTS["stokes_rtol"]
o.stokes_params = bench.stokes_params()
o.resort_every = resort
M0 = s.ntrac
log = []
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(nsteps):
    driver.timestep(s, o, want_kelem=False, phases=True)
    ph = s.phases.result()
    log.append({"step": it + 1, "ms": round(sum(ph.values()), 3), "trac2grid": round(ph.get("trac2grid", 0), 3),
                "stokes": round(ph.get("stokes_solve", 0), 3), "rk4": round(ph.get("advect_rk4", 0), 3),
                "fence_count": round(ph.get("fence_count", 0), 3), "iters": s.stats["stokes_iters"],
                "relres": s.stats["stokes_relres"], "status": s.stats["stokes_status"], "dt": s.tstep})
torch.cuda.synchronize()
wall = time.perf_counter() - t0
mean = lambda rows, k: float(np.mean([r[k] for r in rows]))
first, last = log[:10], log[-10:]
vmax = float(max(s.newvel[0].abs().max().item(), s.newvel[1].abs().max().item()))
out = {"config": "C3 Rayleigh-Taylor, %d^2 cells, 16 markers/cell, %d steps, re-sort every %d steps" % (ncell, nsteps, resort),
       "markers": M0, "markers_end": s.ntrac, "wall_s": wall, "timesteps_per_s": nsteps / wall,
       "model_time_s": s.totaltime, "max_velocity_end": vmax,
       "all_solves_converged": all(r["status"] == "converged" for r in log),
       "first10_vs_last10_ms": {k: [round(mean(first, k), 3), round(mean(last, k), 3)]
                                for k in ("ms", "trac2grid", "stokes", "rk4", "fence_count", "iters")},
       "per_step": log}
print(json.dumps(out))
