#!/bin/bash
# RK4 and the marker-temperature kernels after the round's changes, on the bench's own cloud; CSV only
B="python bench.py --steps 1 --warmup 1 --spinup 4 --e2e-steps 0 --cpu-ncell 0"
timeout 900 ncu --set full --clock-control none -k regex:'k_rk4|k_subgrid_fused' --launch-skip 12 --launch-count 3 \
  -o /tmp/r2_markers_after -f $B > gpurun_out/r2_ncu21.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2_markers_after.ncu-rep --page raw --csv > gpurun_out/r02_ncu_k_rk4_subgrid_after_4096.csv 2>/dev/null
ls -la gpurun_out/
grep -v "^{" gpurun_out/r2_ncu21.log | tail -3
