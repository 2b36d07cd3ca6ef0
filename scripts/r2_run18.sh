#!/bin/bash
# Stokes finest-level kernels under ncu (prof_stokes.py ran clean without ncu in run 17); only the CSV comes back
timeout 900 ncu --set full --clock-control none -k regex:'k_stokes_op_tile|k_precond_rhs|k_vel_op|k_restrict|k_prolong_add' --launch-skip 120 --launch-count 60 \
  -o /tmp/r2_stokes_kernels -f python scripts/prof_stokes.py 4096 > gpurun_out/r2_ncu18.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2_stokes_kernels.ncu-rep --page raw --csv > gpurun_out/r2_ncu_stokes_kernels_4096.csv 2>/dev/null
ls -la gpurun_out/ /tmp/r2_stokes_kernels.ncu-rep
tail -2 gpurun_out/r2_ncu18.log
