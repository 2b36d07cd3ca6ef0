#!/bin/bash
# fused trac2grid under ncu on the bench's own cloud (the bench command ran clean without ncu in run 19); CSVs only
B="python bench.py --steps 1 --warmup 1 --spinup 4 --e2e-steps 0 --cpu-ncell 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_t2g_fused' --launch-skip 4 --launch-count 1 \
  -o /tmp/r2_t2g_fused -f $B > gpurun_out/r2_ncu20.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r2_t2g_fused.ncu-rep --page raw --csv > gpurun_out/r2_ncu_k_t2g_fused_bench.csv 2>/dev/null
ncu -i /tmp/r2_t2g_fused.ncu-rep --page source --csv > gpurun_out/r2_ncu_k_t2g_fused_bench_source.csv 2>/dev/null
ls -la gpurun_out/
grep -v "^{" gpurun_out/r2_ncu20.log | tail -3
