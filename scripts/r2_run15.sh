#!/bin/bash
# delete_outside kernels + a fresh ncu capture of the marker kernels on the bench's own (advected, re-sorted) cloud
python -m pytest -m gpu tests/test_markers_gpu.py tests/test_flowthru_gpu.py tests/test_driver_gpu.py -q --timeout 1200 2>&1 | tail -6 | cut -c1-300
B="python bench.py --steps 1 --warmup 1 --e2e-steps 0 --cpu-ncell 0"
timeout 600 $B > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_t2g_fused|k_rk4|k_subgrid_fused' --launch-skip 36 --launch-count 5 \
  -o gpurun_out/r2_markers_bench -f $B > gpurun_out/r2_ncu15.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu15.log
