#!/bin/bash
# round 2: everything that had never run on a GPU (surfstab, slabgrid halves, slab-owned markers + migration)
export PLB_RUN_UNVERIFIED=1
python -m pytest -m gpu tests/test_surfstab_gpu.py tests/test_slabgrid_gpu.py -q --timeout 600 2>&1 | tail -40 > gpurun_out/r2_unverified.log
NG=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$NG" -ge 2 ]; then
  python -m pytest -m gpu tests/test_multi_gpu.py -q --timeout 900 2>&1 | tail -60 > gpurun_out/r2_multi_gpu.log
fi
tail -25 gpurun_out/r2_unverified.log; tail -40 gpurun_out/r2_multi_gpu.log 2>/dev/null
