#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/multi_gpu_driver_check.py 64 4 slab local > gpurun_out/r2_mgpu_local.log 2>&1
grep -E "step|PARITY|Error|error|rank" gpurun_out/r2_mgpu_local.log | cut -c1-500 | tail -20
python -m pytest -m gpu tests/test_markers_gpu.py tests/test_driver_gpu.py tests/test_stokes_gpu.py tests/test_diff_gpu.py -q -x --timeout 900 2>&1 | tail -3
