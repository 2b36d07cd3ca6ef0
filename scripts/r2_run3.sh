#!/bin/bash
python -m pytest -m gpu tests/test_markers_gpu.py tests/test_driver_gpu.py tests/test_stokes_large_gpu.py -q -x --timeout 900 2>&1 | tail -30 > gpurun_out/r2_pytest3.log
tail -15 gpurun_out/r2_pytest3.log
python bench.py --cpu-ncell 0 --e2e-steps 0 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench3.json'))
print('value',d['value'],'ms',d['ms_per_step']); print(d['phases_ms_per_step']); print({k:(round(v['ms_per_step'],2),v['GBps'] and round(v['GBps'])) for k,v in d['kernel_breakdown'].items()}); print(d['solver_iterations'][-3:]); print(d['roofline'])
PY
tail -5 gpurun_out/r2_bench3.err
