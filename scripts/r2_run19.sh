#!/bin/bash
# preconditioner scale planes: solver parity tests + bench
python -m pytest -m gpu tests/test_stokes_gpu.py tests/test_stokes_large_gpu.py tests/test_surfstab_gpu.py tests/test_flowthru_gpu.py tests/test_fullsize_gpu.py -q --timeout 1200 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-ncell 0 --e2e-steps 0 > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench19.json').read().splitlines() if l.startswith('{')][-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()})
print({k[:12]:(round(v['ms_per_step'],2),v['GBps'] and round(v['GBps'])) for k,v in d['kernel_breakdown'].items()})
print(d['ms_of_each_timed_step'], [i['stokes_iters'] for i in d['solver_iterations']])
PY
tail -3 gpurun_out/r2_bench19.err
