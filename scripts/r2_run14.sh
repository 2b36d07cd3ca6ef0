#!/bin/bash
python -m pytest -m gpu tests/test_diff_gpu.py tests/test_driver_gpu.py tests/test_dropin_loop_gpu.py tests/test_fullsize_gpu.py tests/test_flowthru_gpu.py -q --timeout 1200 2>&1 | tail -6 | cut -c1-300
for g in 1 0; do
PLB_DIFF_GMRES=$g timeout 600 python bench.py --steps 10 --warmup 3 --cpu-ncell 0 --e2e-steps 0 > gpurun_out/r2_bench14_$g.json 2> gpurun_out/r2_bench14.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench14_$g.json').read().splitlines() if l.startswith('{')][-1])
print('gmres=$g value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['solver_iterations'][-1]['heat_iters'])
PY
done
tail -3 gpurun_out/r2_bench14.err
