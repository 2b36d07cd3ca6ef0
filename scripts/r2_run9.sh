#!/bin/bash
python -m pytest -m gpu tests -q --timeout 1200 2>&1 | tail -12 > gpurun_out/r2_pytest9.log
tail -6 gpurun_out/r2_pytest9.log
python -m pytest -m gpu tests/test_stokes_large_gpu.py -q -s --timeout 1200 2>&1 | grep -E "SolCx|order|iters|passed|failed" | cut -c1-400 > gpurun_out/r2_pytest_large.log
cat gpurun_out/r2_pytest_large.log
python scripts/run_c3.py 1024 100 16 > gpurun_out/r2_c3_rt_1024.json 2> gpurun_out/r2_c3.err
python -c "
import json; d=json.load(open('gpurun_out/r2_c3_rt_1024.json')); d.pop('per_step'); print(d)"
tail -3 gpurun_out/r2_c3.err
python bench.py > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench9.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d.get('e2e')); print(d['phases_ms_per_step']); print({k:(round(v['ms_per_step'],2),v['GBps'] and round(v['GBps'])) for k,v in d['kernel_breakdown'].items()}); print(d['solver_iterations'][-2:]); print(d['roofline']); print(d['config']); print(d.get('cpu_baseline',{}).get('value'))
PY
tail -5 gpurun_out/r2_bench9.err
