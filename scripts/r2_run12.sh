#!/bin/bash
python -m pytest -m gpu tests/test_markers_gpu.py tests/test_driver_gpu.py tests/test_flowthru_gpu.py tests/test_dropin_loop_gpu.py -q --timeout 1200 2>&1 | tail -15 > gpurun_out/r2_pytest12.log
tail -12 gpurun_out/r2_pytest12.log | cut -c1-300
python scripts/bench_markers2.py 2048 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
print({k:v for k,v in d.items() if k.startswith('rk4')})"
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-ncell 0 --e2e-steps 0 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench12.json'))
print('value',d['value'],'ms',d['ms_per_step']); print(d['phases_ms_per_step']); print({k:(round(v['ms_per_step'],2),v['GBps'] and round(v['GBps'])) for k,v in d['kernel_breakdown'].items()}); print(d['solver_iterations'][-2:]); print(d['roofline'])
PY
tail -3 gpurun_out/r2_bench12.err
