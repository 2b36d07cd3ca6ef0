#!/bin/bash
python -m pytest -m gpu tests -q -x --timeout 1200 --deselect tests/test_stokes_large_gpu.py 2>&1 | tail -12 > gpurun_out/r2_pytest11.log
tail -8 gpurun_out/r2_pytest11.log
python -m pytest -m gpu tests/test_flowthru_gpu.py -q -s 2>&1 | grep -E "flow-through|step|passed|failed" | cut -c1-500
python scripts/bench_markers2.py 2048 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
print({k:v for k,v in d.items() if k.startswith('rk4')})"
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-ncell 0 --e2e-steps 0 > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench11.json'))
print('value',d['value'],'ms',d['ms_per_step']); print(d['phases_ms_per_step']); print({k:(round(v['ms_per_step'],2),v['GBps'] and round(v['GBps'])) for k,v in d['kernel_breakdown'].items()}); print(d['solver_iterations'][-2:]); print(d['roofline'])
PY
tail -3 gpurun_out/r2_bench11.err
