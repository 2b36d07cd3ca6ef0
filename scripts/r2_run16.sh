#!/bin/bash
# marker kernels after the instruction-count pass (fused t2g addressing, RK4/subgrid FMAs + prefetch)
python -m pytest -m gpu tests/test_markers_gpu.py tests/test_driver_gpu.py tests/test_dropin_loop_gpu.py tests/test_launcher_gpu.py -q --timeout 1200 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-ncell 0 --e2e-steps 0 > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench16.json').read().splitlines() if l.startswith('{')][-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()})
print({k[:12]:(round(v['ms_per_step'],2),v['GBps'] and round(v['GBps'])) for k,v in d['kernel_breakdown'].items()})
print(d['ms_of_each_timed_step'], d['roofline']['frac'])
PY
tail -3 gpurun_out/r2_bench16.err
python scripts/bench_markers2.py 2048 2>&1 | tail -12
