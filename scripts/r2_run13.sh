#!/bin/bash
python -m pytest -m gpu tests/test_markers_gpu.py tests/test_driver_gpu.py -q --timeout 1200 2>&1 | tail -3
python scripts/bench_markers2.py 2048 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
print({k:v for k,v in d.items() if k.startswith('rk4')})"
for f in 0 1; do
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-ncell 0 --e2e-steps 0 --fused-rk4-fence $f > gpurun_out/r2_bench13_$f.json 2> gpurun_out/r2_bench13.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench13_$f.json'))
print('fused=$f value',d['value'],'ms',d['ms_per_step']); print(d['phases_ms_per_step'])
PY
done
tail -3 gpurun_out/r2_bench13.err
