#!/bin/bash
python -m pytest -m gpu tests/test_markers_gpu.py tests/test_stokes_large_gpu.py -q -s --timeout 900 -k "fused or sort or large" 2>&1 | grep -E "passed|failed|iters|Error|error" | cut -c1-700 > gpurun_out/r2_pytest6.log
cat gpurun_out/r2_pytest6.log
python scripts/bench_markers2.py 2048 5 > gpurun_out/r2_bench_markers2_2048.json 2> gpurun_out/r2_bench_markers2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_markers2_2048.json'))
for k,v in d.items(): print(k, v)"
tail -3 gpurun_out/r2_bench_markers2.err
