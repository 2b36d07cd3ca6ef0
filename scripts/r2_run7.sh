#!/bin/bash
python -m pytest -m gpu tests/test_markers_gpu.py -q --timeout 900 -k "fused" 2>&1 | tail -3
python scripts/bench_markers2.py 2048 5 > gpurun_out/r2_bench_markers2_2048.json 2> gpurun_out/r2_bench_markers2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_markers2_2048.json'))
for k,v in d.items(): print(k, v)"
tail -3 gpurun_out/r2_bench_markers2.err
