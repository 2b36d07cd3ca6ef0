#!/bin/bash
python -m pytest -m gpu tests/test_markers_gpu.py tests/test_launcher_gpu.py -q -x --timeout 900 2>&1 | tail -15 > gpurun_out/r2_pytest4.log
tail -8 gpurun_out/r2_pytest4.log
python scripts/bench_markers2.py 2048 5 > gpurun_out/r2_bench_markers2_2048.json 2> gpurun_out/r2_bench_markers2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_markers2_2048.json'))
for k,v in d.items(): print(k, v)"
tail -3 gpurun_out/r2_bench_markers2.err
python scripts/prof_t2g_fused.py 4096 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_t2g_fused -s 2 -c 2 -o gpurun_out/r02_prof_t2g_fused -f python scripts/prof_t2g_fused.py 4096 > gpurun_out/ncu_t2g_fused.log 2>&1
tail -3 gpurun_out/ncu_t2g_fused.log
