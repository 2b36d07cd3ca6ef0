#!/bin/bash
# Round-1 final GPU call: parity of everything touched, marker-kernel A/B, bench, remaining tests, ncu of the new kernel.
set +e
O=gpurun_out/f
mkdir -p $O
python -m pytest tests/test_markers_gpu.py tests/test_fullsize_gpu.py tests/test_driver_gpu.py -m gpu -q -s > $O/pytest_new.log 2>&1
echo "pytest_new rc=$?" >> $O/status.txt
python scripts/bench_markers.py 2048 10 > $O/bench_markers.json 2> $O/bench_markers.err
echo "bench_markers rc=$?" >> $O/status.txt
python bench.py --cpu-ncell 0 --e2e-steps 0 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/status.txt
python -m pytest tests/test_stokes_gpu.py tests/test_diff_gpu.py tests/test_dropin_loop_gpu.py tests/test_multi_gpu.py -m gpu -q > $O/pytest_rest.log 2>&1
echo "pytest_rest rc=$?" >> $O/status.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_t2g_chunk -c 5 -o $O/prof_t2g_chunk python scripts/prof_t2g.py 4096 > $O/ncu_t2g.log 2>&1
echo "ncu rc=$?" >> $O/status.txt
echo done
