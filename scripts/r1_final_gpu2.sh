#!/bin/bash
# Round-1 last GPU call: parity of the trac2grid variants, their A/B timing, ncu launch list of the bench command.
set +e
O=gpurun_out/f2
mkdir -p $O
timeout 60 python -m pytest tests/test_markers_gpu.py -m gpu -q -k "trac2grid or chunk or fence_count or large" > $O/pytest_markers.log 2>&1
echo "pytest_markers rc=$?" >> $O/status.txt
timeout 50 python scripts/bench_markers.py 2048 10 > $O/bench_markers.json 2> $O/bench_markers.err
echo "bench_markers rc=$?" >> $O/status.txt
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 2 --spinup 0 --e2e-steps 0 --cpu-ncell 0 > $O/ncu_bench.log 2>&1
echo "ncu rc=$?" >> $O/status.txt
echo done
