#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 scripts/multi_gpu_driver_check.py 64 6 slab local > gpurun_out/r2_mgpu_local_native.log 2>&1
grep -E "step|PARITY|Error|error|rank|Traceback" gpurun_out/r2_mgpu_local_native.log | cut -c1-400 | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
    print('N=2 value',d['value'],'ms',d['ms_per_step']); print(d['phases_ms_per_step']); print(d['solver_iterations'][-2:]); print(d['config']['parallelism'])
except Exception as e: print('bench failed', e)
PY
tail -5 gpurun_out/r2_bench_2gpu.err
