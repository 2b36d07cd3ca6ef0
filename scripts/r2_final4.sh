#!/bin/bash
# 4 GPUs, the bench line as the driver launches it (final code, with the e2e leg)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 --cpu-ncell 0 > gpurun_out/r02_bench_4gpu_4096_final.json 2> gpurun_out/r02_bench_4gpu_final.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_4gpu_4096_final.json').read().splitlines() if l.startswith('{')][-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()})
print(d.get('e2e'), [i['stokes_iters'] for i in d['solver_iterations']])
PY
tail -3 gpurun_out/r02_bench_4gpu_final.err | cut -c1-300
