#!/bin/bash
# 8 GPUs: the weak-scaling point (11264^2, tolerance scaled with the fp64 floor) and strong scaling at 8192^2 (4 -> 8 GPUs)
run() {
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-steps 0 --cpu-ncell 0 $2 > gpurun_out/r2_scale_$1_n8.json 2> gpurun_out/r2_scale_$1_n8.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_scale_$1_n8.json').read().splitlines() if l.startswith('{')][-1])
    print('$1 N=8 ncell',d['config']['grid_nodes'][0]-1,'value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['solver_iterations'][-1], d['ms_of_each_timed_step'])
except Exception as e:
    print('$1 FAILED', e); print(open('gpurun_out/r2_scale_$1_n8.err').read()[-1500:])
PY
}
run weak "--ncell 11264 --scaling weak --stokes-rtol 7.6e-9" 29581
run strong8192 "--ncell 8192 --scaling strong --stokes-rtol 4e-9" 29583
