#!/bin/bash
# extrapolation order of the Stokes warm start: iterations and solve time, N GPUs (N = $1)
N=$1
for ws in 3 4 6; do
if [ "$N" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$ws bench.py --gpus $N"; fi
timeout 300 $CMD --steps 10 --warmup 3 --cpu-ncell 0 --e2e-steps 0 --warm-start $ws > gpurun_out/ws_${N}_$ws.json 2> gpurun_out/ws_${N}_$ws.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/ws_${N}_$ws.json').read().splitlines() if l.startswith('{')][-1])
print('N=$N ws=$ws', [i['stokes_iters'] for i in d['solver_iterations']], 'stokes', round(d['phases_ms_per_step']['stokes_solve'],1), 'step', round(d['ms_per_step'],1))
PY
done
