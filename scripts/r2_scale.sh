#!/bin/bash
# strong (4096^2) and weak (4096*sqrt(N) cells, rounded to a multigrid-friendly size) scaling points on N GPUs
N=$1; WEAK=$2; PORT=${3:-29571}
run() {  # name, extra args
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 0 --cpu-ncell 0 $2 > gpurun_out/r2_scale_$1_n$N.json 2> gpurun_out/r2_scale_$1_n$N.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_scale_$1_n$N.json').read().splitlines() if l.startswith('{')][-1])
    print('$1 N=$N ncell',d['config']['grid_nodes'][0]-1,'value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['solver_iterations'][-1])
except Exception as e:
    print('$1 N=$N FAILED', e); print(open('gpurun_out/r2_scale_$1_n$N.err').read()[-1500:])
PY
}
if [ "${4:-both}" != "weakonly" ]; then run strong ""; fi
# weak points: the fp64 floor of the scaled residual grows with the cell count (2.0e-10 at 4096^2, measured); keep the
# requested tolerance the same factor above it
if [ -n "$WEAK" ] && [ "$WEAK" != "0" ]; then
  RTOL=$(python -c "print('%.2e' % (1e-9 * ($WEAK / 4096.0) ** 2))")
  PORT=$((PORT+2)); run weak "--ncell $WEAK --scaling weak --stokes-rtol $RTOL"
fi
