#!/bin/bash
# which ingredient makes the 2-GPU run need 8 iterations: ownership mode? the extrapolated warm start?
run() { tag=$1; shift; "$@" > gpurun_out/diag2_$tag.json 2> gpurun_out/diag2_$tag.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/diag2_$tag.json').read().splitlines() if l.startswith('{')][-1])
print('$tag', [i['stokes_iters'] for i in d['solver_iterations']], round(d['phases_ms_per_step']['stokes_solve'],1))
PY
}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
C="--steps 6 --warmup 3 --cpu-ncell 0 --e2e-steps 0"
run n2_index   $T 29521 bench.py --gpus 2 $C --marker-ownership index
run n2_ws1     $T 29522 bench.py --gpus 2 $C --warm-start 1
run n1_ws1     python bench.py $C --warm-start 1
run n2_slabnl  $T 29523 bench.py --gpus 2 $C --slab-local 0
