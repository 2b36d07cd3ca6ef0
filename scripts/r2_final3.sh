#!/bin/bash
# 2 GPUs, the bench line as the driver launches it (with the e2e leg on slab-owned markers)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-ncell 0 > gpurun_out/r02_bench_2gpu_4096_final.json 2> gpurun_out/r02_bench_2gpu_final.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_2gpu_4096_final.json').read().splitlines() if l.startswith('{')][-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()})
print(d.get('e2e'), d['config'].get('parallelism'))
PY
tail -5 gpurun_out/r02_bench_2gpu_final.err | cut -c1-300
