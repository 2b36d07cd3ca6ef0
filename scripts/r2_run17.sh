#!/bin/bash
# default bench (device leg with the sort's buffers warm, e2e leg with the side-stream download, cpu leg) + Stokes kernel capture
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench17.json').read().splitlines() if l.startswith('{')][-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()})
print({k[:12]:(round(v['ms_per_step'],2),v['GBps'] and round(v['GBps'])) for k,v in d['kernel_breakdown'].items()})
print(d['ms_of_each_timed_step'], d['roofline']['frac'], d['e2e'], d['cpu_baseline']['value'])
PY
tail -3 gpurun_out/r2_bench17.err
python scripts/prof_stokes.py 4096 > gpurun_out/r2_prof_stokes_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r2_prof_stokes_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stokes_op_tile|k_precond_rhs|k_vel_op|k_restrict|k_prolong_add' --launch-skip 120 --launch-count 60 \
  -o gpurun_out/r2_stokes_kernels -f python scripts/prof_stokes.py 4096 > gpurun_out/r2_ncu17.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2_ncu17.log
