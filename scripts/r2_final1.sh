#!/bin/bash
# round-2 final state, 1 GPU: smoke, the whole GPU test suite, the bench line as the driver runs it, the launch list
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log | cut -c1-300
python -m pytest -m gpu tests -q --timeout 1500 > gpurun_out/r02_gpu_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gpu_pytest_final.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_4096_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_1gpu_4096_final.json').read().splitlines() if l.startswith('{')][-1])
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phases_ms_per_step'].items()})
print({k[:12]:(round(v['ms_per_step'],2),v['GBps'] and round(v['GBps'])) for k,v in d['kernel_breakdown'].items()})
print(d['roofline']['frac'], d['e2e'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])
PY
tail -2 gpurun_out/r02_bench_final.err
L="python bench.py --steps 2 --warmup 1 --spinup 2 --e2e-steps 0 --cpu-ncell 0"
timeout 600 $L > gpurun_out/r02_launchlist_plain.json 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_ncu_launches_bench_4096.csv $L > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -8
