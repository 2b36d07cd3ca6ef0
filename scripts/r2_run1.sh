#!/bin/bash
# full GPU suite on the new build + solver floor study
python -m pytest -m gpu tests -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2_pytest1.log
python scripts/solver_floor_study.py 512 12 1e-9 1e-10 1e-11 1e-12 > gpurun_out/r2_floor_512.log 2>&1
python scripts/solver_floor_study.py 4096 12 1e-10 1e-11 > gpurun_out/r2_floor_4096.log 2>&1
tail -5 gpurun_out/r2_pytest1.log; grep -v "fgmres it" gpurun_out/r2_floor_512.log | tail -30; grep -v "fgmres it" gpurun_out/r2_floor_4096.log | tail -30
