#!/bin/bash
python scripts/solver_floor_study.py 512 12 1e-9 1e-10 1e-11 1e-12 > gpurun_out/r2_floor_512.log 2>&1
python scripts/solver_floor_study.py 4096 12 1e-10 1e-11 > gpurun_out/r2_floor_4096.log 2>&1
python -m pytest -m gpu tests/test_stokes_large_gpu.py -q -s --timeout 900 2>&1 | tail -30 > gpurun_out/r2_pytest_large.log
grep -v "fgmres it" gpurun_out/r2_floor_512.log | tail -30; grep -v "fgmres it" gpurun_out/r2_floor_4096.log | tail -30; tail -20 gpurun_out/r2_pytest_large.log
