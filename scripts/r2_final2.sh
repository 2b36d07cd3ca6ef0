#!/bin/bash
# round-2 final state, 2 GPUs: the multi-GPU parity tests (index / slab / slab-reduce / slab-local + native migration)
timeout 1200 python -m pytest -m gpu tests/test_multi_gpu.py -q --timeout 900 -rs > gpurun_out/r02_multi_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_multi_gpu_final.log | cut -c1-300
