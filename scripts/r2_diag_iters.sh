#!/bin/bash
# residual histories of the timed Stokes solves: 1 GPU vs 2 GPUs (why does the 2-GPU run need 8 iterations where 1 GPU needs 6-7?)
export PLB_DEBUG_FGMRES=1
timeout 300 python bench.py --steps 4 --warmup 3 --cpu-ncell 0 --e2e-steps 0 > gpurun_out/diag_n1.json 2> gpurun_out/diag_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --cpu-ncell 0 --e2e-steps 0 > gpurun_out/diag_n2.json 2> gpurun_out/diag_n2.err
grep -c fgmres gpurun_out/diag_n1.err gpurun_out/diag_n2.err
