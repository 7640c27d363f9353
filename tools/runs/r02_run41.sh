#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_multigpu.py > gpurun_out/r41_check_multigpu.txt 2>&1; tail -9 gpurun_out/r41_check_multigpu.txt
