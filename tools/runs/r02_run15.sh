#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k sweep_all --durations=3 > gpurun_out/r15_sweep.log 2>&1
tail -15 gpurun_out/r15_sweep.log
