#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_frame.py -x -q -m gpu -k "combine" > gpurun_out/r16.log 2>&1
tail -40 gpurun_out/r16.log
