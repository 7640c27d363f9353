#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_frame.py tests/test_gpu_strings.py -x -q -m gpu > gpurun_out/r36.log 2>&1; tail -30 gpurun_out/r36.log
