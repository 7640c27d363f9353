#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_strings.py tests/test_gpu_scenarios.py tests/test_gpu_frame.py -x -q -m gpu > gpurun_out/r39.log 2>&1; tail -12 gpurun_out/r39.log
