#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scenarios.py tests/test_gpu_frame.py -q -m gpu > gpurun_out/r40.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r40.log | tail; grep -E "^E  " gpurun_out/r40.log | head -30
