#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_golden.py tests/test_gpu_protocol.py tests/test_gpu_frame.py -q -m gpu -k "agg_list or list or readme" > gpurun_out/r42.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r42.log | tail -8; grep -E "^E  " gpurun_out/r42.log | head -12
