#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_strings.py -q -m gpu > gpurun_out/r13_strings.log 2>&1
echo "strings rc=$?" >> gpurun_out/r13_strings.log
tail -30 gpurun_out/r13_strings.log
