#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_strings.py -q -m gpu > gpurun_out/r12_strings.log 2>&1
echo "strings rc=$?" >> gpurun_out/r12_strings.log
tail -60 gpurun_out/r12_strings.log
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_strings.py > gpurun_out/r12_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r12_pytest.log
tail -5 gpurun_out/r12_pytest.log
