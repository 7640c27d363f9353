#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r14_pytest.log
tail -40 gpurun_out/r14_pytest.log
