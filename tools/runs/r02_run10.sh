#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_expr.py -x -q -m gpu > gpurun_out/r10_expr.log 2>&1
echo "expr rc=$?" >> gpurun_out/r10_expr.log
tail -30 gpurun_out/r10_expr.log
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_expr.py > gpurun_out/r10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r10_pytest.log
tail -5 gpurun_out/r10_pytest.log
