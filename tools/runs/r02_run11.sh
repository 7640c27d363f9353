#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_expr.py -q -m gpu > gpurun_out/r11_expr.log 2>&1
echo "expr rc=$?" >> gpurun_out/r11_expr.log
tail -60 gpurun_out/r11_expr.log
