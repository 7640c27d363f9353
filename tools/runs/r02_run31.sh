#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "not sweep_all" > gpurun_out/r31_pytest.log 2>&1; tail -15 gpurun_out/r31_pytest.log
