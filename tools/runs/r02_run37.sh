#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scenarios.py -q -m gpu > gpurun_out/r37_scen.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r37_scen.log | tail -15
grep -E "^E  " gpurun_out/r37_scen.log | head -40
python -m pytest tests -x -q -m gpu -k "not sweep_all" --deselect tests/test_gpu_scenarios.py 2>&1 | tail -3
