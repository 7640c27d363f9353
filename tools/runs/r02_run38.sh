#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r38_pytest.log 2>&1; tail -3 gpurun_out/r38_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
