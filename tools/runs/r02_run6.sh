#!/bin/bash
# round 2, call 6: full GPU test suite, smoke, bench (with also-configs + new e2e)
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r6_pytest.log
tail -5 gpurun_out/r6_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r6_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r6_smoke.log
tail -3 gpurun_out/r6_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r6_bench.err
cat gpurun_out/r6_bench.json
