#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r29_pytest.log 2>&1; tail -4 gpurun_out/r29_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-also > gpurun_out/r29_bench.json 2> gpurun_out/r29_bench.err; tail -c 300 gpurun_out/r29_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r29_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['chunks_16M']['value'], d['gpu_launches'])"
