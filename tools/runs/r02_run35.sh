#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r35_pytest.log 2>&1; tail -3 gpurun_out/r35_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --no-also > gpurun_out/r35_bench.json 2> gpurun_out/r35_bench.err; tail -c 200 gpurun_out/r35_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r35_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['chunks_16M']['value'], d['gpu_launches'])"
