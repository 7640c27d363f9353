#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "not sweep_all" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-also > gpurun_out/r20_bench_n1.json 2> gpurun_out/r20_bench_n1.err
tail -c 600 gpurun_out/r20_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r20_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], json.dumps(d.get("e2e"))[:900])
PY
python tools/probe_e2e_threads.py 28 16,32 20,24 > gpurun_out/r20_probe.txt 2>&1; cat gpurun_out/r20_probe.txt
