#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "not sweep_all" 2>&1 | tail -3
python tools/c4_pass1_breakdown.py 2>&1 | tail -3
python tools/c4_pass1_breakdown.py 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r24_bench.json 2> gpurun_out/r24_bench.err; tail -c 300 gpurun_out/r24_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r24_bench.json").read().strip().splitlines()[-1])
for k,v in d["also"].items(): print(k[:12], v["ms_per_step"], v["frac"], v["parity"], v.get("pass1_ms"), v.get("pass2_ms"))
PY
