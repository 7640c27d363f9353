#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "not sweep_all" 2>&1 | tail -3
python tools/c4_pass1_breakdown.py 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r23_bench_n1.json 2> gpurun_out/r23_bench_n1.err
tail -c 600 gpurun_out/r23_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r23_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
print(json.dumps(d.get("e2e"))[:700])
print(json.dumps(d.get("cpu_baseline"))[:700])
for k,v in d["also"].items(): print(k[:12], v["ms_per_step"], v["frac"], v["parity"], v.get("pass1_ms"), v.get("pass2_ms"))
PY
