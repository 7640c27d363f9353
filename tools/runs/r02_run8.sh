#!/bin/bash
# round 2, call 8: device-side set finalisation (radix sort), bounce ring e2e; hash + frame tests, bench
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r8_pytest.log
tail -15 gpurun_out/r8_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err
echo "bench rc=$?"
tail -c 2000 gpurun_out/r8_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r8_bench.json"))
print("value",d["value"],"frac",d["roofline"]["frac"])
print("e2e",json.dumps(d["e2e"]))
for k,v in d.get("also",{}).items(): print(k[:12], v["rows_per_s"], v["frac"], v["parity"], v.get("pass1_ms"), v.get("pass2_ms"))
PY
