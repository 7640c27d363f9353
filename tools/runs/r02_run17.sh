#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-also > gpurun_out/r17_bench_n1.json 2> gpurun_out/r17_bench_n1.err
tail -c 600 gpurun_out/r17_bench_n1.err
python bench.py --steps 20 --warmup 5 --no-also --no-e2e --no-cpu --no-overlap > gpurun_out/r17_bench_n1_serial.json 2>> gpurun_out/r17_bench_n1.err
python - <<'PY'
import json
for f in ("r17_bench_n1.json","r17_bench_n1_serial.json"):
    d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], json.dumps(d.get("e2e"))[:900])
PY
python tools/probe_e2e_threads.py > gpurun_out/r17_probe.jsonl 2>&1
cat gpurun_out/r17_probe.jsonl
