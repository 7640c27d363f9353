#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29517 tools/check_multigpu.py > gpurun_out/r21_check_multigpu.txt 2>&1; tail -12 gpurun_out/r21_check_multigpu.txt
$T --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r21_bench_n2.json 2> gpurun_out/r21_bench_n2.err; tail -c 400 gpurun_out/r21_bench_n2.err
$T --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-overlap > gpurun_out/r21_bench_n2_serial.json 2>> gpurun_out/r21_bench_n2.err
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also > gpurun_out/r21_bench_n1.json 2>> gpurun_out/r21_bench_n2.err
python - <<'PY'
import json
for f in ("r21_bench_n1.json","r21_bench_n2.json","r21_bench_n2_serial.json"):
    try:
        d=json.loads([l for l in open("gpurun_out/"+f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], json.dumps(d.get("e2e"))[:300])
    except Exception as e: print(f, "ERR", e)
PY
