#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r25_bench_n8.json 2> gpurun_out/r25_bench_n8.err; tail -c 500 gpurun_out/r25_bench_n8.err
$T --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --rows 1.25e9 --no-e2e > gpurun_out/r25_bench_n8_c5.json 2>> gpurun_out/r25_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > gpurun_out/r25_bench_n4.json 2>> gpurun_out/r25_bench_n8.err
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also > gpurun_out/r25_bench_n1.json 2>> gpurun_out/r25_bench_n8.err
python - <<'PY'
import json
for f in ("r25_bench_n1.json","r25_bench_n4.json","r25_bench_n8.json","r25_bench_n8_c5.json"):
    try:
        d=json.loads([l for l in open("gpurun_out/"+f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["workload"][-60:], json.dumps(d.get("e2e"))[:200])
    except Exception as e: print(f, "ERR", e)
PY
