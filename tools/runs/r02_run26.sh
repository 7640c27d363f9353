#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r26_pytest.log 2>&1; tail -12 gpurun_out/r26_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r26_bench_reference.json 2> gpurun_out/r26_ref.err; tail -c 300 gpurun_out/r26_ref.err
python bench.py > gpurun_out/r26_bench_final.json 2> gpurun_out/r26_bench.err; tail -c 300 gpurun_out/r26_bench.err
python bench.py --rows 1.25e9 --no-e2e --no-cpu --no-also > gpurun_out/r26_bench_c5rows.json 2>> gpurun_out/r26_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r26_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-also > gpurun_out/r26_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ring -s 4 -c 2 -o gpurun_out/r26_ring python tools/ab_headline.py --rows 1e9 --reps 1 > gpurun_out/r26_ncu.log 2>&1
ls -la gpurun_out/r26_ring.ncu-rep
python - <<'PY'
import json
for f in ("r26_bench_reference.json","r26_bench_final.json","r26_bench_c5rows.json"):
    try:
        d=json.loads([l for l in open("gpurun_out/"+f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), json.dumps(d.get("e2e"))[:260])
        for k,v in (d.get("also") or {}).items(): print("   ", k[:12], v["ms_per_step"], v["frac"], v["parity"], v.get("pass1_ms"), v.get("pass2_ms"))
    except Exception as e: print(f, "ERR", e)
PY
