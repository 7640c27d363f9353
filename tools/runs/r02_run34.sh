#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "ring or parity and not sweep_all" 2>&1 | tail -2
(for pct in 100 80 90; do
B200_RING_STATIC_PCT=$pct python tools/ab_headline.py --reps 20 --tag static_pct_$pct
B200_RING_STATIC_PCT=$pct python tools/ab_headline.py --reps 20 --tag static_pct_$pct --occupy 16,120
done
B200_RING_STATIC_PCT=80 python tools/ab_headline.py --reps 20 --tag static_pct_80 --occupy 32,200
B200_RING_STATIC_PCT=80 python tools/ab_headline.py --reps 20 --tag static_pct_80 --rows 2.5e8
B200_RING_STATIC_PCT=100 python tools/ab_headline.py --reps 20 --tag static_pct_100 --rows 2.5e8
) > gpurun_out/r34_ab.jsonl 2> gpurun_out/r34_ab.err
tail -3 gpurun_out/r34_ab.err; python - <<'PY'
import json
for l in open("gpurun_out/r34_ab.jsonl"):
    d=json.loads(l); print(d["tag"], d["rows"], d["occupy"], round(d["ms_median"],4), d["count_ok"], d["grid_sha"])
PY
