#!/bin/bash
mkdir -p gpurun_out
python tools/bench_configs.py --rows 1e8 --configs NU,SG > gpurun_out/r30_features.jsonl 2> gpurun_out/r30_features.err; tail -3 gpurun_out/r30_features.err; cat gpurun_out/r30_features.jsonl | cut -c1-700
python - <<'PY'
# where does the sparse groupby spend its time?
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from vaex_b200.frame import Frame
rows=100_000_000
gen = torch.Generator(device="cuda").manual_seed(42)
k1 = torch.randint(0, 30_000, (rows,), device="cuda", dtype=torch.int64, generator=gen)
k2 = torch.randint(0, 30_000, (rows,), device="cuda", dtype=torch.int64, generator=gen)
k2 = (k2 + k1) % 30_000
w = torch.empty(rows, dtype=torch.float64, device="cuda").normal_(generator=gen)
df = Frame(dict(k1=k1, k2=k2, w=w))
def run():
    gb = df.groupby(["k1","k2"], combine="auto"); out = gb.agg({"w":["sum","count"]}); return out
run()
pr=cProfile.Profile(); pr.enable(); out=run(); pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
PY
