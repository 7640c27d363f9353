#!/usr/bin/env python
"""C4 pass 1 (ordered_set over 1e9 int64 rows, 1e6 distinct keys): wall-clock split of update / finalisation, twice (the second
build reuses the capacity hint).  Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vaex_b200 import _lib, superutils
rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
ctx = _lib.context()
gen = torch.Generator(device="cuda").manual_seed(7)
keys = torch.randint(0, 1_000_000, (rows,), device="cuda", dtype=torch.int64, generator=gen) * 256 + 5
torch.cuda.synchronize()
for rep in range(3):
    ctx.sync(0)
    t0 = time.perf_counter()
    s = superutils.ordered_set_int64(7)
    s.update(keys, -1)
    ctx.sync(0)
    t1 = time.perf_counter()
    n = len(s)
    ctx.sync(0)
    t2 = time.perf_counter()
    print(f"rep {rep}: update {1e3 * (t1 - t0):.2f} ms, finalise {1e3 * (t2 - t1):.2f} ms, keys {n}", flush=True)
    del s
