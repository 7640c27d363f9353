#!/usr/bin/env python
"""Where does the host-chunk path spend its time?  Frame.count over pageable numpy columns for several thread counts / chunk
sizes, plus the share of wall time inside b200_bin (ctypes call, GIL released) vs in Python (GIL held)."""
import json, os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from vaex_b200 import _lib, execution
from vaex_b200.frame import Frame

n = 1 << 27
rng = np.random.default_rng(0)
x = rng.standard_normal(n, dtype=np.float32)
y = rng.standard_normal(n, dtype=np.float32)
L = _lib.lib()
orig = L.b200_bin
acc = {"t": 0.0, "n": 0}
lock = threading.Lock()

def timed_bin(*a):
    t0 = time.perf_counter()
    r = orig(*a)
    dt = time.perf_counter() - t0
    with lock:
        acc["t"] += dt
        acc["n"] += 1
    return r

for nthreads in (1, 4, 8, 16, 32):
    for chunk in (1 << 20, 1 << 22, 1 << 24):
        ex = execution.Executor(nthreads=nthreads, chunk_size_max=chunk)
        df = Frame({"x": x, "y": y}, executor=ex)
        for rep in range(3):
            acc["t"], acc["n"] = 0.0, 0
            L.b200_bin = timed_bin if rep == 2 else orig
            t0 = time.perf_counter()
            g = df.count(binby=["x", "y"], limits=[[-3, 3], [-3, 3]], shape=1024, edges=True)
            dt = time.perf_counter() - t0
        L.b200_bin = orig
        assert int(g.sum()) == n
        print(json.dumps(dict(threads=nthreads, chunk=ex.chunk_size_for(n), rows_per_s=n / dt, ms=dt * 1e3, calls=acc["n"], in_bin_ms_sum=acc["t"] * 1e3,
                              in_bin_ms_per_call=acc["t"] * 1e3 / max(acc["n"], 1))), flush=True)
