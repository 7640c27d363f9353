#!/usr/bin/env python
"""Where does the host-chunk path spend its time?  Frame.count over pageable numpy columns for several thread counts / chunk
sizes; per configuration the wall time, the C side's own accounting (b200_ctx_host_stats: waiting for a bounce piece, memcpy,
enqueue, whole b200_bin — summed over threads) and the feed loop's timeline (first start, last end, busy time per worker).

    B200_BOUNCE_PIECE_KB=1024 B200_BOUNCE_COUNT=8 python tools/probe_e2e_threads.py [rows_log2] [threads,...] [chunk_log2,...]"""
import json, os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from vaex_b200 import _lib, execution, taskpart
from vaex_b200.frame import Frame

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
threads = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 8, 16, 32]
chunks = [1 << int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1 << 20, 1 << 22, 1 << 24]
rng = np.random.default_rng(0)
x = rng.standard_normal(n, dtype=np.float32)
y = rng.standard_normal(n, dtype=np.float32)
ctx = _lib.context()
spans = []
lock = threading.Lock()
orig = taskpart.TaskPartAggregation.process

def process(self, thread_index, i1, i2, *a, **k):
    t0 = time.perf_counter()
    r = orig(self, thread_index, i1, i2, *a, **k)
    with lock:
        spans.append((thread_index, t0, time.perf_counter()))
    return r

taskpart.TaskPartAggregation.process = process
# every C-ABI call, timed: which one holds the pass back before the first chunk is fed?
calls = []
L = _lib.lib()
def wrap(name, fn):
    def f(*a):
        t0 = time.perf_counter()
        r = fn(*a)
        t1 = time.perf_counter()
        if t1 - t0 > 2e-3:
            with lock:
                calls.append((name, t0, t1 - t0, threading.get_ident() == main_id))
        return r
    return f
main_id = threading.get_ident()
for name in ("b200_agg_create", "b200_agg_destroy", "b200_agg_reset", "b200_agg_reset_on", "b200_agg_read", "b200_agg_merge", "b200_ctx_sync", "b200_bin"):
    setattr(L, name, wrap(name, getattr(L, name)))
for nthreads in threads:
    for chunk in chunks:
        ex = execution.Executor(nthreads=nthreads, chunk_size_max=chunk)
        df = Frame({"x": x, "y": y}, executor=ex)
        for rep in range(4):
            spans.clear()
            calls.clear()
            ctx.host_stats(reset=True)
            t0 = time.perf_counter()
            g = df.count(binby=["x", "y"], limits=[[-3, 3], [-3, 3]], shape=1024, edges=True)
            t1 = time.perf_counter()
            st = ctx.host_stats()
            assert int(g.sum()) == n
            if rep < 2:
                continue
            busy = {}
            for t, a, b in spans:
                busy[t] = busy.get(t, 0.0) + (b - a)
            print(json.dumps(dict(threads=nthreads, chunk=ex.chunk_size_for(n), rows_per_s=round(n / (t1 - t0) / 1e9, 3), ms=round((t1 - t0) * 1e3, 1), calls=st["calls"],
                                  workers=len(busy), first_start_ms=round((min(a for _, a, _ in spans) - t0) * 1e3, 2), last_end_ms=round((max(b for _, _, b in spans) - t0) * 1e3, 1),
                                  busy_ms_max=round(max(busy.values()) * 1e3, 1), busy_ms_sum=round(sum(busy.values()) * 1e3, 1),
                                  c_wait_ms=round(st["wait_ms"], 1), c_memcpy_ms=round(st["memcpy_ms"], 1), c_enqueue_ms=round(st["enqueue_ms"], 1), c_bin_ms=round(st["bin_ms"], 1),
                                  pieces=st["pieces"],
                                  slow_calls=[(nm, round((a - t0) * 1e3, 1), round(d * 1e3, 1), m) for nm, a, d, m in calls if nm != "b200_bin" or d > 0.03][:8])), flush=True)
