"""End-to-end through the reference-facing front (Frame -> TaskPartAggregation.process -> b200_bin(HOST)) with numpy columns:
pageable vs page-locked (Frame(pin=True) = b200_host_register), executor threads = slots = streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vaex_b200.frame import Frame
from vaex_b200.execution import Executor

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 28
rng = np.random.default_rng(0)
x = rng.standard_normal(n, dtype=np.float32)
y = rng.standard_normal(n, dtype=np.float32)
for pin in (False, True):
    for nthreads in (1, 4, 8):
        df = Frame(dict(x=x, y=y), executor=Executor(nthreads=nthreads), pin=pin)
        kw = dict(binby=["x", "y"], limits=[[-3, 3], [-3, 3]], shape=1024)
        g = df.count(**kw)
        t0 = time.perf_counter()
        g = df.count(**kw)
        dt = time.perf_counter() - t0
        assert int(g.sum()) <= n
        print(f"pin={pin!s:5} threads={nthreads}: {n/dt:.3e} rows/s  ({8*n/dt/1e9:.1f} GB/s over PCIe)", flush=True)
        if df._pinned:
            df._pinned.release()
