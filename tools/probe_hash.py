import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vaex_b200 import _lib, superutils
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**9
ctx = _lib.context(0)
gen = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(0, 1_000_000, (n,), device="cuda", dtype=torch.int64, generator=gen) * 256 + 5
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); s = superutils.ordered_set_int64(7); torch.cuda.synchronize(); t1 = time.perf_counter()
    s.update(keys, -1); ctx.sync(); torch.cuda.synchronize(); t2 = time.perf_counter()
    k = len(s); t3 = time.perf_counter()
    c = s.map_ordinal(keys[:1000]); t4 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms  update {1e3*(t2-t1):.1f} ms  finalize(len) {1e3*(t3-t2):.1f} ms  map1000 {1e3*(t4-t3):.1f} ms  keys {k}", flush=True)
    del s
