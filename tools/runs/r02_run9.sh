#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r9_pytest.log
tail -15 gpurun_out/r9_pytest.log
python - <<'PY' > gpurun_out/r9_minmax_timing.json
import json, torch, numpy as np, sys
sys.path.insert(0, ".")
from vaex_b200 import _lib, engine
from vaex_b200.frame import Frame
ctx = _lib.context(0)
n = 1 << 30
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
out = {}
for dt, name in ((torch.float64, "f8"), (torch.float32, "f4")):
    x = torch.empty(n, dtype=dt, device="cuda").normal_()
    df = Frame({"x": x})
    df.minmax("x")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = engine.slot_stream(ctx, 0)
    e0.record(st)
    for _ in range(5):
        r = df.minmax("x", raw=True)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = n * x.element_size() / (ms * 1e-3) / 1e9
    out[name] = dict(rows=n, ms=ms, gbs=gbs, frac_of_measured_hbm=gbs / peak, result=r.tolist(), check=[float(x.min()), float(x.max())])
    del x, df
print(json.dumps(out))
PY
cat gpurun_out/r9_minmax_timing.json
