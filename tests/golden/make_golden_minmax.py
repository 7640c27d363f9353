"""Generate tests/golden/minmax_golden.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/vaexfast*.so): the limits
pre-pass df.minmax -> TaskStatistic(OP_MIN_MAX) -> vaexfast.statisticNd (src/vaexfast.cpp:1089-1101, 1167-1290), driven by the
restated TaskPartStatistic.process (oracle/ref_driver.py:minmax).  Run where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden_minmax.py

Cases: every dtype of the path, plain / masked / byte-swapped, NaN and infinities, integers beyond 2^24 and 2^53 (the reference
rounds them through float32 / float64 on the way: part of the observable result), all-NaN and empty columns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver as R  # noqa: E402


def cases():
    rng = np.random.default_rng(1089)
    out = {}

    def add(name, data):
        mask = np.ma.getmaskarray(data) if np.ma.isMaskedArray(data) else None
        raw = np.asarray(data.data if mask is not None else data)
        out[f"{name}/data"] = raw.view(raw.dtype.newbyteorder("=")) if raw.dtype.itemsize > 1 else raw
        out[f"{name}/dtype"] = np.array(raw.dtype.str)
        if mask is not None:
            out[f"{name}/mask"] = mask
        out[f"{name}/raw"] = R.minmax(data, raw=True)
        out[f"{name}/result"] = R.minmax(data)

    n = 5000
    for dt in ("float64", "float32", "int64", "int32", "int16", "int8", "uint64", "uint32", "uint16", "uint8", "bool"):
        d = np.dtype(dt)
        if d.kind == "f":
            v = (rng.standard_normal(n) * 1e3).astype(d)
            v[rng.random(n) < 0.1] = np.nan
            v[5], v[6] = np.inf, -np.inf
        elif d.kind == "b":
            v = rng.integers(0, 2, n).astype(d)
        else:
            info = np.iinfo(d)
            v = rng.integers(info.min // 2, info.max // 2, n, dtype=np.int64 if d.kind == "i" else np.uint64).astype(d)
            v[7] = info.max - 1  # beyond 2^24 / 2^53 for the wide types: rounded by the reference's float cast
            if d.kind == "i":
                v[8] = info.min + 1
        add(dt, v)
        add(dt + "_masked", np.ma.array(v, mask=rng.random(n) < 0.3))
        if d.itemsize > 1:
            add(dt + "_swapped", v.astype(d.newbyteorder(">")))
            add(dt + "_swapped_masked", np.ma.array(v.astype(d.newbyteorder(">")), mask=rng.random(n) < 0.3))
    add("all_nan", np.full(100, np.nan))
    add("empty_f8", np.zeros(0, "f8"))
    add("empty_f4", np.zeros(0, "f4"))
    add("all_masked", np.ma.array(np.arange(10.0), mask=np.ones(10, bool)))
    add("one_row", np.array([42.5], "f4"))
    add("neg_zero", np.array([0.0, -0.0, 0.0]))
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    data = cases()
    np.savez_compressed(os.path.join(here, "minmax_golden.npz"), **data)
    print("wrote", len(data), "arrays")
