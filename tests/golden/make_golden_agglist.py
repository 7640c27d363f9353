"""Generate tests/golden/agglist_golden.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/superagg*.so):
AggList_<dtype>_int64 (src/agg_list.cpp) over an ordinal binner, fed in two bin() calls, for the four dropnan / dropnull
combinations, float and integer values, with and without a data mask.  Run where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden_agglist.py

get_result() of the reference hands (offsets, values) to vaex.arrow.convert.list_from_arrays; vaex cannot be imported here, so a
stub module with that one function (returning the two arrays) stands in for it.  The value slots of NULL rows are uninitialised
memory in the reference: the generator zeroes them (their positions follow from the counts) so the file is reproducible."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import ref_driver as R  # noqa: E402


def _stub_vaex():
    vaex = types.ModuleType("vaex")
    arrow = types.ModuleType("vaex.arrow")
    convert = types.ModuleType("vaex.arrow.convert")
    convert.list_from_arrays = lambda offsets, values: (np.array(offsets), np.array(values))
    vaex.arrow, arrow.convert = arrow, convert
    sys.modules.update({"vaex": vaex, "vaex.arrow": arrow, "vaex.arrow.convert": convert})


def main():
    _stub_vaex()
    sa, _ = R.modules()
    n, ncat, cut = 4000, 9, 1777
    out = {}
    rng = np.random.default_rng(47)
    x = rng.integers(-1, ncat + 1, n).astype("i4")
    out["x"] = x
    out["ncat"] = np.array(ncat)
    out["cut"] = np.array(cut)
    for dt in ("float64", "float32", "int32", "int64", "uint8"):
        d = np.dtype(dt)
        v = (rng.standard_normal(n) * 100).astype(d) if d.kind == "f" else rng.integers(0, 200, n).astype(d)
        if d.kind == "f":
            v[rng.random(n) < 0.15] = np.nan
        valid = (rng.random(n) < 0.8).astype("u1")
        out[f"{dt}/v"], out[f"{dt}/valid"] = v, valid
        for masked in (False, True):
            for dropnan in (False, True):
                for dropnull in (False, True):
                    b = sa.BinnerOrdinal_int32(1, "x", ncat, 0, False, False)
                    g = sa.Grid([b])
                    a = getattr(sa, f"AggList_{dt}_int64")(g, 1, 1, dropnan, dropnull)
                    keep = []
                    for i1, i2 in ((0, cut), (cut, n)):
                        xs, vs = np.ascontiguousarray(x[i1:i2]), np.ascontiguousarray(v[i1:i2])
                        keep += [xs, vs]
                        b.set_data(0, xs)
                        a.set_data(0, vs, 0)
                        if masked:
                            ms = np.ascontiguousarray(valid[i1:i2])
                            keep.append(ms)
                            a.set_data_mask(0, ms)
                        g.bin(0, [a], i2 - i1)
                    offsets, values = a.get_result()
                    offsets, values = offsets.astype(np.int64), np.array(values)
                    # zero the null slots (uninitialised in the reference): they are the tail of every cell's list
                    cells = O.flat_indices([O.ordinal(x, ncat, 0)], n)[0].astype(np.int64)  # the binner's cell of every row
                    if masked and not dropnull:
                        # which mask entry the reference looks at: aggregate() runs per 1024-row block of a call and indexes the
                        # mask WITHOUT the block offset (src/agg_list.cpp:96), i.e. row r of a call is judged by mask[r % 1024]
                        seen = np.concatenate([valid[i1:i2][np.arange(i2 - i1) % 1024] for i1, i2 in ((0, cut), (cut, n))])
                        nulls = np.bincount(cells[seen == 0], minlength=len(offsets) - 1)
                        for c in range(len(offsets) - 1):
                            if nulls[c]:
                                values[offsets[c + 1] - nulls[c]:offsets[c + 1]] = 0
                    name = f"{dt}/{'masked' if masked else 'plain'}_dropnan{int(dropnan)}_dropnull{int(dropnull)}"
                    out[f"{name}/offsets"], out[f"{name}/values"] = offsets, values
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "agglist_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays")


if __name__ == "__main__":
    main()
