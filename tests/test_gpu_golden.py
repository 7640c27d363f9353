"""GPU parity against the committed golden vectors of the compiled reference (no oracle in the loop)."""
import numpy as np
import pytest

import golden_util
from helpers import b200_binby, same

pytestmark = pytest.mark.gpu

GOLD = golden_util.load()
BINBY = sorted(k for k in GOLD if not k.startswith(("set_", "hash64")))
SETS = sorted(k for k in GOLD if k.startswith("set_"))


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("name", BINBY)
def test_binby_matches_golden(name, device):
    binners, aggs, n, expected = golden_util.binby_case(GOLD[name])
    got = b200_binby(binners, aggs, n, device=device)
    for a, w, g in zip(aggs, expected, got):
        if a["op"] in ("sum", "sum_moment") and np.asarray(a["data"]).dtype.kind == "f":
            assert w.shape == g.shape and np.allclose(g, w, rtol=1e-6, atol=1e-9 * max(1.0, float(np.abs(w).max()))), (name, a["op"])
        else:
            assert same(w, g), (name, a["op"])


@pytest.mark.parametrize("name", SETS)
def test_sets_match_golden(name):
    from vaex_b200 import superutils
    c = GOLD[name]
    dtype, nmaps = name.split("_")[1], int(name.split("_")[2])
    s = getattr(superutils, "ordered_set_" + dtype)(nmaps)
    vals, mi = s.update(c["keys"], c["mask"], 0, return_values=True)
    assert np.array_equal(vals, c["values"]) and np.array_equal(mi, c["map_index"])
    assert np.array_equal(s.key_array(), c["key_array"], equal_nan=True)
    assert s.offsets() == c["offsets"].tolist()
    mo = s.map_ordinal(c["keys"])
    assert mo.dtype == c["map_ordinal"].dtype and np.array_equal(mo, c["map_ordinal"])
    assert [s.null_index, s.nan_index, s.null_count, s.nan_count] == c["null_nan"].tolist()


def test_hash64_golden():
    from vaex_b200 import superutils
    for i, o in zip(GOLD["hash64"]["in"], GOLD["hash64"]["out"]):
        assert superutils.hash(int(i)) == int(o)


MINMAX = golden_util.load_minmax()


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("name", sorted(MINMAX))
def test_minmax_matches_golden(name, device):
    """Frame.minmax -> b200_minmax (csrc/minmax.cu) against the compiled reference's OP_MIN_MAX vectors: raw grid values and the
    pair cast back to the column dtype, host chunks and device-resident columns."""
    from helpers import to_device
    from vaex_b200.frame import Frame
    data, raw, result = MINMAX[name]
    col = data
    if device:
        if np.ma.isMaskedArray(data) or (data.dtype.itemsize > 1 and data.dtype.byteorder not in ("=", "|", "<")):
            pytest.skip("device columns are plain native arrays")
        col = to_device(np.ascontiguousarray(data))
        if data.dtype.kind == "u" and data.dtype.itemsize > 1:
            pytest.skip("torch carries unsigned columns as signed views")
    df = Frame({"v": col})
    got_raw = df.minmax("v", raw=True)
    assert np.array_equal(got_raw, raw, equal_nan=True), (name, got_raw, raw)
    got = df.minmax("v")
    assert got.dtype == result.dtype and np.array_equal(got, result, equal_nan=True)


AGGLIST = golden_util.load_agglist()


@pytest.mark.parametrize("name", sorted(AGGLIST["cases"]))
def test_agg_list_matches_golden(name):
    """superagg.AggList_<dtype>_int64 (csrc/list.cu) against the compiled reference's vectors: same two bin() calls, offsets and
    values bit-identical (NaN slots are NaN; the NULL slots are unspecified in the reference and zero in the file and on the device)."""
    from vaex_b200 import superagg
    c = AGGLIST["cases"][name]
    x, cut = AGGLIST["x"], AGGLIST["cut"]
    n = len(x)
    b = superagg.BinnerOrdinal_int32(1, "x", AGGLIST["ncat"], 0, False, False)
    g = superagg.Grid([b])
    a = getattr(superagg, f"AggList_{c['dtype']}_int64")(g, 1, 1, c["dropnan"], c["dropnull"])
    for i1, i2 in ((0, cut), (cut, n)):
        b.set_data(0, np.ascontiguousarray(x[i1:i2]))
        a.set_data(0, np.ascontiguousarray(c["v"][i1:i2]), 0)
        if c["masked"]:
            a.set_data_mask(0, np.ascontiguousarray(c["valid"][i1:i2]))
        g.bin(0, [a], i2 - i1)
    offsets, values = a.result_arrays()
    assert np.array_equal(offsets, c["offsets"])
    assert values.dtype == c["values"].dtype and np.array_equal(values, c["values"], equal_nan=True)
