"""String keys on the device (csrc/hashset.cu string section, superutils.ordered_set_string, AggCount_string, AggNUnique_string):
against the golden vectors of the compiled reference's ordered_set<> over StringList64 and against the oracle / plain Python."""
import collections
import random

import numpy as np
import pytest

import golden_util

pytestmark = pytest.mark.gpu

STRINGS = golden_util.load_strings()


@pytest.mark.parametrize("name", sorted(k for k in STRINGS if k.startswith("strset_")))
def test_string_set_matches_golden(name):
    from vaex_b200 import superutils
    c = STRINGS[name]
    s = superutils.ordered_set_string(int(name.split("_")[1]))
    for k in range(int(c["ncalls"])):
        strs = golden_util.unpack_strings(c[f"c{k}_offsets"], c[f"c{k}_bytes"], c[f"c{k}_mask"])
        vals, mi = s.update(strs, 0, return_values=True)
        assert np.array_equal(vals, c[f"c{k}_values"]) and np.array_equal(mi, c[f"c{k}_map_index"]), k
    assert s.keys() == golden_util.unpack_strings(c["key_offsets"], c["key_bytes"], c["key_nulls"])
    assert s.offsets() == c["shard_offsets"].tolist()
    probe = golden_util.unpack_strings(c["probe_offsets"], c["probe_bytes"], c["probe_mask"])
    assert np.array_equal(s.map_ordinal(probe), c["probe_ordinals"])
    assert [len(s), s.null_count, s.null_index] == c["info"].tolist()


def test_string_set_large_arrow_input(oracle):
    """1e5 rows of a pyarrow string array with 5e3 distinct keys of mixed lengths (incl. > 8 and multi-byte), table growth inside one
    call, two calls, against the oracle restatement"""
    import pyarrow as pa
    from vaex_b200 import superutils
    rnd = random.Random(5)
    words = ["".join(rnd.choice("abcdefghijklmnopqrstuvwxyzäö") for _ in range(rnd.randint(0, 24))) for _ in range(5000)]
    for nmaps in (1, 7):
        s, o = superutils.ordered_set_string(nmaps), oracle.StringOrderedSet(nmaps)
        for call in range(2):
            strs = [rnd.choice(words) if rnd.random() > 0.02 else None for _ in range(100_000)]
            arr = pa.array(strs, type=pa.string())
            got, want = s.update(arr, return_values=True), o.update(strs, 0, True)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        assert s.keys() == o.keys() and s.offsets() == o.offsets() and len(s) == len(o)
        probe = [rnd.choice(words + ["nope"]) for _ in range(5000)]
        assert np.array_equal(s.map_ordinal(pa.array(probe)), o.map_ordinal(probe))


def test_groupby_count_nunique_on_string_columns():
    """df.groupby(string column).agg(...), df.count(string column, binby=...), df.nunique(string column, binby=...): a default vaex
    use that round 1 handed back to the CPU"""
    import pyarrow as pa
    from vaex_b200 import execution
    from vaex_b200.frame import Frame
    rnd = random.Random(11)
    n = 60_000
    cities = ["amsterdam", "groningen", "den haag", "utrecht", "a", "", "zürich", "san francisco bay area"]
    key = [rnd.choice(cities) if rnd.random() > 0.05 else None for _ in range(n)]
    tag = [rnd.choice(["x", "yy", "zzz", None]) for _ in range(n)]
    rng = np.random.default_rng(2)
    v = rng.standard_normal(n)
    g = rng.integers(0, 5, n).astype("i4")
    df = Frame({"key": pa.array(key), "tag": pa.array(tag), "v": v, "g": g}, executor=execution.Executor(nthreads=2, chunk_size=17_000))
    out = df.groupby("key").agg({"v": ["sum", "count"]})
    want_n = collections.Counter(key)
    got = {k: (c, s) for k, c, s in zip(out["key"], out["count"], out["v_sum"])}
    assert set(got) == set(want_n)
    for k in want_n:
        assert got[k][0] == want_n[k]
        assert np.isclose(got[k][1], v[[i for i, x in enumerate(key) if x == k]].sum(), rtol=1e-9, atol=1e-9)
    # count(string) per ordinal bin: non-null strings only
    df.categorize("g", 0, 5)
    cnt = df.count("tag", binby=["g"])
    assert cnt.tolist() == [sum(1 for i in range(n) if g[i] == b and tag[i] is not None) for b in range(5)]
    # nunique(string): the reference's per-cell counter counts the null as ONE more value, and dropmissing subtracts the cell's null
    # ROW count from that (src/agg_nunique_string.cpp:22-27: `counter->count() - counter->null_count`)
    def ref_nunique(col, b, dropmissing):
        rows = [col[i] for i in range(n) if g[i] == b]
        nulls = sum(1 for r in rows if r is None)
        total = len({r for r in rows if r is not None}) + (1 if nulls else 0)
        return total - nulls if dropmissing else total
    for col, name in ((tag, "tag"), (key, "key")):
        for drop in (False, True):
            nu = df.nunique(name, binby=["g"], dropmissing=drop)
            assert nu.tolist() == [ref_nunique(col, b, drop) for b in range(5)], (name, drop)


def test_string_set_pickle_round_trip():
    # vaex/hash.py:28-36: ordered_set_string pickles as (keys in ordinal order, null_index, ...) and comes back with the same ordinals
    import pickle
    import pyarrow as pa
    from vaex_b200 import superutils
    words = ["noot", "aap", None, "mies", "aap", "", "noot", "kees"]
    s = superutils.ordered_set_string(3)
    s.update(pa.array(words))
    t = pickle.loads(pickle.dumps(s))
    assert t.keys() == s.keys() and t.null_count == s.null_count and len(t) == len(s)
    probe = pa.array(["kees", "aap", "xx", None, ""])
    assert np.array_equal(np.asarray(t.map_ordinal(probe)), np.asarray(s.map_ordinal(probe)))
