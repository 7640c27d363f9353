"""More scenarios shaped after the reference's own tests (tests/groupby_test.py, tests/agg_test.py, tests/unique_test.py): key dtypes the
reference routes through other groupers (datetime64, bool, int8), first / last / min / max inside a groupby, string keys with value
aggregators, value_counts / unique on strings, list selections.  Expected values from numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frame(**cols):
    from vaex_b200.frame import Frame
    return Frame(cols)


def test_groupby_datetime_keys():
    # tests/groupby_test.py:35-54 (groupby on a datetime column): keys come back as datetime64
    t = np.array(["2020-01-01", "2020-01-02", "2020-01-01", "2020-01-03", "2020-01-02", "2020-01-01"], dtype="M8[ns]")
    v = np.arange(6, dtype="f8")
    out = _frame(t=t, v=v).groupby("t", agg={"v": ["sum", "count"]}, sort=True)
    assert np.asarray(out["t"]).dtype.kind == "M"
    assert np.asarray(out["t"]).astype("M8[D]").astype(str).tolist() == ["2020-01-01", "2020-01-02", "2020-01-03"]
    assert out["v_sum"].tolist() == [0 + 2 + 5, 1 + 4, 3] and out["v_count"].tolist() == [3, 2, 1]


@pytest.mark.parametrize("dtype", ["?", "i1", "u1", "i2", "f4"])
def test_groupby_small_key_dtypes(dtype):
    # the reference groups bool / int8 / uint8 keys with BinnerInteger (vaex/groupby.py:598-600): same groups, in key order
    rng = np.random.default_rng(3)
    n = 5000
    k = rng.integers(0, 2 if dtype == "?" else 7, n).astype(dtype)
    v = rng.normal(0, 1, n)
    out = _frame(k=k, v=v).groupby("k", agg={"v": ["sum", "count", "min", "max"]}, sort=True)
    uniq = np.unique(k)
    assert np.array_equal(np.asarray(out["k"]), uniq)
    for j, u in enumerate(uniq):
        sel = v[k == u]
        assert out["v_count"][j] == len(sel) and out["v_min"][j] == sel.min() and out["v_max"][j] == sel.max()
        np.testing.assert_allclose(out["v_sum"][j], sel.sum(), rtol=1e-9, atol=1e-12)


def test_groupby_first_last_with_order_column():
    # tests/agg_test.py:503-540 (first / last by an order column) inside a groupby
    from vaex_b200 import agg
    rng = np.random.default_rng(4)
    n = 20_000
    k = rng.integers(0, 40, n).astype("i8")
    t = rng.permutation(n).astype("i8")  # unique order values: no ties
    v = rng.normal(0, 1, n)
    out = _frame(k=k, t=t, v=v).groupby("k", agg=[agg.first("v", "t"), agg.last("v", "t")], sort=True)
    for j, u in enumerate(np.asarray(out["k"])):
        rows = np.nonzero(k == u)[0]
        assert out["v_first"][j] == v[rows[np.argmin(t[rows])]]
        assert out["v_last"][j] == v[rows[np.argmax(t[rows])]]


def test_groupby_string_keys_with_value_aggregators():
    # tests/groupby_test.py:116-176 with a string key: sum / mean / min / max / nunique of a numeric column per string group
    import pyarrow as pa
    from vaex_b200 import agg
    rng = np.random.default_rng(5)
    n = 30_000
    words = np.array(["aap", "noot", "mies", "kees", "", "été", None], dtype=object)
    pick = rng.integers(0, len(words), n)
    s = pa.array(words[pick].tolist())
    v = rng.integers(-50, 50, n).astype("i4")
    out = _frame(s=s, v=v).groupby("s", agg=[agg.sum("v"), agg.mean("v"), agg.min("v"), agg.max("v"), agg.nunique("v")], sort=True)
    keys = out["s"].tolist()
    assert keys[-1] is None and keys[:-1] == sorted(w for w in words if w is not None)  # arrow's bytewise order, the null group last
    for j, w in enumerate(keys):
        sel = v[np.array([x is None for x in words[pick]])] if w is None else v[words[pick] == w]
        assert out["v_sum"][j] == sel.sum() and out["v_min"][j] == sel.min() and out["v_max"][j] == sel.max()
        assert out["v_nunique"][j] == len(np.unique(sel))
        np.testing.assert_allclose(out["v_mean"][j], sel.mean(), rtol=1e-12)


def test_value_counts_and_unique_on_strings():
    # tests/value_counts_test.py / tests/unique_test.py on a string column with missing values
    import pyarrow as pa
    rng = np.random.default_rng(6)
    words = np.array(["x", "yy", "zzz", None, "x" * 300], dtype=object)
    pick = rng.integers(0, len(words), 8000)
    df = _frame(s=pa.array(words[pick].tolist()))
    keys, counts = df.value_counts("s")
    got = dict(zip(keys, np.asarray(counts).tolist()))
    want = {w: int((pick == i).sum()) for i, w in enumerate(words)}
    assert got == want
    assert set(df.unique("s")) == set(words.tolist())


def test_count_with_a_list_of_selections():
    # tests/agg_test.py:62-78 / vaex/cpu.py:744-786: selection=[None, 'x > 0'] -> one grid per selection, stacked first
    rng = np.random.default_rng(7)
    x = rng.normal(0, 1, 10_000)
    y = rng.normal(0, 1, 10_000)
    df = _frame(x=x, y=y)
    both = df.count(binby="y", limits=[-2, 2], shape=8, selection=[None, "x > 0"])
    assert np.asarray(both).shape == (2, 8)
    h_all, _ = np.histogram(y, bins=8, range=(-2, 2))
    h_sel, _ = np.histogram(y[x > 0], bins=8, range=(-2, 2))
    edge_all, edge_sel = int((y == 2).sum()), int(((y == 2) & (x > 0)).sum())  # numpy closes the last bin, the reference does not
    assert np.array_equal(np.asarray(both)[0], h_all - np.eye(8, dtype=int)[-1] * edge_all)
    assert np.array_equal(np.asarray(both)[1], h_sel - np.eye(8, dtype=int)[-1] * edge_sel)


def test_groupby_agg_argument_forms():
    # tests/groupby_test.py:15-33, 56-80: agg='count', {'label': descriptor}, {column: descriptor}, {column: [names]}
    from vaex_b200 import agg
    rng = np.random.default_rng(9)
    n = 8000
    g = rng.integers(0, 5, n).astype("i8")
    x = rng.normal(0, 1, n)
    df = _frame(g=g, x=x)
    counts = np.bincount(g)
    out = df.groupby("g", agg="count", sort=True)
    assert out["count"].tolist() == counts.tolist()
    out = df.groupby("g", agg={"mean_x": agg.mean("x"), "biggest": agg.max("x")}, sort=True)
    np.testing.assert_allclose(out["mean_x"], [x[g == k].mean() for k in range(5)], rtol=1e-12)
    assert out["biggest"].tolist() == [x[g == k].max() for k in range(5)]
    out = df.groupby("g", agg={"x": agg.std("x")}, sort=True)
    np.testing.assert_allclose(out["x"], [x[g == k].std() for k in range(5)], rtol=1e-9)
    out = df.groupby("g", agg={"x": ["min", agg.sum("x")]}, sort=True)
    assert out["x_min"].tolist() == [x[g == k].min() for k in range(5)]
    np.testing.assert_allclose(out["x_sum"], [x[g == k].sum() for k in range(5)], rtol=1e-9)


@pytest.mark.parametrize("device", [False, True])
def test_groupby_on_virtual_columns_and_expressions(device):
    # tests/groupby_test.py:82-98: the key is a virtual column / an expression, evaluated on the device chunk by chunk
    import torch
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(10)
    n = 25_000
    k = rng.integers(0, 12, n).astype("i8")
    v = rng.normal(0, 1, n)
    cols = dict(k=k, v=v)
    if device:
        cols = {c: torch.from_numpy(a).cuda() for c, a in cols.items()}
    df = Frame(cols)
    df.add_virtual_column("kk", "k * 2 + 1")
    for by in ("kk", "k * 2 + 1"):
        out = df.groupby(by, agg={"v": ["sum", "count"]}, sort=True)
        uniq = np.unique(k * 2 + 1)
        assert np.array_equal(np.asarray(out[by]), uniq)
        assert out["v_count"].tolist() == [int((k * 2 + 1 == u).sum()) for u in uniq]
        np.testing.assert_allclose(out["v_sum"], [v[k * 2 + 1 == u].sum() for u in uniq], rtol=1e-9, atol=1e-12)
    # a filtered frame keeps its filter through the groupby
    out = df.filter("v > 0").groupby("kk", agg="count", sort=True)
    assert out["count"].tolist() == [int(((k * 2 + 1 == u) & (v > 0)).sum()) for u in np.unique((k * 2 + 1)[v > 0])]
