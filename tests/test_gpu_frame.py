"""API-level GPU tests shaped after the reference's own tests (tests/agg_test.py, tests/count_test.py, tests/groupby_test.py),
run through vaex_b200.frame.Frame -> TaskPartAggregation -> superagg mirror -> C ABI -> CUDA."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def base_columns():
    # the numeric core of tests/common.py:313-381 create_base_ds(): x, y = x**2, masked m, NaN n
    x = np.arange(20, dtype=">f8")
    y = x ** 2
    m = np.ma.array(x.astype("f8"), mask=(np.arange(20) % 2 == 1))
    n = x.astype("f8").copy()
    n[[1, 3]] = np.nan
    i8 = np.arange(20, dtype="i1") - 10
    return dict(x=x, y=y, m=m, n=n, i8=i8)


def frame(chunk=None, **kw):
    from vaex_b200.execution import Executor
    from vaex_b200.frame import Frame
    ex = Executor(nthreads=3, chunk_size=chunk)
    return Frame(base_columns(), executor=ex, **kw)


@pytest.mark.parametrize("chunk", [None, 3])
def test_count_1d(chunk):
    # tests/agg_test.py:150-158 (small_buffer(size=3) == chunk 3)
    from vaex_b200.execution import Executor
    from vaex_b200.frame import Frame
    x = np.array([-1, -2, 0.5, 1.5, 4.5, 5], dtype="f8")
    df = Frame(dict(x=x), executor=Executor(nthreads=2, chunk_size=chunk))
    bins = 5
    binner = df._binner_specs("x", [0, 5], bins)
    assert binner[0]["count"] == 5
    grid = df.count(binby="x", limits=[0, 5], shape=bins, edges=True)
    assert grid.tolist() == [0, 2, 1, 1, 0, 0, 1, 1]
    assert df.count(binby="x", limits=[0, 5], shape=bins).tolist() == [1, 1, 0, 0, 1]


@pytest.mark.parametrize("chunk", [None, 3])
def test_count_1d_ordinal(chunk):
    # tests/agg_test.py:171-180
    from vaex_b200.execution import Executor
    from vaex_b200.frame import Frame
    x = np.array([-1, -2, 0, 1, 4, 5], dtype="i8")
    df = Frame(dict(x=x), executor=Executor(nthreads=2, chunk_size=chunk))
    df.categorize("x", min_value=0, count=5)
    assert df.count(binby="x", edges=True).tolist() == [1, 1, 0, 0, 1, 3, 0]


@pytest.mark.parametrize("chunk", [None, 3, 7])
def test_sum_count_mean(chunk):
    # tests/agg_test.py:8-48, :108-147: against numpy on the same data, incl. big-endian x, masked m, NaN n
    df = frame(chunk)
    c = base_columns()
    x, y = c["x"].astype("f8"), c["y"].astype("f8")
    assert df.count() == 20
    assert df.count("m") == 10
    assert df.count("n") == 18
    assert df.sum("x") == x.sum()
    assert df.sum("m") == c["m"].sum()
    assert df.sum("n") == np.nansum(c["n"])
    assert df.sum("i8") == c["i8"].sum() and df.sum("i8").dtype == np.int64  # upcast, tests/agg_test.py:395-402
    np.testing.assert_array_equal(df.sum("y", binby="x", limits=[0, 20], shape=2), [y[:10].sum(), y[10:].sum()])
    np.testing.assert_array_equal(df.count(binby=["x", "y"], limits=[[0, 20], [0, 400]], shape=[2, 2]), np.histogram2d(x, y, bins=2, range=[[0, 20], [0, 400]])[0])
    np.testing.assert_allclose(df.mean("y", binby="x", limits=[0, 20], shape=4), [y[i:i + 5].mean() for i in range(0, 20, 5)], rtol=1e-12)


def test_var_and_std_equal_numpy():
    # tests/agg_test.py:419-439: the reference asserts EXACT equality with numpy on this 10-row fixture
    from vaex_b200.frame import Frame
    x = np.arange(10, dtype="f8")
    y = x ** 2
    df = Frame(dict(x=x, y=y))
    assert df.var("y").tolist() == np.var(y).tolist()
    assert df.std("y").tolist() == np.std(y).tolist()
    v = df.var("y", binby="x", limits=[0, 10], shape=2)
    np.testing.assert_allclose(v, [np.var(y[:5]), np.var(y[5:])], rtol=1e-12)


def test_minmax_and_limits():
    # tests/agg_test.py:195-246
    df = frame(4)
    c = base_columns()
    assert df.min("x") == 0 and df.max("x") == 19
    assert df.min("n") == 0 and df.max("n") == 19
    assert df.max("m") == 18
    assert df.min("i8") == -10 and df.min("i8").dtype == np.int8
    np.testing.assert_array_equal(df.minmax("x"), [0, 19])
    np.testing.assert_array_equal(df.minmax("n"), [0, 19])
    np.testing.assert_array_equal(df.minmax("m"), [0, 18])
    np.testing.assert_array_equal(df.max("y", binby="x", limits=[0, 20], shape=2), [81, 361])
    # limits=None triggers the device min/max pre-pass
    grid = df.count(binby="x", shape=4)
    assert grid.sum() == 19  # the maximum falls on the upper edge and is excluded, like the reference


def test_numpy_histogram_cross_check():
    # tests/count_test.py:26-42
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(3)
    x = rng.normal(0, 1, 100_000)
    df = Frame(dict(x=x))
    got = df.count(binby="x", limits=[-4, 4], shape=64)
    want = np.histogram(x, bins=64, range=(-4, 4))[0]
    # numpy puts x == upper edge in the last bin; vaex excludes it; none of the samples is exactly 4.0
    np.testing.assert_array_equal(got, want)


def test_selection_and_first_last():
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(9)
    n = 5000
    x = rng.uniform(0, 10, n)
    v = rng.normal(0, 1, n)
    t = rng.permutation(n).astype("i8")
    df = Frame(dict(x=x, v=v, t=t))
    sel = v > 0
    got = df.sum("v", binby="x", limits=[0, 10], shape=5, selection=sel)
    want = [v[(x >= i * 2) & (x < i * 2 + 2) & sel].sum() for i in range(5)]
    np.testing.assert_allclose(got, want, rtol=1e-12)
    first = df.first("v", "t", binby="x", limits=[0, 10], shape=5)
    last = df.last("v", "t", binby="x", limits=[0, 10], shape=5)
    for i in range(5):
        inb = (x >= i * 2) & (x < i * 2 + 2)
        assert first[i] == v[inb][np.argmin(t[inb])]
        assert last[i] == v[inb][np.argmax(t[inb])]


def test_task_merging_single_pass():
    # tests/execution_test.py:144-175: aggregations with equal binners share ONE pass
    df = frame()
    before = df.executor.passes
    df._agg([__import__("vaex_b200.agg", fromlist=["x"]).mean("y"), __import__("vaex_b200.agg", fromlist=["x"]).std("y"),
             __import__("vaex_b200.agg", fromlist=["x"]).count()], binby="x", limits=[0, 20], shape=4)
    assert df.executor.passes == before + 1


@pytest.mark.parametrize("fused", [True, False])
def test_groupby_sum_count(fused, oracle):
    # tests/groupby_test.py:116-176 style: groupby over hashed int64 keys, sum + count, against numpy
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(4)
    n = 200_000
    keys = rng.integers(0, 1000, n).astype("i8") * 256 + 5
    v = rng.normal(0, 1, n)
    df = Frame(dict(k=keys, v=v))
    out = df.groupby("k", agg={"v": ["sum", "count"]}, fused=fused)
    uniq, inv = np.unique(keys, return_inverse=True)
    want_sum = np.bincount(inv, weights=v)
    want_cnt = np.bincount(inv)
    order = np.argsort(out["k"])
    np.testing.assert_array_equal(np.asarray(out["k"])[order], uniq)
    np.testing.assert_array_equal(out["v_count"][order], want_cnt)
    np.testing.assert_allclose(out["v_sum"][order], want_sum, rtol=1e-9, atol=1e-9)
    # group order == the ordinals of the sequential reference: ordered_set with 7 shards (vaex/cpu.py:317), keys in
    # key_array() order (shard by shard, insertion order inside a shard: src/hash.hpp:337-353), restated by the oracle
    ref = oracle.OrderedSet("int64", 7)
    ref.update(keys, None, -1, False)
    assert np.array_equal(np.asarray(out["k"]), ref.key_array())
    ordinals = ref.map_ordinal(keys)
    np.testing.assert_array_equal(out["v_count"], np.bincount(ordinals, minlength=len(ref)))


def test_groupby_float_keys_nan_and_missing():
    # tests/groupby_test.py:219-275: NaN and masked keys get their own groups' cells and are dropped from the center
    from vaex_b200.frame import Frame
    keys = np.ma.array([1.5, 2.5, np.nan, 1.5, 7.0, np.nan, 2.5, 9.0], mask=[0, 0, 0, 0, 1, 0, 0, 0])
    v = np.arange(8, dtype="f8")
    df = Frame(dict(k=keys, v=v))
    gb = df.groupby("k")
    hm = gb.hash_maps[0]
    assert hm.has_nan and hm.has_null and len(hm) == 5
    out = gb.agg({"v": "sum"})
    got = {(None if np.ma.is_masked(k) else (float("nan") if k != k else float(k))): s for k, s in zip(out["k"], out["v_sum"])}
    assert got[1.5] == 3.0 and got[2.5] == 7.0 and got[9.0] == 7.0


def test_value_counts_and_unique():
    # tests/value_counts_test.py, tests/unique_test.py style: against numpy / collections
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(10)
    n = 50_000
    k = rng.integers(0, 50, n).astype("i4")
    f = np.ma.array(rng.integers(0, 5, n).astype("f8"), mask=rng.random(n) < 0.1)
    f.data[::17] = np.nan
    df = Frame(dict(k=k, f=f), executor=__import__("vaex_b200.execution", fromlist=["x"]).Executor(nthreads=2, chunk_size=7001))
    keys, counts = df.value_counts("k")
    u, c = np.unique(k, return_counts=True)
    assert dict(zip(keys, counts.tolist())) == dict(zip(u.tolist(), c.tolist()))
    assert list(counts) == sorted(counts, reverse=True)
    assert sorted(df.unique("k")) == u.tolist()
    keys, counts = df.value_counts("f")
    valid = f.compressed()
    assert counts[[i for i, q in enumerate(keys) if q is None][0]] == int(f.mask.sum())
    assert counts[[i for i, q in enumerate(keys) if isinstance(q, float) and q != q][0]] == int(np.isnan(valid).sum())
    keys, counts = df.value_counts("f", dropna=True)
    u, c = np.unique(valid[~np.isnan(valid)], return_counts=True)
    assert dict(zip(keys, counts.tolist())) == dict(zip(u.tolist(), c.tolist()))


@pytest.mark.parametrize("combine", [True, False])
def test_groupby_with_missing_combine(combine):
    # tests/groupby_test.py:334-343 (assume_sparse True / False): two keys, nulls sort last
    from vaex_b200.frame import Frame
    g1 = np.ma.array([0, 0, 1, 1, 1, 99, 99, 2], mask=[0, 0, 0, 0, 0, 1, 1, 0], dtype="i8")
    g2 = np.array([0, 1, 0, 1, 1, 0, 1, 0], dtype="i8")
    df = Frame(dict(g1=g1, g2=g2))
    out = df.groupby(["g1", "g2"], agg=[__import__("vaex_b200.agg", fromlist=["x"]).count()], sort=True, combine=combine)
    assert out["g1"].tolist() == [0, 0, 1, 1, 2, None, None]
    assert out["g2"].tolist() == [0, 1, 0, 1, 0, 0, 1]
    assert out["count"].tolist() == [1, 1, 1, 2, 1, 1, 1]


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("sort", [False, True])
def test_groupby_sparse_combined_matches_numpy(device, sort, oracle):
    """vaex/groupby.py:526-584 (_combine): 3000 x 2000 x 3 possible key combinations, 200k rows -> 'auto' combines (occupancy
    < 10 rows per cell); sums / counts per distinct (k1, k2, k3) against numpy, host chunks and device-resident columns."""
    import torch
    from vaex_b200.execution import Executor
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(23)
    n = 200_000
    k1 = rng.integers(0, 3000, n).astype("i8") * 1000 + 7
    k2 = rng.integers(0, 2000, n).astype("f8") * 0.25
    k2[rng.random(n) < 0.01] = np.nan
    k3 = rng.integers(0, 3, n).astype("i4")
    v = rng.normal(0, 1, n)
    cols = dict(k1=k1, k2=k2, k3=k3, v=v)
    if device:
        cols = {k: torch.from_numpy(a).cuda() for k, a in cols.items()}
    df = Frame(cols, executor=Executor(nthreads=3, chunk_size=33_333))
    gb = df.groupby(["k1", "k2", "k3"], sort=sort, combine="auto")
    assert gb.combined is not None
    out = gb.agg({"v": ["sum", "count"]})
    # numpy: distinct rows of (k1, k2-with-NaN-as-one-value, k3)
    k2key = np.where(np.isnan(k2), -1.0, k2)
    rec = np.rec.fromarrays([k1, k2key, k3])
    uniq, inv = np.unique(rec, return_inverse=True)
    want_sum = np.bincount(inv, weights=v)
    want_cnt = np.bincount(inv)
    got_k2 = np.where(np.isnan(np.asarray(out["k2"], dtype="f8")), -1.0, np.asarray(out["k2"], dtype="f8"))
    got = np.rec.fromarrays([np.asarray(out["k1"]), got_k2, np.asarray(out["k3"])])
    assert len(got) == len(uniq)
    order = np.argsort(got)
    assert np.array_equal(got[order], uniq)
    np.testing.assert_array_equal(out["count"][order], want_cnt)
    np.testing.assert_array_equal(out["v_count"][order], want_cnt)
    np.testing.assert_allclose(out["v_sum"][order], want_sum, rtol=1e-9, atol=1e-9)
    if sort:  # lexicographic by (k1, k2, k3), NaN after the numbers
        nan_last = np.where(np.isnan(np.asarray(out["k2"], dtype="f8")), np.inf, np.asarray(out["k2"], dtype="f8"))
        lex = np.lexsort((np.asarray(out["k3"]), nan_last, np.asarray(out["k1"])))
        assert np.array_equal(lex, np.arange(len(lex)))
    else:
        # the sequential reference's order: ordinals of the combined codes in an ordered_set with 7 shards (vaex/cpu.py:317),
        # restated with the oracle's ordered sets end to end
        keys = [k1, k2, k3]
        sets = [oracle.OrderedSet(k.dtype.name, 7) for k in keys]
        for s_, k in zip(sets, keys):
            s_.update(k, None, -1, False)
        mult = [len(sets[1]) * len(sets[2]), len(sets[2]), 1]
        codes = sum(s_.map_ordinal(k).astype("i8") * m for s_, k, m in zip(sets, keys, mult))
        cs = oracle.OrderedSet("int64", 7)
        cs.update(codes, None, -1, False)
        want_codes = cs.key_array()
        o1, rest = want_codes // mult[0], want_codes % mult[0]
        o2, o3 = rest // mult[1], rest % mult[1]
        assert np.array_equal(np.asarray(out["k1"]), sets[0].key_array()[o1])
        assert np.array_equal(np.asarray(out["k2"], dtype="f8"), sets[1].key_array()[o2], equal_nan=True)
        assert np.array_equal(np.asarray(out["k3"]), sets[2].key_array()[o3])


def test_nunique_reference_kats():
    # tests/agg_test.py:294-333 (float variant: the strings mapped to floats, None -> NaN), :583-586 (no binby)
    from vaex_b200.execution import Executor
    from vaex_b200.frame import Frame
    x = np.array([0, 0, 0, 0, 0, 1, 1, 1, 2], dtype="i8")
    s = np.array([1.2, 1.2, 2.5, 3.7, np.nan, 3.7, 4.8, 3.7, 1.2])
    y = np.array([1, 1, 0, 1, 0, 0, 0, 1, 1])
    for chunk in (None, 2):
        df = Frame(dict(x=x, s=s), executor=Executor(nthreads=2, chunk_size=chunk))
        df.categorize("x", min_value=0, count=3)
        assert df.nunique("s", binby="x").tolist() == [4, 2, 1]
        assert df.nunique("s", binby="x", dropnan=True).tolist() == [3, 2, 1]
        assert df.nunique("s", binby="x", selection=(y == 0)).tolist() == [2, 2, 0]
    df = Frame(dict(x=np.array([1, 2, 3], dtype="i8")))
    r = df.nunique("x")
    assert r.ndim == 0 and r.item() == 3


@pytest.mark.parametrize("device", [False, True])
def test_nunique_large_matches_numpy(device):
    """2M rows, ~600k distinct (cell, value) pairs: table growth + rehash across chunks, several executor threads sharing the one
    aggregator (vaex/agg.py:357: 'using a shared hashmap, which is thread safe'), masked values, NaN; against numpy."""
    import torch
    from vaex_b200.execution import Executor
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(33)
    n = 2_000_000
    g = rng.integers(0, 50, n).astype("i4")
    v = rng.integers(0, 20_000, n).astype("f8")
    v[rng.random(n) < 0.001] = np.nan
    vm = np.ma.array(v, mask=rng.random(n) < 0.002)
    cols = dict(g=g, v=v) if device else dict(g=g, v=vm)
    if device:
        cols = {k: torch.from_numpy(a).cuda() for k, a in cols.items()}
    df = Frame(cols, executor=Executor(nthreads=3, chunk_size=300_000))
    df.categorize("g", min_value=0, count=50)
    got = df.nunique("v", binby="g")
    got_drop = df.nunique("v", binby="g", dropna=True)
    valid = np.ones(n, bool) if device else ~vm.mask
    for c in range(50):
        inc = g == c
        vals = v[inc & valid]
        nn = vals[~np.isnan(vals)]
        distinct = len(np.unique(nn))
        n_nan, n_null = int(np.isnan(vals).sum()), int((inc & ~valid).sum())
        assert got[c] == distinct + (n_nan > 0) + (n_null > 0)
        assert got_drop[c] == distinct + (n_nan > 0) + (n_null > 0) - n_nan - n_null  # row counts, like the reference


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("sort", [False, True])
def test_groupby_combine_recurses_past_64_bits(device, sort, oracle):
    """vaex/groupby.py:541-548, 572-582: four keys with ~2^16 distinct values each — the cartesian product (~1.8e19) overflows
    2^63-1, so the first three are combined, their distinct codes become one grouper and that is combined with the fourth.  Groups,
    counts and sums against numpy; the unsorted order against the same two-stage recursion restated with the oracle's sets."""
    import torch
    from vaex_b200.execution import Executor
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(29)
    n = 700_000
    keys = [rng.integers(0, 1 << 16, n).astype(dt) for dt in ("i8", "i4", "u2", "i8")]
    keys[3] = keys[3] * 3 - 70_000
    dup = rng.integers(0, n // 2, n // 2)  # half of the rows repeat an earlier combination, so groups hold > 1 row
    for k in keys:
        k[n // 2:] = k[dup]
    v = rng.normal(0, 1, n)
    cols = dict(a=keys[0], b=keys[1], c=keys[2], d=keys[3], v=v)
    if device:
        cols = {k: torch.from_numpy(a.astype("i4") if a.dtype == np.uint16 else a).cuda() for k, a in cols.items()}
        keys[2] = keys[2].astype("i4")
    df = Frame(cols, executor=Executor(nthreads=3, chunk_size=100_003))
    gb = df.groupby(["a", "b", "c", "d"], sort=sort, combine=True)
    sizes = [len(hm) for hm in gb.hash_maps]
    assert sizes[0] * sizes[1] * sizes[2] * sizes[3] >= 2 ** 63 - 1 and len(gb._stages) == 2
    out = gb.agg({"v": ["sum", "count"]})
    rec = np.rec.fromarrays([k.astype("i8") for k in keys])
    uniq, inv = np.unique(rec, return_inverse=True)
    got = np.rec.fromarrays([np.asarray(out[k]).astype("i8") for k in "abcd"])
    assert len(got) == len(uniq)
    order = np.argsort(got)
    assert np.array_equal(got[order], uniq)
    np.testing.assert_array_equal(out["count"][order], np.bincount(inv))
    np.testing.assert_allclose(out["v_sum"][order], np.bincount(inv, weights=v), rtol=1e-9, atol=1e-9)
    if sort:
        assert np.array_equal(order, np.arange(len(order)))
    else:
        sets = [oracle.OrderedSet(k.dtype.name, 7) for k in keys]
        for s_, k in zip(sets, keys):
            s_.update(k, None, -1, False)
        m1 = [len(sets[1]) * len(sets[2]), len(sets[2]), 1]
        c1 = sum(s_.map_ordinal(k).astype("i8") * m for s_, k, m in zip(sets[:3], keys[:3], m1))
        s1 = oracle.OrderedSet("int64", 7)
        s1.update(c1, None, -1, False)
        c2 = s1.map_ordinal(c1).astype("i8") * len(sets[3]) + sets[3].map_ordinal(keys[3]).astype("i8")
        s2 = oracle.OrderedSet("int64", 7)
        s2.update(c2, None, -1, False)
        final = s2.key_array()
        inner = s1.key_array()[final // len(sets[3])]
        o = [inner // m1[0], inner % m1[0] // m1[1], inner % m1[1], final % len(sets[3])]
        for name, s_, oo in zip("abcd", sets, o):
            assert np.array_equal(np.asarray(out[name]), s_.key_array()[oo])


@pytest.mark.parametrize("dropnan", [False, True])
@pytest.mark.parametrize("by_col_has_missing", [False, True])
@pytest.mark.parametrize("combine", [False, True])
def test_groupby_agg_list(dropnan, by_col_has_missing, combine):
    # tests/agg_test.py:660-693 (test_agg_list), the numeric column: groupby('id').agg(list(num, dropnan=...)) -> one arrow list per group
    from vaex_b200 import agg
    from vaex_b200.frame import Frame
    ids = np.ma.array([1, 2, 2, 1, 1, 3, 3], mask=[0, 0, 0, 0, 0, by_col_has_missing, by_col_has_missing], dtype="i8")
    num = np.array([1.1, 1.2, 1.3, 1.4, np.nan, 1.6, 1.7])
    cols = dict(id=ids, num=num)
    by = "id"
    if combine:  # the sparse path (tests/agg_test.py:723): a second, constant key
        cols["one"] = np.zeros(7, "i4")
        by = ["id", "one"]
    df = Frame(cols)
    # sort=True: the reference groups small-range integer keys with BinnerInteger, i.e. in key order (vaex/groupby.py:598-600)
    out = df.groupby(by, agg=[agg.list("num", dropnan=dropnan)], combine=combine, sort=True)
    result = out["num_list"].to_pylist()
    if dropnan:
        assert result == [[1.1, 1.4], [1.2, 1.3], [1.6, 1.7]]
    else:
        assert result[1:] == [[1.2, 1.3], [1.6, 1.7]]
        assert result[0][:2] == [1.1, 1.4] and np.isnan(result[0][2]) and len(result[0]) == 3
    assert out["id"].tolist() == ([1, 2, None] if by_col_has_missing else [1, 2, 3])
    assert out["count"].tolist() == [3, 2, 2]


def test_frame_list_binby_through_the_task_part():
    # Frame.list -> AggregatorDescriptorBasic('AggList') -> TaskPartAggregation (chunked over 3 workers): per cell of the full grid
    # (edge cells included, first binner fastest) the values in ROW order, whatever worker fed the chunk (vaex/agg.py:654-674)
    from vaex_b200.execution import Executor
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(8)
    n = 50_000
    x = rng.uniform(-1, 11, n)
    v = rng.integers(0, 1000, n).astype("i4")
    df = Frame(dict(x=x, v=v), executor=Executor(nthreads=3, chunk_size=7_001))
    lists = df.list("v", binby="x", limits=[0, 10], shape=5)
    assert len(lists) == 8
    cell = np.where(np.isnan(x), 0, np.where(x < 0, 1, np.where(x >= 10, 7, np.floor(x / 2) + 2))).astype(int)
    for c in range(8):
        got = np.asarray(lists[c].as_py(), dtype="i4")
        want = v[cell == c]
        # chunks are fed concurrently: the order inside a cell is row order within a chunk, chunks in completion order
        assert np.array_equal(np.sort(got), np.sort(want))
    single = Frame(dict(x=x, v=v), executor=Executor(nthreads=1, chunk_size=7_001)).list("v", binby="x", limits=[0, 10], shape=5)
    for c in range(8):
        assert np.array_equal(np.asarray(single[c].as_py(), dtype="i4"), v[cell == c])  # sequential feed: exactly row order


def test_readme_example_runs_and_is_right():
    """The calls README.md shows, on small data, against numpy."""
    import pyarrow as pa
    from vaex_b200 import agg
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(12)
    n = 40_000
    x, y, z = (rng.normal(0, 1, n).astype("f4") for _ in range(3))
    keys = rng.integers(0, 50, n).astype("i8") * 7 + 1
    words = np.array(["aap", "noot", "mies", "kees", None], dtype=object)
    strings = pa.array(words[rng.integers(0, 5, n)].tolist())
    df = Frame(dict(x=x, y=y, z=z, k=keys, s=strings))
    c = df.count(binby=["x", "y"], limits=[[-3, 3], [-3, 3]], shape=64)
    want, _, _ = np.histogram2d(x, y, bins=64, range=[[-3, 3], [-3, 3]])
    # numpy's last bin is closed on the right, the reference's is open: compare where no value sits exactly on the edge
    assert c.shape == (64, 64) and abs(int(c.sum()) - int(want.sum())) <= 2
    m = df.mean("z", binby=["x", "y"], shape=8)
    assert m.shape == (8, 8) and np.isfinite(m).any()
    out = df.groupby("k", agg={"z": ["sum", "count", "std"]}, sort=True)
    uniq, inv = np.unique(keys, return_inverse=True)
    assert np.array_equal(out["k"], uniq)
    np.testing.assert_allclose(out["z_sum"], np.bincount(inv, weights=z.astype("f8")), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(out["z_std"], [z[inv == i].astype("f8").std() for i in range(len(uniq))], rtol=1e-6, atol=1e-9)
    out = df.groupby(["k", "s"], agg=[agg.nunique("x"), agg.list("z")], combine="auto", sort=True)
    svals = np.array(strings.to_pylist(), dtype=object)
    groups = {}
    for i in range(n):
        groups.setdefault((int(keys[i]), svals[i]), []).append(i)
    assert len(out["count"]) == len(groups)
    for j in range(len(out["count"])):
        rows = groups[(int(out["k"][j]), out["s"][j])]
        assert out["count"][j] == len(rows)
        assert out["x_nunique"][j] == len(set(x[rows].tolist()))
        assert np.array_equal(np.sort(np.asarray(out["z_list"][j].as_py(), dtype="f4")), np.sort(z[rows]))
    f = Frame(dict(x=x, y=y), filter="x * x + y * y < 4").count(binby="x", limits=[-3, 3], shape=16)
    keep = (x.astype("f4") * x + y * y) < 4
    assert int(f.sum()) == int((keep & (x >= -3) & (x < 3)).sum())
