"""GPU tests of the remaining native-class protocol: merge, initial values / load, buffer protocol shape, reset, device-resident
Frame columns, multi-threaded host-chunk feeding (concurrent process() on one shared part, like the reference executor)."""
import numpy as np
import pytest

from helpers import same

pytestmark = pytest.mark.gpu


def _mk(superagg, x, y=None, bins=8):
    b = [superagg.BinnerScalar_float64(2, "x", -3.0, 3.0, bins)]
    if y is not None:
        b.append(superagg.BinnerScalar_float64(2, "y", -3.0, 3.0, bins))
    return b, superagg.Grid(b)


def test_merge_equals_single_pass(oracle):
    """Aggregator.merge (src/agg_count.cpp:15-23, agg_sum.cpp:69-76, agg_minmax.cpp:19-27) + the first/last fold
    (src/agg_first.cpp:68-99): two objects fed with halves of the rows, merged, equal one object fed with everything"""
    from vaex_b200 import superagg
    rng = np.random.default_rng(2)
    n = 20000
    x, v = rng.normal(0, 1, n), rng.normal(0, 1, n)
    order = rng.integers(0, 100, n).astype("i8")
    want = oracle.binby([oracle.scalar(x, -3, 3, 8)], [oracle.agg("count", v), oracle.agg("sum", v), oracle.agg("min", v), oracle.agg("max", v),
                                                        oracle.agg("first", v, None, order=order), oracle.agg("last", v, None, order=order)], n)
    halves = []
    for i1, i2 in ((0, n // 2), (n // 2, n)):
        b, g = _mk(superagg, x)
        aggs = [superagg.AggCount_float64(g, 1, 2), superagg.AggSum_float64(g, 1, 2), superagg.AggMin_float64(g, 1, 2), superagg.AggMax_float64(g, 1, 2),
                superagg.AggFirst_float64_int64(g, 1, 2, False), superagg.AggFirst_float64_int64(g, 1, 2, True)]
        b[0].set_data(0, x[i1:i2])
        for a in aggs:
            a.set_data(0, v[i1:i2], 0)
        for a in aggs[4:]:
            a.set_data(0, order[i1:i2], 1)
        g.bin(0, aggs, i2 - i1, row_offset=i1)
        halves.append(aggs)
    for a0, a1, w in zip(halves[0], halves[1], want):
        a0.merge([a1])
        g = a0.get_result()
        if a0._op == 1:
            assert np.allclose(g, w, rtol=1e-12, atol=1e-12)
        else:
            assert same(w, g), type(a0).__name__


def test_buffer_protocol_load_and_reset():
    from vaex_b200 import superagg
    x = np.array([-1.0, 0.5, 0.6, 2.0])
    b, g = _mk(superagg, x, bins=4)
    a = superagg.AggCount_float64(g, 3, 4)
    b[0].set_data(0, x)
    g.bin(0, [a], 4)
    full = np.asarray(a)  # (grids, *shapes) like src/agg_base.hpp:106-125
    assert full.shape == (3, 7) and full[0].sum() == 4 and full[1:].sum() == 0
    import sys
    assert sys.getsizeof(a) >= 3 * 7 * 8 and a.__sizeof__() == 3 * 7 * 8  # the reference's bytes_used() figure
    c = superagg.AggCount_float64(g, 3, 4)
    c.load(full)  # TaskPartAggregation initial_values (vaex/cpu.py:654-658)
    g.bin(0, [c], 4)
    assert np.array_equal(c.get_result(), 2 * a.get_result())
    a.reset()
    assert a.get_result().sum() == 0
    m = superagg.AggMin_float64(g, 1, 1)
    m.set_data(0, x, 0)
    g.bin(0, [m], 4)
    m.reset()
    assert np.all(np.isinf(m.get_result()))


def test_frame_over_device_tensors_matches_host():
    import torch
    from vaex_b200.frame import Frame
    rng = np.random.default_rng(6)
    n = 300_000
    cols = dict(x=rng.normal(0, 1, n).astype("f4"), y=rng.normal(0, 1, n).astype("f4"), v=rng.normal(0, 1, n))
    host = Frame(cols, nthreads=4)
    dev = Frame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    kw = dict(binby=["x", "y"], limits=[[-3, 3], [-3, 3]], shape=[64, 32])
    assert np.array_equal(host.count(**kw), dev.count(**kw))
    assert np.array_equal(host.count("v", **kw), dev.count("v", **kw))
    assert np.allclose(host.mean("v", **kw), dev.mean("v", **kw), rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.array_equal(host.max("v", **kw), dev.max("v", **kw))
    np.testing.assert_allclose(host.minmax("v"), dev.minmax("v"))
    # limits=None on device columns: device min/max pre-pass feeds the binner
    assert np.array_equal(host.count(binby="x", shape=16), dev.count(binby="x", shape=16))


def test_concurrent_process_on_one_shared_part(oracle):
    """the reference executor calls process() from nthreads Python threads on ONE task part (vaex/execution.py:404-406)"""
    from concurrent.futures import ThreadPoolExecutor
    import threading
    from vaex_b200 import taskpart
    rng = np.random.default_rng(8)
    n = 400_000
    x = rng.normal(0, 1, n).astype("f4")
    v = rng.integers(-100, 100, n).astype("i4")
    spec = {"binners": [{"binner-type": "scalar", "expression": "x", "dtype": "float32", "count": 32, "minimum": -3.0, "maximum": 3.0}],
            "aggregations": [{"aggregation": "count"}, {"aggregation": "sum", "expressions": ["v"]}, {"aggregation": "max", "expressions": ["v"], "edges": True}],
            "dtypes": {"x": "float32", "v": "int32"}}
    nthreads = 6
    part = taskpart.TaskPartAggregation.decode(None, spec, df=None, nthreads=nthreads)
    assert part.ideal_splits(nthreads) == 1 and part.expressions == ["x", "v", "v"] and part.get_bin_count() == 35
    local, lock, counter = threading.local(), threading.Lock(), [0]

    def work(r):
        if not hasattr(local, "i"):
            with lock:
                local.i = counter[0]
                counter[0] += 1
        i1, i2 = r
        part.process(local.i, i1, i2, None, [None, None, None], [x[i1:i2], v[i1:i2], v[i1:i2]])

    chunk = 17_001
    with ThreadPoolExecutor(nthreads) as pool:
        list(pool.map(work, [(i, min(i + chunk, n)) for i in range(0, n, chunk)]))
    part.reduce([])
    count, total, vmax = part.get_result()
    want = oracle.binby([oracle.scalar(x, -3, 3, 32)], [oracle.agg("count"), oracle.agg("sum", v), oracle.agg("max", v)], n)
    assert np.array_equal(count, want[0][2:-1]) and np.array_equal(total, want[1][2:-1]) and np.array_equal(vmax, want[2])
    assert part.memory_usage() > 0


def test_binner_combined_and_map_many(oracle):
    """BinnerCombined (src/binner_combined.cpp:25-29: member indices composed with strides 1, shape_0, ...) and hash_map::map_many
    (src/hash_primitives.hpp:567-590: global ordinals as int64, NaN -> NaN ordinal, absent -> -1)"""
    import pickle
    from vaex_b200 import superagg, superutils
    rng = np.random.default_rng(12)
    n = 20_000
    x, y = rng.standard_normal(n), rng.standard_normal(n).astype("f4")
    z = rng.integers(0, 4, n).astype("i4")
    bx, by = superagg.BinnerScalar_float64(1, "x", -3, 3, 10), superagg.BinnerScalar_float32(1, "y", -2, 2, 5)
    bz = superagg.BinnerOrdinal_int32(1, "z", 4, 0, False, False)
    comb = superagg.BinnerCombined(1, [bx, by])
    assert len(comb) == len(by) and comb.strides == [1, len(bx)] and len(comb.copy().binners) == 2
    assert [type(b).__name__ for b in pickle.loads(pickle.dumps(comb)).binners] == ["BinnerScalar_float64", "BinnerScalar_float32"]
    grid = superagg.Grid([comb, bz])
    assert grid.shapes == [13, 8, 6] and grid.strides == [1, 13, 104]
    agg = superagg.AggCount_int64(grid, 1, 1)
    for b, a in ((bx, x), (by, y), (bz, z)):
        b.set_data(0, a)
    grid.bin(0, [agg], n)
    want = oracle.binby([oracle.scalar(x, -3, 3, 10), oracle.scalar(y, -2, 2, 5), oracle.ordinal(z, 4)], [oracle.agg("count")])[0]
    assert np.array_equal(agg.get_result(), want)
    keys = rng.integers(0, 50, n).astype("f8") * 0.5
    keys[::17] = np.nan
    s = superutils.ordered_set_float64(3)
    s.update(keys[: n // 2], -1)
    so = oracle.OrderedSet("float64", 3)
    so.update(keys[: n // 2], None, -1, False)
    probe = np.concatenate([keys, [1e9, -7.25]])
    out = np.full(100, -99, np.int64)
    s.map_many(probe, len(probe) - 100, 100, out)
    assert np.array_equal(out, so.map_ordinal(probe)[-100:].astype(np.int64))


def test_agg_list_matches_the_reference_semantics():
    """AggList_<dtype>_int64 (src/agg_list.cpp:84-113 aggregate, :47-83 get_result): per cell the non-NaN values of the valid rows in
    arrival order, then one NaN per NaN value (unless dropnan), then one slot per null row (unless dropnull); several calls append."""
    from vaex_b200 import superagg
    rng = np.random.default_rng(21)
    n = 30_000
    x = rng.integers(0, 6, n).astype("i4")
    y = rng.standard_normal(n)
    v = rng.standard_normal(n).astype("f4")
    v[rng.random(n) < 0.1] = np.nan
    valid = (rng.random(n) < 0.85).astype("u1")
    for dropnan, dropnull in ((False, False), (True, False), (False, True), (True, True)):
        bx = superagg.BinnerOrdinal_int32(1, "x", 6, 0, False, False)
        by = superagg.BinnerScalar_float64(1, "y", -1, 1, 3)
        grid = superagg.Grid([bx, by])
        agg = superagg.AggList_float32_int64(grid, 1, 1, dropnan, dropnull)
        cells = len(grid)
        want = [[] for _ in range(cells)]
        nans, nulls = np.zeros(cells, int), np.zeros(cells, int)
        for i1, i2 in ((0, 9_001), (9_001, n)):  # two calls: the lists keep growing in call order
            bx.set_data(0, x[i1:i2])
            by.set_data(0, y[i1:i2])
            agg.set_data(0, v[i1:i2], 0)
            agg.set_data_mask(0, valid[i1:i2])
            grid.bin(0, [agg], i2 - i1)
        seen = np.concatenate([valid[i1:i2][np.arange(i2 - i1) % 1024] for i1, i2 in ((0, 9_001), (9_001, n))])  # see below
        for i in range(n):
            s = (y[i] + 1) * 0.5
            cy = 1 if s < 0 else 5 if s >= 1 else int(s * 3) + 2
            c = int(x[i]) + 8 * cy
            # reference quirk (src/agg_list.cpp:96-97, pinned by tests/golden/agglist_golden.npz): row r of a bin() call is judged by
            # mask[r % 1024] of that call — the 1024-row block offset is applied to the data, not to the mask
            if seen[i] == 1:
                if v[i] == v[i]:
                    want[c].append(v[i])
                elif not dropnan:
                    nans[c] += 1
            elif not dropnull:
                nulls[c] += 1
        offsets, values = agg.result_arrays()
        assert len(offsets) == cells + 1 and offsets[-1] == len(values)
        for c in range(cells):
            got = values[offsets[c]:offsets[c + 1]]
            k = len(want[c])
            assert len(got) == k + nans[c] + nulls[c], (c, dropnan, dropnull)
            assert np.array_equal(got[:k], np.array(want[c], "f4"))
            assert np.isnan(got[k:k + nans[c]]).all()
        assert agg.get_result().type.value_type == "float" and len(agg.get_result()) == cells
