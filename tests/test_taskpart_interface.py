"""CPU-side shape checks of boundary B1: the B200 task parts expose the interface the reference executor drives
(packages/vaex-core/vaex/cpu.py:629-845, :285-405; calls at vaex/execution.py:385-412, :564, :451-453) and encode the same
spec keys vaex's descriptors do (vaex/agg.py:240-252).  No GPU work happens here."""
import inspect

import numpy as np
import pytest


def test_task_part_methods_and_signatures():
    from vaex_b200 import taskpart
    for cls in (taskpart.TaskPartAggregation, taskpart.TaskPartHashmapUniqueCreate):
        for name in ("decode", "process", "reduce", "get_result", "ideal_splits", "memory_usage", "get_bin_count"):
            assert callable(getattr(cls, name)), (cls.__name__, name)
        assert list(inspect.signature(cls.process).parameters)[1:] == ["thread_index", "i1", "i2", "filter_mask", "selection_masks", "blocks"]
        assert list(inspect.signature(cls.decode).parameters)[:2] == ["encoding", "spec"]
        assert cls.stopped is False
    assert taskpart.TaskPartAggregation.snake_name == "aggregations"                # vaex/cpu.py:631
    assert taskpart.TaskPartHashmapUniqueCreate.snake_name == "hash_map_unique_create"  # vaex/cpu.py:287
    assert set(taskpart.REGISTRY) == {"aggregations", "hash_map_unique_create"}


def test_descriptor_specs_match_vaex_encoding():
    from vaex_b200 import agg
    assert agg.count().encode() == {"aggregation": "count"}
    assert agg.sum("x", selection="s", edges=True).encode() == {"aggregation": "sum", "expressions": ["x"], "selection": "s", "edges": True}
    assert agg._sum_moment("x", 2).encode() == {"aggregation": "_sum_moment", "expressions": ["x"], "parameters": [2]}
    assert agg.first("x", "t").encode() == {"aggregation": "first", "expressions": ["x", "t"]}  # agg_args not encoded for first/last
    for spec in (agg.count("x").encode(), agg.min("x").encode(), agg.last("x", "t").encode(), agg._sum_moment("x", 3).encode()):
        d = agg.from_spec(spec)
        assert d.encode() == spec
    # mean / std decompose into the same primitives as vaex/agg.py:386-455
    assert [p.short_name for p in agg.mean("x").primitives()] == ["sum", "count"]
    assert [p.short_name for p in agg.std("x").primitives()] == ["_sum_moment", "sum", "count"]
    # vaex/agg.py:344-350, 600-612: dropna sets both flags; flags are only encoded when set
    assert agg.nunique("x").encode() == {"aggregation": "nunique", "expressions": ["x"]}
    spec = {"aggregation": "nunique", "expressions": ["x"], "dropmissing": True, "dropnan": True}
    assert agg.nunique("x", dropna=True).encode() == spec and agg.from_spec(spec).encode() == spec
    spec = {"aggregation": "list", "expressions": ["x"], "parameters": [True, False]}  # vaex/agg.py:240-252: agg_args travel as "parameters"
    assert agg.from_spec(spec).encode() == spec and agg.list("x", dropnan=True).encode() == spec
    with pytest.raises(ValueError):
        agg.from_spec({"aggregation": "describe", "expressions": ["x"]})


def test_prepare_types_and_class_lookup():
    from vaex_b200 import agg, superagg
    d = agg.sum("x")
    d._prepare_types({"x": np.dtype("float32")})
    assert d.dtype_out == np.float64
    d = agg.sum("i")
    d._prepare_types({"i": np.dtype("uint8")})
    assert d.dtype_out == np.uint64
    c = agg.count()
    c._prepare_types({})
    assert c.dtype_in == np.int64
    assert agg.find_type_from_dtype(superagg, "AggSum_", np.dtype(">f8")) is superagg.AggSum_float64_non_native
    assert agg.find_type_from_dtype(superagg, "AggFirst_", np.dtype("f4"), np.dtype("i8")) is superagg.AggFirst_float32_int64
    assert agg.find_type_from_dtype(superagg, "BinnerOrdinal_", np.dtype("bool")) is superagg.BinnerOrdinal_bool


def test_multi_finish_math_matches_reference_formulas():
    from vaex_b200 import agg
    rng = np.random.default_rng(0)
    v = rng.normal(3, 2, 1000)
    n, s1, s2, s3, s4 = len(v), v.sum(), (v ** 2).sum(), (v ** 3).sum(), (v ** 4).sum()
    assert np.isclose(agg.mean("v").combine(s1, n), v.mean())
    assert np.isclose(agg.var("v").combine(s2, s1, n), v.var())          # E[x^2] - E[x]^2 (vaex/agg.py:448-450)
    assert np.isclose(agg.std("v").combine(s2, s1, n), v.std())
    m = v.mean()
    assert np.isclose(agg.skew("v").combine(s1, s2, s3, n), ((v - m) ** 3).mean() / v.var() ** 1.5)
    assert np.isclose(agg.kurtosis("v").combine(s1, s2, s3, s4, n), ((v - m) ** 4).mean() / v.var() ** 2 - 3.0)


def test_executor_chunking_matches_reference_rule():
    from vaex_b200.execution import Executor
    ex = Executor(nthreads=8, chunk_size_max=1024 ** 2)  # vaex/execution.py:283-292 with the reference's cap
    assert ex.chunk_size_for(10) == 1024
    assert ex.chunk_size_for(8 * 5000) == 5000
    assert ex.chunk_size_for(10 ** 9) == 1024 ** 2
    assert Executor(nthreads=8, chunk_size=3).chunk_size_for(10 ** 6) == 3


def test_plugin_install_fails_cleanly_without_vaex():
    from vaex_b200 import taskpart
    try:
        import vaex  # noqa: F401
    except Exception:
        with pytest.raises(Exception):
            taskpart.install_into_vaex()


def test_combined_codes_decode_matches_reference_div_mod_chain():
    """GrouperCombined recovers the parents' ordinals with a div/mod chain over cumulative_counts[1:]
    (vaex/groupby.py:352-358; the docstring example there: N = 10 and 4 -> cumulative counts [40, 4, 1])."""
    from vaex_b200.hash import CombinedCodes
    cc = CombinedCodes([np.zeros(3), np.zeros(3)], [None, None], [4, 1])
    o0, o1 = np.array([0, 9, 3, 7]), np.array([3, 0, 2, 1])
    d0, d1 = cc.decode(o0 * 4 + o1)
    assert d0.tolist() == o0.tolist() and d1.tolist() == o1.tolist()
    cc = CombinedCodes([np.zeros(1)] * 3, [None] * 3, [35, 7, 1])  # N = (_, 5, 7)
    d = cc.decode(np.array([2 * 35 + 4 * 7 + 6]))
    assert [int(x[0]) for x in d] == [2, 4, 6]
    assert len(cc) == 1 and cc.dtype == np.int64 and cc.device_virtual
