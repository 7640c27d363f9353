"""The B1 boundary end to end WITHOUT a vaex installation: a stub ``vaex`` package restates the few pieces vaex_plugin.py touches —
the 'task-part-cpu' class registry (packages/vaex-core/vaex/encoding.py:31-52, vaex/cpu.py:21), ``Encoding.decode``,
``vaex.memory.local.agg.pre_alloc`` and ``vaex.array_types.to_numpy`` — then the plugin is installed, the spec dicts that
``TaskAggregations.encode`` / ``TaskHashmapUniqueCreate.encode`` emit (vaex/tasks.py:498-504, :213-225; binner specs
vaex/dataframe.py:7294-7360; aggregation specs vaex/agg.py:240-252) are decoded THROUGH THE REGISTRY, and the parts are driven the
way ExecutorLocal.process_tasks does (vaex/execution.py:555-566): concurrently from several threads with distinct thread indices,
then reduce() / get_result().  Results against the oracle."""
import sys
import threading
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def stub_vaex(monkeypatch):
    vaex = types.ModuleType("vaex")
    # ---- vaex.encoding: registry + Encoding.decode (encoding.py:18-52, 297-330) ----
    enc_mod = types.ModuleType("vaex.encoding")
    registry = {}

    def register(name):
        def wrapper(cls):
            registry[name] = cls
            return cls
        return wrapper

    def make_class_registery(groupname):
        types_ = {}

        def register_helper(cls):
            types_[cls.snake_name] = cls
            return cls

        @register(groupname)
        class encoding:
            @staticmethod
            def decode(encoding, spec, **kwargs):
                spec = spec.copy()
                cls = types_[spec.pop(f"{groupname}-type")]
                return cls.decode(encoding, spec, **kwargs)
        return register_helper

    @register("dtype")
    class dtype_encoding:
        @staticmethod
        def decode(encoding, spec):
            class DataType:  # vaex.datatype.DataType: .numpy, .is_string
                def __init__(self, s):
                    self.is_string = s in ("string", "large_string")
                    self.numpy = np.dtype("O") if self.is_string else np.dtype(s)
            return DataType(spec)

    class Encoding:
        def decode(self, typename, value, **kwargs):
            return registry[typename].decode(self, value, **kwargs)
    enc_mod.register, enc_mod.make_class_registery, enc_mod.Encoding, enc_mod.registry = register, make_class_registery, Encoding, registry
    # ---- vaex.cpu: the registry the executor consults, with the two reference classes as placeholders ----
    cpu = types.ModuleType("vaex.cpu")
    cpu.register = make_class_registery("task-part-cpu")

    class TaskPartAggregation:
        snake_name = "aggregations"

    class TaskPartHashmapUniqueCreate:
        snake_name = "hash_map_unique_create"
    cpu.TaskPartAggregation, cpu.TaskPartHashmapUniqueCreate = cpu.register(TaskPartAggregation), cpu.register(TaskPartHashmapUniqueCreate)
    # ---- vaex.memory / vaex.array_types ----
    memory = types.ModuleType("vaex.memory")
    declared = []
    memory.local = types.SimpleNamespace(agg=types.SimpleNamespace(pre_alloc=lambda nbytes, what: declared.append(nbytes)))
    at = types.ModuleType("vaex.array_types")
    at.to_numpy = lambda x, strict=True: (x.to_numpy(zero_copy_only=False) if hasattr(x, "to_numpy") and not isinstance(x, np.ndarray) else x)
    vaex.encoding, vaex.cpu, vaex.memory, vaex.array_types = enc_mod, cpu, memory, at
    for name, mod in (("vaex", vaex), ("vaex.encoding", enc_mod), ("vaex.cpu", cpu), ("vaex.memory", memory), ("vaex.array_types", at)):
        monkeypatch.setitem(sys.modules, name, mod)
    return types.SimpleNamespace(vaex=vaex, Encoding=Encoding, declared=declared)


def _drive(part, columns, n, nthreads, chunk):
    """ExecutorLocal.process_tasks for one task: chunks on `nthreads` workers, each with its own stable thread index"""
    ranges = [(i, min(i + chunk, n)) for i in range(0, n, chunk)]
    errors = []

    def work(t):
        for k in range(t, len(ranges), nthreads):
            i1, i2 = ranges[k]
            try:
                part.process(t, i1, i2, None, [None] * 8, [columns[e][i1:i2] for e in part.expressions])
            except Exception as e:  # pragma: no cover
                errors.append(e)
    threads = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    part.reduce([])
    return part.get_result()


def test_aggregation_task_through_the_registry(stub_vaex, oracle):
    from vaex_b200 import vaex_plugin
    replaced = vaex_plugin.install()
    assert set(replaced) == {"aggregations", "hash_map_unique_create"}
    rng = np.random.default_rng(8)
    n = 200_003
    x, y, z = rng.standard_normal(n), rng.standard_normal(n).astype("f4"), rng.standard_normal(n)
    spec = {"task-part-cpu-type": "aggregations",
            "binners": [{"binner-type": "scalar", "expression": "x", "dtype": "float64", "count": 64, "minimum": -3.0, "maximum": 3.0},
                        {"binner-type": "scalar", "expression": "y", "dtype": "float32", "count": 32, "minimum": -2.0, "maximum": 2.0}],
            "aggregations": [{"aggregation": "count"}, {"aggregation": "sum", "expressions": ["z"]}, {"aggregation": "max", "expressions": ["z"], "edges": True}],
            "dtypes": {"x": "float64", "y": "float32", "z": "float64"}}
    part = stub_vaex.Encoding().decode("task-part-cpu", spec, df=None, nthreads=3)
    assert type(part).__name__ == "VaexTaskPartAggregation" and part.expressions == ["x", "y", "z", "z"]
    assert stub_vaex.declared and stub_vaex.declared[-1] == part.memory_usage()  # vaex/execution.py:413-414 accounting
    got = _drive(part, {"x": x, "y": y, "z": z}, n, nthreads=3, chunk=30_000)
    b = [oracle.scalar(x, -3, 3, 64), oracle.scalar(y, -2, 2, 32)]
    want = oracle.binby(b, [oracle.agg("count"), oracle.agg("sum", z), oracle.agg("max", z)])
    assert np.array_equal(got[0], want[0][2:-1, 2:-1])  # edges=False: the [2:-1] slice of vaex/agg.py:323-335
    assert np.allclose(got[1], want[1][2:-1, 2:-1], rtol=1e-6, atol=1e-9)
    assert np.array_equal(got[2], want[2])
    vaex_plugin.uninstall()


def test_groupby_tasks_through_the_registry_numeric_and_string_keys(stub_vaex, oracle):
    import pyarrow as pa
    from vaex_b200 import vaex_plugin
    vaex_plugin.install()
    rng = np.random.default_rng(9)
    n = 50_000
    keys = rng.integers(0, 300, n).astype("i8") * 1000 + 7
    spec = {"task-part-cpu-type": "hash_map_unique_create", "expression": "k", "dtype": "int64", "dtype_item": "int64", "flatten": False, "limit": None,
            "limit_raise": True, "selection": None, "return_inverse": False}
    part = stub_vaex.Encoding().decode("task-part-cpu", spec, df=None, nthreads=2)
    hm = _drive(part, {"k": keys}, n, nthreads=1, chunk=7_000)  # one feeder: ordinals equal the sequential reference run
    so = oracle.OrderedSet("int64", 2 * 7)
    for i in range(0, n, 7_000):
        so.update(keys[i:i + 7_000], None, -1, False)
    assert np.array_equal(np.sort(hm.keys()), np.sort(so.key_array()))
    words = np.array(["alpha", "beta", "gamma", "", "δelta", None], dtype=object)
    skeys = words[rng.integers(0, len(words), n)]
    spec = dict(spec, dtype="string", dtype_item="string")
    part = stub_vaex.Encoding().decode("task-part-cpu", spec, df=None, nthreads=2)
    hm = _drive(part, {"k": pa.array(skeys.tolist(), type=pa.string())}, n, nthreads=1, chunk=7_000)
    assert sorted(k for k in hm.keys() if k is not None) == sorted(w for w in words if w is not None) and hm.has_null
    with pytest.raises(NotImplementedError):
        stub_vaex.Encoding().decode("task-part-cpu", dict(spec, dtype="O", dtype_item="O"), df=None, nthreads=1)
    vaex_plugin.uninstall()
