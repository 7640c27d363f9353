"""vaex_b200 — B200-native binned statistics / groupby aggregation behind vaex's native-module interface.

Only what the hot path needs lives here:
  csrc/           hand-written sm_100a CUDA kernels + the C ABI (include/b200agg.h) -> libb200agg.so
  _lib.py         ctypes binding (no torch types cross the boundary)
  superagg.py     mirror of ``vaex.superagg``   (Binner*/Grid/Agg* class protocol)
  superutils.py   mirror of ``vaex.superutils`` (ordered_set_<dtype>, hash)
  hash.py         mirror of ``vaex.hash.HashMapUnique``
  taskpart.py     TaskPartAggregation / TaskPartHashmapUniqueCreate drop-ins for the 'task-part-cpu' registry
  engine.py       device-resident, row-sharded driver (one process per GPU, NCCL all-reduce of the grids)

There is no CPU fallback anywhere in this package.
"""
from . import _lib  # noqa: F401
from ._lib import build, context  # noqa: F401

__all__ = ["build", "context", "superagg", "superutils", "hash", "taskpart", "engine"]


def __getattr__(name):
    if name in ("superagg", "superutils", "hash", "taskpart", "engine"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
