"""Drop-in mirror of the ordinal-encoder part of ``vaex.superutils``: ``ordered_set_<dtype>`` and ``hash``.

Reference: packages/vaex-core/src/hash_primitives.cpp:3-25, 45-56 (bindings), src/hash_primitives.hpp:437-725 (ordered_set),
src/superutils.cpp:265 (``hash``).  Same constructor overloads, method names, keyword arguments, return dtypes and error
messages.  Keys may be numpy arrays (host, staged per call) or device arrays (``__cuda_array_interface__``).

Ordinals are identical to the reference run sequentially with chunks in order (one ``update`` call per chunk).  With more
than one thread the reference's own ordinals are timing dependent (insertion order under per-shard mutexes); the device
version always returns the sequential answer.
"""
import copyreg
import ctypes as C
import math

import numpy as np

from . import _lib


def hash(x):
    """``vaex.superutils.hash`` — splitmix64 finaliser (src/hash.hpp:40-45)."""
    return _lib.lib().b200_hash64(int(x) & 0xFFFFFFFFFFFFFFFF)


class _OrderedSet:
    _dtype = None
    _code = None

    def __init__(self, *args, **kwargs):
        self._ctx = _lib.context()
        self._h = C.c_void_p()
        self.fingerprint = ""
        self.sealed = False
        first = args[0] if args else kwargs.get("nmaps", 1)
        if isinstance(first, (int, np.integer)) and not isinstance(first, bool):
            nmaps = int(first)
            limit = int(args[1]) if len(args) > 1 else int(kwargs.get("limit", -1))
            self.nmaps = nmaps
            _lib.check(_lib.lib().b200_set_create(self._ctx._h, self._code, nmaps, limit, C.byref(self._h)))
        else:
            # ordered_set_<T>(keys, null_index, nan_count, null_count, fingerprint)  (src/hash_primitives.hpp:486-537)
            keys, null_index, nan_count, null_count = args[0], int(args[1]), int(args[2]), int(args[3])
            self.fingerprint = args[4] if len(args) > 4 else ""
            keys = np.ascontiguousarray(np.asarray(keys), dtype=self._np_dtype())
            if keys.ndim != 1:
                raise RuntimeError("Expected a 1d array")
            self.nmaps = 1
            _lib.check(_lib.lib().b200_set_from_keys(self._ctx._h, self._code, keys.ctypes.data if keys.size else None, len(keys), null_index, nan_count,
                                                     null_count, C.byref(self._h)))
            self.sealed = True

    @classmethod
    def _np_dtype(cls):
        return np.dtype(cls._dtype)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().b200_set_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- reference protocol --------------------------------------------------------------------------
    def update(self, values, *args, **kwargs):
        """update(values[, masks], start_index=0, chunk_size=..., bucket_size=..., return_values=False)"""
        names = ["start_index", "chunk_size", "bucket_size", "return_values"]
        masks = kwargs.pop("masks", None)
        args = list(args)
        if args and not isinstance(args[0], (int, np.integer, bool)):
            masks = args.pop(0)
        opts = dict(start_index=0, chunk_size=1024 * 128, bucket_size=1024 * 128, return_values=False)
        for n, v in zip(names, args):
            opts[n] = v
        opts.update(kwargs)
        if opts["bucket_size"] < opts["chunk_size"]:
            raise RuntimeError("bucket size should be larger than chunk_size")
        col = _lib.column(values)
        if col.dtype.itemsize != self._np_dtype().itemsize:
            raise RuntimeError("stride not equal to bytesize")
        mcol = None
        if masks is not None:
            mcol = _lib.mask_column(masks)
            if mcol.length != col.length:
                raise RuntimeError("array and mask should be of same size")
        n = col.length
        rv = bool(opts["return_values"])
        out_values = np.empty(n if rv else 1, np.int64)
        out_map = np.empty(n if rv else 1, np.int16)
        _lib.check(_lib.lib().b200_set_update(self._h, 0, col.ptr, None if mcol is None else mcol.ptr, n, int(opts["start_index"]), int(rv),
                                              out_values.ctypes.data, out_map.ctypes.data,
                                              col.memspace if mcol is None or mcol.memspace == col.memspace else _lib.MEM_MIXED, 0))
        if rv:
            return out_values, out_map
        return None

    def merge(self, others):
        if self.sealed:
            raise RuntimeError("hashmap is sealed, cannot merge")
        if not others:
            return
        arr = (C.c_void_p * len(others))(*[o._h for o in others])
        _lib.check(_lib.lib().b200_set_merge(self._h, arr, len(others)))

    def seal(self):
        self.sealed = True

    def __len__(self):
        return int(_lib.lib().b200_set_count(self._h))

    @property
    def count(self):
        return len(self)

    nan_count = property(lambda self: int(_lib.lib().b200_set_nan_count(self._h)))
    null_count = property(lambda self: int(_lib.lib().b200_set_null_count(self._h)))
    has_nan = property(lambda self: self.nan_count > 0)
    has_null = property(lambda self: self.null_count > 0)
    nan_index = property(lambda self: int(_lib.lib().b200_set_nan_index(self._h)))
    null_index = property(lambda self: int(_lib.lib().b200_set_null_index(self._h)))

    @property
    def offset(self):
        return int(self.null_count > 0) + int(self.nan_count > 0)  # src/hash.hpp:355-358

    def offsets(self):
        out = np.zeros(self.nmaps, np.int64)
        _lib.check(_lib.lib().b200_set_offsets(self._h, out.ctypes.data))
        return out.tolist()

    def key_array(self):
        out = np.empty(len(self), self._np_dtype())
        if len(out):
            _lib.check(_lib.lib().b200_set_key_array(self._h, out.ctypes.data))
        return out

    def keys(self):
        # hash_common::keys (src/hash.hpp:290-317): python list with nan / None in the special slots
        ar = self.key_array()
        out = ar.tolist()
        if self.nan_count:
            out[self.nan_index] = math.nan
        if self.null_count:
            out[self.null_index] = None
        return out

    def map_ordinal(self, values):
        col = _lib.column(values)
        if col.dtype.itemsize != self._np_dtype().itemsize:
            raise RuntimeError("stride not equal to bytesize for key values")
        odt = np.dtype(_lib.DTYPES[_lib.lib().b200_set_ordinal_dtype(self._h)])
        if col.memspace == _lib.MEM_DEVICE:
            import torch
            out = torch.empty(col.length, dtype=getattr(torch, odt.name), device=f"cuda:{self._ctx.device}")
            _lib.check(_lib.lib().b200_set_map_ordinal(self._h, 0, col.ptr, col.length, out.data_ptr(), _lib.MEM_DEVICE, 0))
            self._ctx.sync(0)
            return out
        out = np.empty(col.length, odt)
        if col.length:
            _lib.check(_lib.lib().b200_set_map_ordinal(self._h, 0, col.ptr, col.length, out.ctypes.data, _lib.MEM_HOST, 0))
        return out

    def isin(self, values):
        col = _lib.column(values)
        out = np.empty(col.length, np.uint8)
        if col.memspace == _lib.MEM_DEVICE:
            import torch
            t = torch.empty(col.length, dtype=torch.uint8, device=f"cuda:{self._ctx.device}")
            _lib.check(_lib.lib().b200_set_isin(self._h, 0, col.ptr, col.length, t.data_ptr(), _lib.MEM_DEVICE, 0))
            self._ctx.sync(0)
            return t.bool()
        if col.length:
            _lib.check(_lib.lib().b200_set_isin(self._h, 0, col.ptr, col.length, out.ctypes.data, _lib.MEM_HOST, 0))
        return out.astype(bool)

    def flatten_values(self, values, map_index, out):
        # hash_common::flatten_values (src/hash.hpp:267-288)
        values = np.asarray(values)
        map_index = np.asarray(map_index)
        if values.size != out.size:
            raise RuntimeError("output array does not match length of values")
        if values.size != map_index.size:
            raise RuntimeError("map_index array does not match length of values")
        out[:] = values + np.asarray(self.offsets(), np.int64)[map_index]
        return out

    def extract(self):
        keys = self.key_array()
        offsets = self.offsets() + [len(self)]
        maps = []
        for m in range(self.nmaps):
            d = {}
            for i in range(offsets[m], offsets[m + 1]):
                if (self.nan_count and i == self.nan_index) or (self.null_count and i == self.null_index):
                    continue
                d[keys[i].item()] = i - offsets[m]
            maps.append(d)
        return maps

    def __sizeof__(self):
        return int(_lib.lib().b200_set_bytes(self._h))


def _pickle(x):
    # vaex/hash.py:21-25
    return type(x), (x.key_array(), x.null_index, x.nan_count, x.null_count, x.fingerprint)


for _name in _lib.DTYPES:
    _cls = type("ordered_set_" + _name, (_OrderedSet,), dict(_dtype=_name, _code=_lib.DTYPE_CODE[_name]))
    _cls.__module__ = __name__
    globals()["ordered_set_" + _name] = _cls
    copyreg.pickle(_cls, _pickle)


class _Counter(_OrderedSet):
    """``vaex.superutils.counter_<dtype>`` (src/hash_primitives.hpp:344-433, bound in src/hash_primitives.cpp:36-43): key ->
    number of occurrences — what ``value_counts`` / ``unique`` run on (vaex/cpu.py:141-283).  Layout of ``keys()`` /
    ``key_array()`` / ``counts()``: NaN first, then null, then the keys (the reference lists the keys in its container's
    iteration order, which is unspecified; here they come in first-seen order)."""

    def __init__(self, nmaps=1):
        self._ctx = _lib.context()
        self._h = C.c_void_p()
        self.fingerprint = ""
        self.sealed = False
        self.nmaps = int(nmaps)
        _lib.check(_lib.lib().b200_counter_create(self._ctx._h, self._code, self.nmaps, C.byref(self._h)))

    def _order(self):
        """ordinal positions reordered to the counter layout [nan][null][keys...]"""
        n = len(self)
        special = []
        if self.nan_count:
            special.append(_OrderedSet.nan_index.fget(self))
        if self.null_count:
            special.append(_OrderedSet.null_index.fget(self))
        rest = [i for i in range(n) if i not in special]
        return np.array(special + rest, dtype=np.int64)

    nan_index = property(lambda self: 0)                                  # src/hash.hpp:289
    null_index = property(lambda self: 1 if self.nan_count else 0)        # src/hash.hpp:290

    def key_array(self):
        return _OrderedSet.key_array(self)[self._order()] if len(self) else _OrderedSet.key_array(self)

    def counts(self):
        n = len(self)
        out = np.zeros(n, np.int64)
        if n:
            _lib.check(_lib.lib().b200_set_counts(self._h, out.ctypes.data))
            out = out[self._order()]
        return out

    def keys(self):
        out = self.key_array().tolist()
        if self.nan_count:
            out[0] = math.nan
        if self.null_count:
            out[self.null_index] = None
        return out

    def map_ordinal(self, values):
        raise AttributeError("counter has no map_ordinal")


for _name in _lib.DTYPES:
    _cls = type("counter_" + _name, (_Counter,), dict(_dtype=_name, _code=_lib.DTYPE_CODE[_name]))
    _cls.__module__ = __name__
    globals()["counter_" + _name] = _cls
