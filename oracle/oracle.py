"""ctypes front for oracle/libbinstats_oracle.so — the CPU restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under vaex_b200/ imports this module.

The spec objects below (``scalar``, ``ordinal``, ``agg``) are plain dicts so that the very same specs
can be handed to the product (vaex_b200) in the parity tests.

Parity status: PINNED.  The C library is checked against the compiled reference and its known-answer vectors
(tests/test_oracle_pinning.py); the two pieces restated in Python here — ``nunique`` (AggNUniquePrimitive) and
``flat_indices`` — are pinned by the ``nunique_*`` golden vectors that tests/golden/make_golden.py produced with the
compiled reference (incl. its -0.0 / row-count quirks).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DTYPES = ["float64", "float32", "int64", "int32", "int16", "int8", "uint64", "uint32", "uint16", "uint8", "bool"]
DTYPE_CODE = {name: i for i, name in enumerate(DTYPES)}
BINNER_SCALAR, BINNER_ORDINAL = 0, 1
OPS = {"count": 0, "sum": 1, "sum_moment": 2, "min": 3, "max": 4, "first": 5, "last": 5}


def dtype_code(dtype):
    dtype = np.dtype(dtype)
    return DTYPE_CODE[dtype.newbyteorder("=").name]


def is_swapped(ar):
    return ar.dtype.byteorder not in ("=", "|") and ar.dtype.byteorder != ("<" if np.little_endian else ">")


class _Binner(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dtype", C.c_int32), ("flip", C.c_int32), ("allow_other", C.c_int32), ("invert", C.c_int32), ("pad_", C.c_int32),
                ("data", C.c_void_p), ("mask", C.c_void_p), ("vmin", C.c_double), ("vmax", C.c_double), ("bins", C.c_uint64),
                ("ordinal_count", C.c_int64), ("min_value", C.c_int64)]


class _Agg(C.Structure):
    _fields_ = [("op", C.c_int32), ("dtype", C.c_int32), ("dtype2", C.c_int32), ("flip", C.c_int32), ("invert", C.c_int32), ("use_moment", C.c_int32),
                ("moment", C.c_uint32), ("pad_", C.c_uint32), ("data", C.c_void_p), ("data2", C.c_void_p), ("mask", C.c_void_p),
                ("grid", C.c_void_p), ("grid_order", C.c_void_p), ("cell_masked", C.c_void_p)]


def build():
    """Compile the C restatement (gcc only)."""
    subprocess.check_call(["make", "-C", _HERE, "libbinstats_oracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libbinstats_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_hash64.restype = C.c_uint64
        L.orc_hash64.argtypes = [C.c_uint64]
        L.orc_hash_bits.restype = C.c_uint64
        L.orc_hash_bits.argtypes = [C.c_int, C.c_uint64]
        L.orc_bin.restype = C.c_int
        L.orc_bin.argtypes = [C.POINTER(_Binner), C.c_int, C.POINTER(_Agg), C.c_int, C.c_uint64]
        L.orc_fill_minmax.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_int]
        L.orc_fill_first.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.orc_set_create.restype = C.c_void_p
        L.orc_set_create.argtypes = [C.c_int, C.c_int, C.c_int64]
        L.orc_set_destroy.argtypes = [C.c_void_p]
        L.orc_set_from_keys.restype = C.c_void_p
        L.orc_set_from_keys.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        for name in ("count", "nan_count", "null_count", "nan_index", "null_index"):
            f = getattr(L, "orc_set_" + name)
            f.restype = C.c_int64
            f.argtypes = [C.c_void_p]
        L.orc_set_offsets.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_update.restype = C.c_int
        L.orc_set_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_key_array.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_set_map_ordinal.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_set_isin.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_set_merge.restype = C.c_int
        L.orc_set_merge.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_scalar_to_bins.restype = C.c_int
        L.orc_scalar_to_bins.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_ordinal_to_bins.restype = C.c_int
        L.orc_ordinal_to_bins.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_minmax.restype = C.c_int
        L.orc_minmax.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        _LIB = L
    return _LIB


# ------------------------------------------------------------------------------------------------
# spec constructors (shared with the product in parity tests)
# ------------------------------------------------------------------------------------------------
def scalar(data, vmin, vmax, bins, mask=None):
    """BinnerScalar spec; ``mask`` follows numpy (1 = masked)."""
    return dict(kind="scalar", data=data, mask=mask, vmin=float(vmin), vmax=float(vmax), bins=int(bins))


def ordinal(data, count, min_value=0, allow_other=False, invert=False, mask=None):
    return dict(kind="ordinal", data=data, mask=mask, count=int(count), min_value=int(min_value), allow_other=bool(allow_other), invert=bool(invert))


def agg(op, data=None, mask=None, moment=None, order=None, selection=None, dropmissing=False, dropnan=False):
    """Aggregator spec; ``mask`` follows the aggregator convention (1 = use the row).  ``nunique`` only: ``mask`` = 0 marks a null
    row, ``selection`` = 0 a row that is skipped (set_selection_mask), plus the dropmissing / dropnan constructor flags."""
    return dict(op=op, data=data, mask=mask, moment=moment, order=order, selection=selection, dropmissing=bool(dropmissing), dropnan=bool(dropnan))


def binner_shape(b):
    return b["bins"] + 3 if b["kind"] == "scalar" else b["count"] + (3 if b["allow_other"] else 2)


def upcast(dtype):
    dtype = np.dtype(dtype).newbyteorder("=")
    if dtype.kind == "f":
        return np.dtype("float64")
    if dtype.kind in "ib":
        return np.dtype("int64")
    return np.dtype("uint64")


def _ptr(ar):
    return None if ar is None else ar.ctypes.data


def _mask_u8(mask):
    if mask is None:
        return None
    mask = np.ascontiguousarray(mask)
    return mask.view(np.uint8) if mask.dtype == np.bool_ else mask.astype(np.uint8)


def binby(binners, aggs, length=None):
    """Run Grid::bin_ (agg.hpp:106-137) over one chunk.  Returns a list with, per aggregator, the full
    grid *with* edge cells, shaped (shape_0, shape_1, ...) (dim 0 = first binner, fastest in memory);
    first/last return a numpy.ma array like the reference (agg_first.cpp:61-114)."""
    L = lib()
    keep = []
    nb = len(binners)
    B = (_Binner * max(nb, 1))()
    shapes = []
    for i, b in enumerate(binners):
        data = np.asarray(b["data"])
        if data.ndim != 1 or not (data.flags.c_contiguous):
            data = np.ascontiguousarray(data)
        mask = _mask_u8(b["mask"])
        keep += [data, mask]
        B[i].kind = BINNER_SCALAR if b["kind"] == "scalar" else BINNER_ORDINAL
        B[i].dtype = dtype_code(data.dtype)
        B[i].flip = int(is_swapped(data))
        B[i].data = _ptr(data)
        B[i].mask = _ptr(mask)
        if b["kind"] == "scalar":
            B[i].vmin, B[i].vmax, B[i].bins = b["vmin"], b["vmax"], b["bins"]
        else:
            B[i].ordinal_count, B[i].min_value = b["count"], b["min_value"]
            B[i].allow_other, B[i].invert = int(b["allow_other"]), int(b["invert"])
        shapes.append(binner_shape(b))
        if length is None:
            length = len(data)
    cells = int(np.prod(shapes)) if shapes else 1
    if length is None:
        length = 0
    # nunique is restated in Python (nunique() below); everything else goes through the C driver
    nunique_at = {k: a for k, a in enumerate(aggs) if a["op"] == "nunique"}
    if nunique_at:
        rest = [a for a in aggs if a["op"] != "nunique"]
        others = iter(binby(binners, rest, length) if rest else [])
        return [nunique(binners, a["data"], a["mask"], a.get("selection"), a.get("dropmissing", False), a.get("dropnan", False), length)
                if k in nunique_at else next(others) for k, a in enumerate(aggs)]
    na = len(aggs)
    A = (_Agg * max(na, 1))()
    outs = []
    for k, a in enumerate(aggs):
        op = a["op"]
        data = None if a["data"] is None else np.ascontiguousarray(a["data"])
        order = None if a.get("order") is None else np.ascontiguousarray(a["order"])
        mask = _mask_u8(a["mask"])
        keep += [data, order, mask]
        dt = np.dtype("int64") if data is None else data.dtype.newbyteorder("=")
        A[k].op = OPS[op]
        A[k].dtype = dtype_code(dt)
        A[k].flip = int(data is not None and is_swapped(data))
        A[k].data = _ptr(data)
        A[k].mask = _ptr(mask)
        if op == "count":
            grid = np.zeros(cells, np.int64)
        elif op in ("sum", "sum_moment"):
            grid = np.zeros(cells, upcast(dt))
            if op == "sum_moment":
                A[k].use_moment, A[k].moment = 1, int(a["moment"])
        elif op in ("min", "max"):
            grid = np.zeros(cells, dt)
            L.orc_fill_minmax(A[k].dtype, grid.ctypes.data, cells, int(op == "max"))
        elif op in ("first", "last"):
            dt2 = np.dtype("int64") if order is None else order.dtype.newbyteorder("=")
            grid = np.zeros(cells, dt)
            grid_order = np.zeros(cells, dt2)
            cell_masked = np.zeros(cells, np.uint8)
            A[k].dtype2 = dtype_code(dt2)
            A[k].invert = int(op == "last")
            A[k].data2 = _ptr(order)
            L.orc_fill_first(A[k].dtype, A[k].dtype2, grid.ctypes.data, grid_order.ctypes.data, cell_masked.ctypes.data, cells, A[k].invert)
            A[k].grid_order = grid_order.ctypes.data
            A[k].cell_masked = cell_masked.ctypes.data
            keep += [grid_order, cell_masked]
            outs.append((grid, cell_masked))
            A[k].grid = grid.ctypes.data
            continue
        else:
            raise ValueError(op)
        A[k].grid = grid.ctypes.data
        outs.append(grid)
    rc = L.orc_bin(B, nb, A, na, length)
    if rc == -2:
        raise RuntimeError("data not set")
    if rc:
        raise RuntimeError(f"oracle error {rc}")
    results = []
    for o in outs:
        if isinstance(o, tuple):
            grid, cm = o
            results.append(np.ma.array(grid.reshape(shapes, order="F"), mask=cm.astype(bool).reshape(shapes, order="F")))
        else:
            results.append(o.reshape(shapes, order="F"))
    return results


def flat_indices(binners, length=None):
    """The index half of Grid::bin_ (agg.hpp:106-124): every binner's to_bins accumulated into one flat cell index per row."""
    L = lib()
    out, stride, shapes = None, 1, []
    for b in binners:
        data = np.ascontiguousarray(b["data"])
        if length is None:
            length = len(data)
        if out is None:
            out = np.zeros(length, np.uint64)
        mask = _mask_u8(b["mask"])
        if b["kind"] == "scalar":
            rc = L.orc_scalar_to_bins(dtype_code(data.dtype), int(is_swapped(data)), _ptr(data), _ptr(mask), b["vmin"], b["vmax"], b["bins"], 0, length, stride,
                                      out.ctypes.data)
        else:
            rc = L.orc_ordinal_to_bins(dtype_code(data.dtype), int(is_swapped(data)), _ptr(data), _ptr(mask), b["count"], b["min_value"], int(b["allow_other"]),
                                       int(b["invert"]), 0, length, stride, out.ctypes.data)
        if rc:
            raise RuntimeError(f"oracle error {rc}")
        shapes.append(binner_shape(b))
        stride *= shapes[-1]
    if out is None:
        out = np.zeros(length or 0, np.uint64)
    return out, shapes


def nunique(binners, data, valid=None, selection=None, dropmissing=False, dropnan=False, length=None):
    """AggNUniquePrimitive (src/agg_nunique.cpp): aggregate (:57-86) keeps one counter<T> per cell — a row outside
    ``selection`` (1 = take part) is skipped, a row with ``valid`` == 0 is a null, NaN a nan, the rest distinct keys; get_result
    (:16-42) returns keys + (any null) + (any nan), and dropmissing / dropnan subtract the cell's null / nan ROW counts
    (``null_count`` / ``nan_count`` are incremented per row, src/hash_primitives.hpp:296-301).  Restated with Python sets."""
    data = np.asarray(data)
    if length is None:
        length = len(data)
    idx, shapes = flat_indices(binners, length)
    cells = int(np.prod(shapes)) if shapes else 1
    native = data.astype(data.dtype.newbyteorder("=")) if is_swapped(data) else data
    keys = [set() for _ in range(cells)]
    nan_rows = np.zeros(cells, np.int64)
    null_rows = np.zeros(cells, np.int64)
    isnan = np.isnan(native) if native.dtype.kind == "f" else np.zeros(length, bool)
    # keys are told apart by their BITS: the hash of a double is the hash of its bit pattern (src/hash.hpp:50-152), so -0.0 and
    # 0.0 live in different buckets and are never compared — two keys (pinned by the golden vectors)
    bits = np.ascontiguousarray(native).view("u%d" % native.dtype.itemsize).tolist()
    for j in range(length):
        if selection is not None and not selection[j]:
            continue
        c = int(idx[j])
        if valid is not None and not valid[j]:
            null_rows[c] += 1
        elif isnan[j]:
            nan_rows[c] += 1
        else:
            keys[c].add(bits[j])
    out = np.array([len(k) for k in keys], np.int64) + (null_rows > 0) + (nan_rows > 0)
    if dropmissing:
        out -= null_rows
    if dropnan:
        out -= nan_rows
    return out.reshape(shapes, order="F") if shapes else out.reshape(())


def hash64(x):
    return lib().orc_hash64(int(x) & 0xFFFFFFFFFFFFFFFF)


class OrderedSet:
    """Restatement of vaex.superutils.ordered_set_<dtype> (hash_primitives.hpp:437-725)."""

    def __init__(self, dtype, nmaps=1, limit=-1, _handle=None):
        self.dtype = np.dtype(dtype)
        self.nmaps = nmaps
        self._h = lib().orc_set_create(dtype_code(self.dtype), nmaps, limit) if _handle is None else _handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_set_destroy(self._h)
            self._h = None

    @classmethod
    def from_keys(cls, keys, null_value=-1, nan_count=0, null_count=0):
        keys = np.ascontiguousarray(keys)
        h = lib().orc_set_from_keys(dtype_code(keys.dtype), keys.ctypes.data, len(keys), null_value, nan_count, null_count)
        if not h:
            raise RuntimeError("key array does not match the claimed null/nan state")
        return cls(keys.dtype, 1, _handle=h)

    def update(self, keys, masks=None, start_index=0, return_values=False):
        keys = np.ascontiguousarray(keys, dtype=self.dtype)
        masks = _mask_u8(masks)
        n = len(keys)
        values = np.zeros(n if return_values else 1, np.int64)
        map_index = np.zeros(n if return_values else 1, np.int16)
        rc = lib().orc_set_update(self._h, keys.ctypes.data, _ptr(masks), n, start_index, int(return_values), values.ctypes.data, map_index.ctypes.data)
        if rc == -3:
            raise RuntimeError("Cannot combine limit with return_inverse")
        if return_values:
            return values, map_index

    def __len__(self):
        return lib().orc_set_count(self._h)

    count = property(__len__)
    nan_count = property(lambda self: lib().orc_set_nan_count(self._h))
    null_count = property(lambda self: lib().orc_set_null_count(self._h))
    nan_index = property(lambda self: lib().orc_set_nan_index(self._h))
    null_index = property(lambda self: lib().orc_set_null_index(self._h))
    has_nan = property(lambda self: self.nan_count > 0)
    has_null = property(lambda self: self.null_count > 0)

    def offsets(self):
        out = np.zeros(self.nmaps, np.int64)
        lib().orc_set_offsets(self._h, out.ctypes.data)
        return out.tolist()

    def key_array(self):
        out = np.zeros(len(self), self.dtype)
        lib().orc_set_key_array(self._h, out.ctypes.data)
        return out

    def map_ordinal(self, keys):
        keys = np.ascontiguousarray(keys, dtype=self.dtype)
        out = np.zeros(len(keys), np.int64)
        lib().orc_set_map_ordinal(self._h, keys.ctypes.data, len(keys), out.ctypes.data)
        size = len(self)
        if size < (1 << 7):
            return out.astype(np.int8)
        if size < (1 << 15):
            return out.astype(np.int16)
        if size < (1 << 31):
            return out.astype(np.int32)
        return out

    def isin(self, keys):
        keys = np.ascontiguousarray(keys, dtype=self.dtype)
        out = np.zeros(len(keys), np.uint8)
        lib().orc_set_isin(self._h, keys.ctypes.data, len(keys), out.ctypes.data)
        return out.astype(bool)

    def flatten_values(self, values, map_index, out):
        # hash.hpp:267-288
        offsets = np.asarray(self.offsets())
        out[:] = values + offsets[map_index]
        return out

    def merge(self, others):
        for o in others:
            if lib().orc_set_merge(self._h, o._h):
                raise RuntimeError("cannot merge with an unequal maps")

    def flatten(self):
        return OrderedSet.from_keys(self.key_array(), self.null_index if self.has_null else -1, self.nan_count, self.null_count)


# ------------------------------------------------------------------------------------------------
# limits pre-pass (df.minmax): legacy TaskStatistic(OP_MIN_MAX) over vaexfast.statisticNd
# ------------------------------------------------------------------------------------------------
def minmax(data, raw=False):
    """df.minmax(expression) for one column (numpy or numpy.ma array, any of the 11 dtypes, either byte order).
    raw=True: the (min, max) doubles of the statistic grid; default: cast back to the column dtype like
    vaex/dataframe.py:1524-1528 does (`value.astype(data_type0.numpy)`)."""
    mask = None
    if np.ma.isMaskedArray(data):
        mask = np.ascontiguousarray(np.ma.getmaskarray(data)).view(np.uint8)
        data = data.data
    data = np.ascontiguousarray(data)
    code, flip = dtype_code(data.dtype), int(is_swapped(data) and data.dtype.itemsize > 1)
    out = np.empty(2, np.float64)
    rc = lib().orc_minmax(code, flip, data.ctypes.data, None if mask is None else mask.ctypes.data, len(data), out.ctypes.data)
    assert rc == 0
    if raw:
        return out
    with np.errstate(invalid="ignore"):
        return out.astype(data.dtype.newbyteorder("="))


# ------------------------------------------------------------------------------------------------
# string key sets: ordered_set<> over StringList64, restated (small cases: pure Python)
# ------------------------------------------------------------------------------------------------
_MUL = ((0xc6a4a793 << 32) + 0x5bd1e995) & 0xFFFFFFFFFFFFFFFF
_M64 = 0xFFFFFFFFFFFFFFFF


def string_hash(data: bytes) -> int:
    """std::hash<string_view> of the reference build (src/hash.hpp:59-86 -> string-view-lite -> libstdc++ _Hash_bytes: the 64-bit
    Murmur-2 variant, seed 0xc70f6907); pinned against the compiled reference in tests/test_oracle_pinning.py"""
    n = len(data)
    h = (0xc70f6907 ^ (n * _MUL)) & _M64
    body = n & ~7
    for p in range(0, body, 8):
        d = int.from_bytes(data[p:p + 8], "little")
        d = (d * _MUL) & _M64
        d ^= d >> 47
        d = (d * _MUL) & _M64
        h = ((h ^ d) * _MUL) & _M64
    if n & 7:
        h = ((h ^ int.from_bytes(data[body:], "little")) * _MUL) & _M64
    h ^= h >> 47
    h = (h * _MUL) & _M64
    h ^= h >> 47
    return h


class StringOrderedSet:
    """ordered_set<> for strings (src/hash_string.hpp:56-180 update, :437-560 ordered_set): shard = hash % nmaps, ordinal = insertion
    rank within the shard; nulls go to shard 0 at the END of the update call that first sees one; global ordinal = local + offsets."""

    def __init__(self, nmaps=1):
        self.nmaps = nmaps
        self.maps = [dict() for _ in range(nmaps)]  # str -> local ordinal (insertion ordered)
        self.null_count = 0
        self.null_value = 0x7fffffff

    def update(self, strings, start_index=0, return_values=False):
        n = len(strings)
        values, map_index = np.zeros(n, np.int64), np.zeros(n, np.int16)
        buckets = [[] for _ in range(self.nmaps)]
        nulls = []
        for i, s in enumerate(strings):
            if s is None:
                nulls.append(i)
            else:
                buckets[string_hash(s.encode("utf8")) % self.nmaps].append(i)
        for m, rows in enumerate(buckets):
            for i in rows:
                mp = self.maps[m]
                if strings[i] not in mp:
                    mp[strings[i]] = len(mp)  # add_new: map.size(); the null of shard 0 is one of its entries
                values[i], map_index[i] = mp[strings[i]], m
        for i in nulls:
            if self.null_count == 0:
                self.null_value = len(self.maps[0])
                self.maps[0][None] = self.null_value
            self.null_count += 1
            values[i], map_index[i] = self.null_value, 0
        return (values, map_index) if return_values else None

    def offsets(self):
        out, acc = [], 0
        for mp in self.maps:
            out.append(acc)
            acc += len(mp)
        return out

    def keys(self):
        return [k for mp in self.maps for k in mp]

    def __len__(self):
        return sum(len(mp) for mp in self.maps)

    @property
    def null_index(self):
        return self.null_value  # 0x7fffffff until a null was seen (src/hash_string.hpp ordered_set<>::null_index)

    def map_ordinal(self, strings):
        off = self.offsets()
        out = np.empty(len(strings), np.int64)
        for i, s in enumerate(strings):
            if s is None:
                out[i] = self.null_value if self.null_count else -1
            else:
                m = string_hash(s.encode("utf8")) % self.nmaps
                out[i] = off[m] + self.maps[m][s] if s in self.maps[m] else -1
        return out


def agg_list(cells, values, valid=None, ncells=None, dropnan=False, dropnull=False, calls=None):
    """AggListPrimitive (src/agg_list.cpp:84-113 aggregate, :47-83 get_result) restated: per cell of the flat grid the non-NaN values
    of the valid rows in arrival order, then one NaN per NaN value (unless dropnan), then one slot per null row (unless dropnull).
    `cells`: flat cell index per row; `valid`: 1 = use the row (the data mask the task part hands over, vaex/cpu.py:765-784), None =
    all; `calls`: row ranges fed as separate bin() calls.  REFERENCE QUIRK, kept: aggregate() is called per 1024-row block of a
    call with the block's offset applied to the data pointer but NOT to the mask (src/agg_list.cpp:96 `data_mask_ptr[j]` against
    :97 `data_ptr[offset + j]`), so row r of a call is judged by mask[r % 1024] of that call — the same slip as AggFirst
    (src/agg_first.cpp:131).  Returns
    (offsets int64[ncells + 1], values, nan_count[ncells], null_count[ncells]); the null slots of `values` are unspecified in the
    reference (uninitialised memory) and hold 0 here."""
    cells = np.asarray(cells, dtype=np.int64)
    values = np.asarray(values)
    n = len(cells)
    ncells = int(cells.max()) + 1 if ncells is None else int(ncells)
    lists = [[] for _ in range(ncells)]
    nan_count, null_count = np.zeros(ncells, np.int64), np.zeros(ncells, np.int64)
    isf = values.dtype.kind == "f"
    for i1, i2 in (calls or [(0, n)]):
        for j in range(i1, i2):
            c = int(cells[j])
            jm = i1 + (j - i1) % 1024  # the mask entry the reference looks at
            if valid is None or valid[jm] == 1:
                v = values[j]
                if not isf or v == v:
                    lists[c].append(v)
                elif not dropnan:
                    nan_count[c] += 1
            elif valid is not None and valid[jm] == 0 and not dropnull:
                null_count[c] += 1
    offsets = np.zeros(ncells + 1, np.int64)
    for c in range(ncells):
        offsets[c + 1] = offsets[c] + len(lists[c]) + nan_count[c] + null_count[c]
    out = np.zeros(int(offsets[-1]), values.dtype.newbyteorder("="))
    for c in range(ncells):
        k = len(lists[c])
        out[offsets[c]:offsets[c] + k] = lists[c]
        if nan_count[c]:
            out[offsets[c] + k:offsets[c] + k + nan_count[c]] = np.nan
    return offsets, out, nan_count, null_count
