"""Drive the COMPILED, UNMODIFIED reference (oracle/_ref/superagg*.so, superutils*.so) from numpy.

TEST INFRASTRUCTURE ONLY.  `import vaex` is impossible offline (dask/frozendict/aplus/future missing),
so this restates the Python marshalling the reference does around its native classes:
  * TaskPartAggregation.process   — /root/reference/packages/vaex-core/vaex/cpu.py:678-786
  * _create_operation grid count  — vaex/agg.py:278-321 (grids = nthreads capped 32/16/8 by cell count)
  * chunk size                    — vaex/execution.py:283-292
  * ThreadPoolIndex stable thread — vaex/multithreading.py:64-80
  * HashMapUnique.add/map/flatten — vaex/hash.py:62-214
Used to (a) pin oracle/binstats_oracle.c against the real thing, (b) generate tests/golden fixtures,
(c) time the reference CPU path for bench.py's cpu_baseline / --impl reference legs.
"""
import importlib
import math
import os
import sys
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_mods = None


def available():
    import glob
    return bool(glob.glob(os.path.join(_REF, "superagg*.so"))) and bool(glob.glob(os.path.join(_REF, "superutils*.so")))


def modules():
    """(superagg, superutils) — superutils must load first (BinnerHash takes its hash_map types)."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("compiled reference missing: run `make -C oracle ref` where /root/reference exists")
        sys.path.insert(0, _REF)
        try:
            superutils = importlib.import_module("superutils")
            superagg = importlib.import_module("superagg")
        finally:
            sys.path.remove(_REF)
        _mods = (superagg, superutils)
    return _mods


def _suffix(ar):
    ar = np.asarray(ar)
    name = ar.dtype.newbyteorder("=").name
    swapped = ar.dtype.byteorder not in ("=", "|") and ar.dtype.byteorder != ("<" if np.little_endian else ">")
    return name + ("_non_native" if swapped else "")


def chunk_size_for(rows, nthreads, cmin=1024, cmax=1024 * 1024):
    # execution.py:283-292
    chunk = math.ceil(rows / nthreads) if nthreads else rows
    return max(cmin, min(cmax, chunk))


def grids_for(ncells, nthreads):
    # agg.py:292-303
    grids = nthreads
    if ncells >= 1e4:
        grids = min(32, nthreads)
    if ncells >= 1e5:
        grids = min(16, nthreads)
    if ncells >= 1e6:
        grids = min(8, nthreads)
    return max(grids, 1)


class RefBinby:
    """One TaskPartAggregation-like object over the reference's native classes."""

    def __init__(self, binners, aggs, nthreads=1):
        superagg, _ = modules()
        self.nthreads = nthreads
        self.binner_specs = binners
        self.agg_specs = aggs
        self.binners = []
        for b in binners:
            sfx = _suffix(b["data"])
            if b["kind"] == "scalar":
                self.binners.append(getattr(superagg, "BinnerScalar_" + sfx)(nthreads, "x", b["vmin"], b["vmax"], b["bins"]))
            else:
                self.binners.append(getattr(superagg, "BinnerOrdinal_" + sfx)(nthreads, "x", b["count"], b["min_value"], b["allow_other"], b["invert"]))
        self.grid = superagg.Grid(self.binners)
        ncells = len(self.grid)
        grids = grids_for(ncells, nthreads)
        self.aggs = []
        for a in aggs:
            op = a["op"]
            data = a["data"]
            sfx = "int64" if data is None else _suffix(data)
            if op == "count":
                agg = getattr(superagg, "AggCount_" + sfx)(self.grid, grids, nthreads)
            elif op == "sum":
                agg = getattr(superagg, "AggSum_" + sfx)(self.grid, grids, nthreads)
            elif op == "sum_moment":
                agg = getattr(superagg, "AggSumMoment_" + sfx)(self.grid, grids, nthreads, a["moment"])
            elif op == "min":
                agg = getattr(superagg, "AggMin_" + sfx)(self.grid, grids, nthreads)
            elif op == "max":
                agg = getattr(superagg, "AggMax_" + sfx)(self.grid, grids, nthreads)
            elif op in ("first", "last"):
                order = a.get("order")
                sfx2 = "int64" if order is None else np.asarray(order).dtype.newbyteorder("=").name
                name = "AggFirst_" + np.asarray(data).dtype.newbyteorder("=").name + "_" + sfx2
                if sfx.endswith("_non_native"):
                    name += "_non_native"
                agg = getattr(superagg, name)(self.grid, grids, nthreads, op == "last")
            elif op == "nunique":  # vaex/agg.py:356-369: grids = 1, one shared thread-safe set structure
                agg = getattr(superagg, "AggNUnique_" + sfx)(self.grid, 1, nthreads, a.get("dropmissing", False), a.get("dropnan", False))
            else:
                raise ValueError(op)
            self.aggs.append(agg)

    def process(self, thread, i1, i2):
        """cpu.py:678-786 for rows [i1, i2) of the spec arrays."""
        keep = []
        for binner, spec in zip(self.binners, self.binner_specs):
            block = np.ascontiguousarray(spec["data"][i1:i2])
            binner.set_data(thread, block)
            keep.append(block)
            if spec["mask"] is not None:
                m = np.ascontiguousarray(spec["mask"][i1:i2])
                binner.set_data_mask(thread, m)
                keep.append(m)
            else:
                binner.clear_data_mask(thread)
        for agg, spec in zip(self.aggs, self.agg_specs):
            if spec["data"] is not None:
                block = np.ascontiguousarray(spec["data"][i1:i2])
                agg.set_data(thread, block, 0)
                keep.append(block)
            if spec.get("order") is not None:
                block = np.ascontiguousarray(spec["order"][i1:i2])
                agg.set_data(thread, block, 1)
                keep.append(block)
            if spec["mask"] is not None:
                m = np.ascontiguousarray(spec["mask"][i1:i2])
                agg.set_data_mask(thread, m)
                keep.append(m)
            else:
                agg.clear_data_mask(thread)
            if spec["op"] == "nunique":
                if spec.get("selection") is not None:
                    m = np.ascontiguousarray(spec["selection"][i1:i2])
                    agg.set_selection_mask(thread, m)
                    keep.append(m)
                else:
                    agg.clear_selection_mask(thread)
        self.grid.bin(thread, self.aggs, i2 - i1)

    def run(self, length, chunk=None):
        """execution.py:432-435: chunks mapped over a pool whose workers keep a stable thread index."""
        chunk = chunk or chunk_size_for(length, self.nthreads)
        ranges = [(i, min(i + chunk, length)) for i in range(0, length, chunk)]
        if self.nthreads == 1:
            for i1, i2 in ranges:
                self.process(0, i1, i2)
        else:
            local = threading.local()
            lock = threading.Lock()
            counter = [0]

            def work(r):
                if not hasattr(local, "index"):
                    with lock:
                        local.index = counter[0]
                        counter[0] += 1
                self.process(local.index, *r)

            with ThreadPoolExecutor(self.nthreads) as pool:
                list(pool.map(work, ranges))
        return self.results()

    def results(self):
        return [agg.get_result() for agg in self.aggs]


def binby(binners, aggs, length=None, nthreads=1, chunk=None):
    if length is None:
        length = len(binners[0]["data"])
    return RefBinby(binners, aggs, nthreads).run(length, chunk)


def ordered_set(dtype, nmaps=1, limit=-1):
    _, superutils = modules()
    return getattr(superutils, "ordered_set_" + np.dtype(dtype).name)(nmaps, limit)


def ordered_set_from_keys(keys, null_index=-1, nan_count=0, null_count=0, fingerprint=""):
    _, superutils = modules()
    keys = np.asarray(keys)
    return getattr(superutils, "ordered_set_" + keys.dtype.name)(keys, null_index, nan_count, null_count, fingerprint)


def groupby_sum_count(keys, values, nthreads=1, nmaps=None):
    """The reference's two-pass groupby restated: pass 1 ordered_set.update per chunk
    (hash.py:152-171 with chunk_size 1M / bucket 4M), flatten; pass 2 map_ordinal -> BinnerOrdinal ->
    AggSum + AggCount (groupby.py:303-317, functions.py:2454-2463)."""
    superagg, superutils = modules()
    n = len(keys)
    nmaps = nmaps or nthreads * 7  # cpu.py:317
    oset = getattr(superutils, "ordered_set_" + keys.dtype.name)(nmaps, -1)
    chunk = chunk_size_for(n, nthreads)
    ranges = [(i, min(i + chunk, n)) for i in range(0, n, chunk)]

    def add(r):
        oset.update(keys[r[0]:r[1]], -1, 1024 * 1024, 4 * 1024 * 1024, False)

    if nthreads == 1:
        for r in ranges:
            add(r)
    else:
        with ThreadPoolExecutor(nthreads) as pool:
            list(pool.map(add, ranges))
    flat = type(oset)(oset.key_array(), oset.null_index, oset.nan_count, oset.null_count, "")
    nkeys = len(flat)
    ord_dtype = flat.map_ordinal(keys[:1]).dtype
    binner = getattr(superagg, "BinnerOrdinal_" + ord_dtype.name)(nthreads, "k", nkeys, 0, False, False)
    grid = superagg.Grid([binner])
    grids = grids_for(len(grid), nthreads)
    agg_sum = getattr(superagg, "AggSum_" + values.dtype.name)(grid, grids, nthreads)
    agg_count = getattr(superagg, "AggCount_" + values.dtype.name)(grid, grids, nthreads)
    local = threading.local()
    lock = threading.Lock()
    counter = [0]

    def work(r):
        if not hasattr(local, "index"):
            with lock:
                local.index = counter[0]
                counter[0] += 1
        t = local.index
        codes = flat.map_ordinal(keys[r[0]:r[1]])
        v = values[r[0]:r[1]]
        binner.set_data(t, codes)
        binner.clear_data_mask(t)
        agg_sum.set_data(t, v, 0)
        agg_sum.clear_data_mask(t)
        agg_count.set_data(t, v, 0)
        agg_count.clear_data_mask(t)
        grid.bin(t, [agg_sum, agg_count], r[1] - r[0])

    if nthreads == 1:
        for r in ranges:
            work(r)
    else:
        with ThreadPoolExecutor(nthreads) as pool:
            list(pool.map(work, ranges))
    return flat.key_array(), agg_sum.get_result(), agg_count.get_result()


def vaexfast():
    """the reference's legacy statistics module (src/vaexfast.cpp), compiled unmodified into oracle/_ref"""
    import glob
    if not glob.glob(os.path.join(_REF, "vaexfast*.so")):
        raise RuntimeError("compiled vaexfast missing: run `make -C oracle ref` where /root/reference exists")
    sys.path.insert(0, _REF)
    try:
        return importlib.import_module("vaexfast")
    finally:
        sys.path.remove(_REF)


def minmax(data, raw=False, chunk=1024 * 1024):
    """df.minmax(expression) restated around the compiled vaexfast.statisticNd: TaskPartStatistic.process (vaex/cpu.py:513-606:
    masked rows dropped, dtype class -> statisticNd_f8 / _f4 via as_flat_array, non-native byte order handled inside), grid
    initialised by StatOpMinMax.init, chunks of 1M rows, and the final astype to the column dtype (vaex/dataframe.py:1524-1528)."""
    vf = vaexfast()
    grid = np.zeros((2,), np.float64)
    grid[0], grid[1] = np.inf, -np.inf
    n = len(data)
    for i1 in range(0, max(n, 1), chunk):
        block = data[i1:i1 + chunk]
        if np.ma.isMaskedArray(block):
            block = np.asarray(block.data)[~np.ma.getmaskarray(block)]
        block = np.asarray(block)
        if len(block) == 0:  # the executor hands out no empty chunks; a fully masked chunk leaves nothing to reduce
            continue
        dtype = np.result_type(block.dtype)  # vaex/cpu.py:519-521: the NATIVE common dtype
        if dtype.str in ">f8 <f8 =f8":
            fn = vf.statisticNd_f8
        elif dtype.str in ">f4 <f4 =f4":
            fn = vf.statisticNd_f4
        elif dtype.str in ">i8 <i8 =i8":
            dtype, fn = np.dtype(np.float64), vf.statisticNd_f8
        else:
            dtype, fn = np.dtype(np.float32), vf.statisticNd_f4
        # as_flat_array(block, dtype) (vaex/utils.py:691-695): float64 columns with an 8-byte stride pass through whatever their
        # byte order (statisticNd then reads them through functor_double_to_native); everything else is astype'd to the native dtype
        if block.dtype.type == dtype and block.strides[0] == 8:
            flat = block
        else:
            flat = block.astype(dtype, copy=True)
        fn([], [flat], grid, [], [], 2, 0)
    if raw:
        return grid
    with np.errstate(invalid="ignore"):
        return grid.astype(np.asarray(data).dtype.newbyteorder("="))


# ---- string key sets: ordered_set<> over StringList64 (src/hash_string.hpp) through oracle/ref_strset_shim.cpp ----------------
def pack_strings(strings):
    """list of str / None -> (int64 offsets[n+1], uint8 bytes, uint8 null mask): the arrow large_string layout StringList64 uses"""
    enc = [(s.encode("utf8") if s is not None else b"") for s in strings]
    off = np.zeros(len(enc) + 1, np.int64)
    if enc:
        off[1:] = np.cumsum([len(e) for e in enc])
    by = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if off[-1] else np.zeros(0, np.uint8)
    return off, by, np.array([s is None for s in strings], np.uint8)


def unpack_strings(offsets, data, nulls=None):
    return [None if (nulls is not None and nulls[i]) else bytes(data[offsets[i]:offsets[i + 1]]).decode("utf8") for i in range(len(offsets) - 1)]


def strset_module():
    import glob
    if not glob.glob(os.path.join(_REF, "strset_shim*.so")):
        raise RuntimeError("compiled string-set shim missing: run `make -C oracle ref` where /root/reference exists")
    sys.path.insert(0, _REF)
    try:
        return importlib.import_module("strset_shim")
    finally:
        sys.path.remove(_REF)


class RefStringSet:
    """vaex.superutils.ordered_set_string(nmaps) of the compiled reference, fed with Python string lists"""

    def __init__(self, nmaps=1, limit=-1):
        self._s = strset_module().ordered_set_string(nmaps, limit)

    def update(self, strings, start_index=0, return_values=False):
        return self._s.update(*pack_strings(strings), start_index, return_values)

    def map_ordinal(self, strings):
        return np.asarray(self._s.map_ordinal(*pack_strings(strings)))

    def keys(self):
        off, by, nulls = self._s.key_array()
        return unpack_strings(np.asarray(off), np.asarray(by), np.asarray(nulls))

    def offsets(self):
        return list(self._s.offsets())

    def __len__(self):
        return len(self._s)

    null_count = property(lambda self: self._s.null_count)
    null_index = property(lambda self: self._s.null_index)
