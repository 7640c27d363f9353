"""``HashMapUnique`` — host-side mirror of vaex.hash.HashMapUnique over the device ordered_set.

Reference: packages/vaex-core/vaex/hash.py:62-292 (add / merge / flatten / keys / map / isin / sorted / limit and the
null/nan bookkeeping).  String and object keys are out of scope."""
import numpy as np

from . import superutils


def ordered_set_type_from_dtype(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind in "mM":
        dtype = np.dtype("int64")
    name = "ordered_set_" + dtype.newbyteorder("=").name
    if not hasattr(superutils, name):
        raise ValueError(f"Could not find a class ({name}), seems {dtype} is not supported.")
    return getattr(superutils, name)


class HashMapUnique:
    """HashMap that maps keys to unique integers (vaex/hash.py:62-64)."""

    def __init__(self, dtype, nmaps=1, limit=None, _internal=None):
        self.dtype = np.dtype(dtype)
        self.dtype_item = self.dtype
        limit = -1 if limit is None else limit
        if _internal is None:
            self._internal = ordered_set_type_from_dtype(self.dtype)(nmaps, limit)
        else:
            self._internal = _internal

    def flatten(self):
        # vaex/hash.py:75-80: rebuild a 1-shard set from key_array() so ordinals become global
        keys = self._internal.key_array()
        m = type(self._internal)(keys, self.null_index, self.nan_count, self.null_count, self.fingerprint)
        return HashMapUnique(self.dtype, _internal=m)

    @classmethod
    def from_keys(cls, keys, dtype=None, fingerprint=""):
        # vaex/hash.py:112-150
        dtype = np.dtype(keys.dtype if dtype is None else dtype)
        null_count, null_index = 0, -1
        if np.ma.isMaskedArray(keys):
            null_count = int(keys.mask.sum())
            if null_count == 1:
                null_index = int(np.where(np.ma.getmaskarray(keys))[0][0])
            elif null_count > 1:
                raise ValueError("key arrays contained more than 1 null value")
            keys = keys.data
        keys = np.asarray(keys)
        nancount = int(np.isnan(keys).sum()) if dtype.kind == "f" else 0
        internal = ordered_set_type_from_dtype(dtype)(keys, null_index, nancount, null_count, fingerprint)
        return cls(dtype, _internal=internal)

    def add(self, ar, return_inverse=False):
        # vaex/hash.py:152-171
        chunk_size = 1024 * 1024
        is_device = hasattr(ar, "__cuda_array_interface__") and not isinstance(ar, np.ndarray)
        if not is_device and np.ma.isMaskedArray(ar):
            mask = np.ma.getmaskarray(ar)
            data = np.ascontiguousarray(ar.data)
            return self._internal.update(data, mask, -1, chunk_size=chunk_size, bucket_size=chunk_size * 4, return_values=return_inverse)
        if not is_device:
            ar = np.ascontiguousarray(ar)
        return self._internal.update(ar, -1, chunk_size=chunk_size, bucket_size=chunk_size * 4, return_values=return_inverse)

    def __len__(self):
        return len(self._internal)

    def merge(self, others):
        self._internal.merge([getattr(o, "_internal", o) for o in others])

    def keys(self, mask=True):
        # vaex/hash.py:179-191
        ar = self._internal.key_array()
        if self.dtype_item.kind in "mM":
            ar = ar.view(self.dtype_item)
        if mask and self.has_null:
            m = np.zeros(ar.shape, dtype="?")
            m[self.null_index] = 1
            ar = np.ma.array(ar, mask=m)
        return ar

    def map(self, keys, check_missing=False):
        """Map key values to unique integers (vaex/hash.py:193-214)."""
        masked = isinstance(keys, np.ndarray) and np.ma.isMaskedArray(keys)
        data = np.ascontiguousarray(keys.data) if masked else keys
        indices = self._internal.map_ordinal(data)
        if masked:
            m = np.ma.getmaskarray(keys)
            if self.null_index > np.iinfo(indices.dtype).max:
                indices[m] = -1
            else:
                indices[m] = self.null_index
        if check_missing:
            indices = np.ma.array(indices, mask=indices == -1)
        return indices

    def isin(self, values):
        if isinstance(values, np.ndarray) and np.ma.isMaskedArray(values):
            isin = self._internal.isin(np.ascontiguousarray(values.data))
            isin[np.ma.getmaskarray(values)] = False
            return isin
        return self._internal.isin(values)

    has_null = property(lambda self: self._internal.has_null)
    has_nan = property(lambda self: self._internal.has_nan)
    null_index = property(lambda self: self._internal.null_index)
    nan_index = property(lambda self: self._internal.nan_index)
    nan_count = property(lambda self: self._internal.nan_count)
    null_count = property(lambda self: self._internal.null_count)

    @property
    def fingerprint(self):
        return self._internal.fingerprint

    def sorted(self, ascending=True, return_keys=False):
        # vaex/hash.py:246-268 — arrow sorts nulls last; NaN sorts after every number
        keys = self.keys(mask=False)
        has_null = self.has_null
        idx = np.arange(len(keys))
        if has_null:
            idx = idx[idx != self.null_index]
        k = keys[idx]
        order = np.argsort(k, kind="stable")
        if not ascending:
            if k.dtype.kind == "f":
                nan = np.isnan(k[order])
                order = np.concatenate([order[~nan][::-1], order[nan]])
            else:
                order = order[::-1]
        sorted_keys = k[order]
        null_index = -1
        if has_null:
            sorted_keys = np.concatenate([sorted_keys, keys[self.null_index:self.null_index + 1]])
            null_index = len(sorted_keys) - 1
        internal = type(self._internal)(sorted_keys, null_index, self._internal.nan_count, self._internal.null_count, self.fingerprint + "-sorted")
        out = HashMapUnique(self.dtype, _internal=internal)
        return (out, sorted_keys) if return_keys else out

    def limit(self, limit):
        # vaex/hash.py:270-289
        keys = self.keys(mask=False)[:limit]
        null_index, null_count = self.null_index, 1
        if null_index >= limit or not self.has_null:
            null_index, null_count = -1, 0
        nan_count = int(np.isnan(keys).sum()) if self.dtype_item.kind == "f" else 0
        internal = type(self._internal)(keys, null_index, nan_count, null_count, self.fingerprint + f"-limit-{limit}")
        return HashMapUnique(self.dtype, _internal=internal)

    def __sizeof__(self):
        return self._internal.__sizeof__()
