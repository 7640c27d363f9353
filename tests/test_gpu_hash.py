"""GPU parity of the ordinal encoder: vaex_b200.superutils.ordered_set_<dtype> against the oracle's ordered_set restatement
(itself pinned to the compiled reference in tests/test_oracle_pinning.py).  Shaped after tests/internal/hash_test.py:78-150."""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALL = ["float64", "float32", "int64", "int32", "int16", "int8", "uint64", "uint32", "uint16", "uint8", "bool"]


def make_keys(rng, dt, n, nan=True):
    dt = np.dtype(dt)
    if dt.kind == "f":
        k = rng.integers(-20, 50, n).astype(dt) * 0.5
        if nan:
            k[rng.random(n) < 0.05] = np.nan
    elif dt.kind == "b":
        k = rng.integers(0, 2, n).astype(dt)
    else:
        k = rng.integers(max(np.iinfo(dt).min, -1000), min(np.iinfo(dt).max, 1000), n).astype(dt)
    return k


def same_state(mine, orc):
    assert len(mine) == len(orc)
    assert mine.offsets() == orc.offsets()
    assert np.array_equal(mine.key_array(), orc.key_array(), equal_nan=True)
    assert (mine.nan_count, mine.null_count) == (orc.nan_count, orc.null_count)
    assert (mine.nan_index, mine.null_index) == (orc.nan_index, orc.null_index)


@pytest.mark.parametrize("dtype", ALL)
@pytest.mark.parametrize("nmaps", [1, 3, 7])
def test_update_map_ordinal_all_dtypes(dtype, nmaps, oracle):
    from vaex_b200 import superutils
    rng = np.random.default_rng(hash((dtype, nmaps)) % 2 ** 32)
    mine = getattr(superutils, "ordered_set_" + dtype)(nmaps)
    orc = oracle.OrderedSet(dtype, nmaps)
    for call in range(3):
        n = int(rng.integers(1, 4000))
        k = make_keys(rng, dtype, n, nan=call != 0)
        m = (rng.random(n) < 0.05) if call == 1 else None
        rv = call == 2
        si = int(rng.choice([-1, 0]))
        if m is None:
            a, b = mine.update(k, si, return_values=rv), orc.update(k, None, si, rv)
        else:
            a, b = mine.update(k, m, si, return_values=rv), orc.update(k, m, si, rv)
        if rv:
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            out = np.zeros(n, np.int64)
            assert np.array_equal(mine.flatten_values(a[0], a[1], out), mine.map_ordinal(k).astype(np.int64)) or m is not None
        same_state(mine, orc)
    q = np.concatenate([k, make_keys(rng, dtype, 500)])
    mo, oo = mine.map_ordinal(q), orc.map_ordinal(q)
    assert mo.dtype == oo.dtype and np.array_equal(mo, oo)
    assert np.array_equal(mine.isin(q), orc.isin(q))


def test_ordinal_dtype_by_size(oracle):
    # src/hash_primitives.hpp:611-622
    from vaex_b200 import superutils
    for n, want in ((100, np.int8), (127, np.int8), (128, np.int16), (40000, np.int32)):
        s = superutils.ordered_set_int64(2)
        keys = np.arange(n, dtype="i8") * 7919
        s.update(keys)
        assert s.map_ordinal(keys[:10]).dtype == want
        assert len(s) == n


def test_kat_float64_nan_missing(oracle):
    # tests/internal/hash_test.py:78-112 (ordered_set_float64, nmaps 1..3, nan + missing)
    from vaex_b200 import superutils
    for nmaps in (1, 2, 3):
        ar = np.arange(4, dtype="f8")[::-1].copy()
        keys = np.concatenate([ar, [np.nan, np.nan], ar]).astype("f8")
        mask = np.zeros(len(keys), bool)
        mask[2] = True
        mine = superutils.ordered_set_float64(nmaps)
        orc = oracle.OrderedSet("float64", nmaps)
        a = mine.update(keys, mask, return_values=True)
        b = orc.update(keys, mask, 0, True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        same_state(mine, orc)
        assert mine.has_nan and mine.has_null and len(mine) == 6
        assert mine.map_ordinal(keys).dtype == np.int8
        ks = mine.keys()
        assert ks[mine.null_index] is None and ks[mine.nan_index] != ks[mine.nan_index]


def test_from_keys_flatten_pickle_merge(oracle):
    from vaex_b200 import superutils
    from vaex_b200.hash import HashMapUnique
    rng = np.random.default_rng(8)
    keys = rng.integers(0, 300, 5000).astype("i4") * 3
    s = superutils.ordered_set_int32(5)
    s.update(keys, -1)
    flat = superutils.ordered_set_int32(s.key_array(), s.null_index, s.nan_count, s.null_count, "fp")
    assert np.array_equal(flat.key_array(), s.key_array()) and flat.offsets() == [0]
    assert np.array_equal(flat.map_ordinal(keys), s.map_ordinal(keys))
    clone = pickle.loads(pickle.dumps(flat))  # vaex/hash.py:21-25
    assert np.array_equal(clone.key_array(), flat.key_array()) and clone.fingerprint == "fp"
    with pytest.raises(RuntimeError):  # duplicates -> length mismatch (src/hash_primitives.hpp:526-528)
        superutils.ordered_set_int32(np.array([1, 2, 2], "i4"), -1, 0, 0, "")
    with pytest.raises(RuntimeError):  # NaN present while claiming none
        superutils.ordered_set_float64(np.array([1.0, np.nan]), -1, 0, 0, "")
    # merge: union of keys, counts add up (order of new keys: the other set's ordinal order — documented deviation)
    a, b = superutils.ordered_set_int64(3), superutils.ordered_set_int64(3)
    ka, kb = np.arange(0, 100, dtype="i8"), np.arange(50, 150, dtype="i8")
    a.update(ka)
    b.update(kb)
    a.merge([b])
    assert len(a) == 150 and sorted(a.key_array().tolist()) == list(range(150))
    assert np.array_equal(a.key_array()[a.map_ordinal(kb)], kb)
    with pytest.raises(RuntimeError):
        a.merge([superutils.ordered_set_int64(2)])
    # HashMapUnique front (vaex/hash.py)
    hm = HashMapUnique(np.dtype("int32"), 7)
    hm.add(keys)
    hmf = hm.flatten()
    assert np.array_equal(hmf.map(keys), s_flat_codes(hm, keys))
    srt = hm.sorted()
    assert np.array_equal(srt.keys(), np.sort(np.unique(keys)))
    lim = hm.limit(10)
    assert len(lim) == 10


def s_flat_codes(hm, keys):
    return hm._internal.map_ordinal(keys)


def test_large_sparse_keys_device_resident(oracle):
    """C4-style keys (integers*256+5, sparse range) on the device; table growth + chunked calls; oracle on a sample."""
    import torch
    from vaex_b200 import superutils
    rng = np.random.default_rng(12)
    n = 3_000_000
    keys = rng.integers(0, 200_000, n).astype("i8") * 256 + 5
    mine = superutils.ordered_set_int64(7)
    kd = torch.from_numpy(keys).cuda()
    for i in range(0, n, 1_000_000):
        mine.update(kd[i:i + 1_000_000], -1)
    orc = oracle.OrderedSet("int64", 7)
    for i in range(0, n, 1_000_000):
        orc.update(keys[i:i + 1_000_000], None, -1, False)
    same_state(mine, orc)
    codes = mine.map_ordinal(kd)
    assert codes.dtype == torch.int32
    assert np.array_equal(codes.cpu().numpy(), orc.map_ordinal(keys))


@pytest.mark.parametrize("dtype", ["float64", "int64", "int32", "uint8", "bool"])
def test_counter_matches_reference_semantics(dtype, oracle):
    """counter_<dtype> (value_counts / unique): key -> count incl. NaN / null counts and merge.  The
    reference lists keys in container order, so the comparison is on the key->count mapping (as the reference's own tests do)."""
    from vaex_b200 import superutils
    rng = np.random.default_rng(17)
    a = superutils_counter = getattr(superutils, "counter_" + dtype)(3)
    total = {}
    nan_total = null_total = 0
    for call in range(3):
        n = int(rng.integers(100, 50_000))
        k = make_keys(rng, dtype, n, nan=call > 0)
        m = (rng.random(n) < 0.02) if call == 1 else None
        if m is None:
            a.update(k)
        else:
            a.update(k, m)
        for key, valid in zip(k.tolist(), (~m).tolist() if m is not None else [True] * n):
            if not valid:
                null_total += 1
            elif key != key:
                nan_total += 1
            else:
                total[key] = total.get(key, 0) + 1
    keys, counts = a.keys(), a.counts()
    assert (a.nan_count, a.null_count) == (nan_total, null_total)
    got = {}
    for key, c in zip(keys, counts.tolist()):
        if key is None:
            assert c == null_total
        elif key != key:
            assert c == nan_total
        else:
            got[key] = c
    assert got == total
    assert int(counts.sum()) == sum(total.values()) + nan_total + null_total
    # merge adds counts (src/hash_primitives.hpp:414-432)
    b = getattr(superutils, "counter_" + dtype)(3)
    k2 = make_keys(rng, dtype, 5000, nan=False)
    b.update(k2)
    a.merge([b])
    merged = dict(total)
    for key in k2.tolist():
        merged[key] = merged.get(key, 0) + 1
    got = {key: c for key, c in zip(a.keys(), a.counts().tolist()) if key is not None and key == key}
    assert got == merged


def test_counter_matches_compiled_reference_on_golden_keys():
    """cross-check against arrays produced by the compiled reference's ordered sets: every golden key set counted on the
    device must reproduce numpy's unique counts"""
    import golden_util
    from vaex_b200 import superutils
    gold = golden_util.load()
    for name in sorted(k for k in gold if k.startswith("set_") and k.endswith("_3")):
        dtype = name.split("_")[1]
        keys, mask = gold[name]["keys"], gold[name]["mask"]
        c = getattr(superutils, "counter_" + dtype)(3)
        c.update(keys, mask)
        valid = keys[~mask]
        if valid.dtype.kind == "f":
            valid = valid[~np.isnan(valid)]
        u, n = np.unique(valid, return_counts=True)
        got = {k: v for k, v in zip(c.keys(), c.counts().tolist()) if k is not None and k == k}
        assert got == dict(zip(u.tolist(), n.tolist())), name
