"""The N>1 path on CPU: world_size-2 `gloo` run of the row-sharding + grid all-reduce logic (vaex_b200.engine), with the
oracle standing in for the per-rank kernel.  Checks that shard ranges tile the rows and that sum / min / max grids reduce
to exactly what one pass over all rows gives."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from vaex_b200 import _lib, engine
    rng = np.random.default_rng(123)  # every rank generates the same full columns, then takes its shard
    x = rng.normal(0, 1, n).astype("f4")
    y = rng.normal(0, 1, n).astype("f4")
    v = rng.normal(0, 1, n)
    u = rng.integers(0, 2 ** 32 - 1, n).astype("u4")
    i1, i2 = engine.shard_range(n, rank, world)
    b = [O.scalar(x[i1:i2], -3, 3, 32), O.scalar(y[i1:i2], -3, 3, 32)]
    aggs = [O.agg("count"), O.agg("sum", v[i1:i2]), O.agg("min", v[i1:i2]), O.agg("max", v[i1:i2]), O.agg("max", u[i1:i2]), O.agg("min", u[i1:i2])]
    ops = [_lib.AGG_COUNT, _lib.AGG_SUM, _lib.AGG_MIN, _lib.AGG_MAX, _lib.AGG_MAX, _lib.AGG_MIN]
    grids = O.binby(b, aggs, i2 - i1)
    reduced = []
    for g, op in zip(grids, ops):
        flat = np.ascontiguousarray(g.reshape(-1, order="F"))
        unsigned = flat.dtype.kind == "u"
        t = torch.from_numpy(flat.view("i4") if unsigned else flat)
        engine.all_reduce_tensor(t, op, unsigned_as_signed=unsigned)
        reduced.append(t.numpy().view(flat.dtype) if unsigned else t.numpy())
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), *reduced)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_tile_rows():
    sys.path.insert(0, ROOT)
    from vaex_b200 import engine
    for n in (0, 1, 7, 1000, 10 ** 10 + 3):
        for world in (1, 2, 3, 8):
            r = [engine.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_allreduce_matches_single_pass(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    n, world = 40_001, 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "reduced.npz")
    reduced = [z[k] for k in z.files]
    rng = np.random.default_rng(123)
    x = rng.normal(0, 1, n).astype("f4")
    y = rng.normal(0, 1, n).astype("f4")
    v = rng.normal(0, 1, n)
    u = rng.integers(0, 2 ** 32 - 1, n).astype("u4")
    b = [O.scalar(x, -3, 3, 32), O.scalar(y, -3, 3, 32)]
    aggs = [O.agg("count"), O.agg("sum", v), O.agg("min", v), O.agg("max", v), O.agg("max", u), O.agg("min", u)]
    want = [g.reshape(-1, order="F") for g in O.binby(b, aggs, n)]
    assert np.array_equal(reduced[0], want[0])                      # counts: bit-exact
    assert np.allclose(reduced[1], want[1], rtol=1e-12, atol=1e-12)  # fp64 sums: order of addition differs
    for k in (2, 3, 4, 5):
        assert np.array_equal(reduced[k], want[k]), k                # min/max incl. unsigned through the signed view


def _worker_sets(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from vaex_b200 import engine
    rng = np.random.default_rng(99)
    keys = rng.integers(0, 500, n).astype("f8") * 0.5
    keys[rng.random(n) < 0.01] = np.nan
    i1, i2 = engine.shard_range(n, rank, world)
    local = O.OrderedSet("float64", 3)
    local.update(keys[i1:i2], None, -1, False)
    union = engine.union_key_sets(local, make_set=lambda: O.OrderedSet("float64", 3))
    np.save(os.path.join(out_dir, f"keys_{rank}.npy"), union.key_array())
    codes = union.map_ordinal(keys[i1:i2]).astype(np.int64)
    counts = torch.from_numpy(np.bincount(codes, minlength=len(union)).astype(np.int64))
    dist.all_reduce(counts)
    if rank == 0:
        np.save(os.path.join(out_dir, "counts.npy"), counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_groupby_key_union_gives_identical_ordinals(tmp_path):
    """pass 1 of a row-sharded groupby: both ranks end up with the same key order, and the all-reduced per-ordinal counts
    equal a single-process groupby over all rows"""
    n, world = 30_000, 2
    mp.spawn(_worker_sets, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    k0, k1 = np.load(tmp_path / "keys_0.npy"), np.load(tmp_path / "keys_1.npy")
    assert np.array_equal(k0, k1, equal_nan=True)
    counts = np.load(tmp_path / "counts.npy")
    rng = np.random.default_rng(99)
    keys = rng.integers(0, 500, n).astype("f8") * 0.5
    keys[rng.random(n) < 0.01] = np.nan
    assert counts.sum() == n
    for key, c in zip(k0, counts):
        want = np.isnan(keys).sum() if key != key else (keys == key).sum()
        assert c == want


def _local_first_state(idx, cells, order, row0, v, last):
    """per-cell winner of one shard, the way csrc/first.cu keeps it: key = order as u64 bits with the sign flipped (LAST: its
    complement), row = global row index; lexicographic minimum wins"""
    key = (order.astype(np.int64).view(np.uint64) ^ np.uint64(1 << 63))
    if last:
        key = ~key
    rows = (row0 + np.arange(len(idx))).astype(np.uint64)
    best_key = np.full(cells, np.iinfo(np.uint64).max, np.uint64)
    best_row = np.full(cells, np.iinfo(np.uint64).max, np.uint64)
    value = np.full(cells, 99.0)
    ordv = np.zeros(cells, np.int64)
    masked = np.ones(cells, np.int8)
    for j in np.lexsort((rows, key))[::-1]:  # worst first, so the best row of a cell is written last
        c = int(idx[j])
        best_key[c], best_row[c], value[c], ordv[c], masked[c] = key[j], rows[j], v[j], order[j], 0
    return best_key, best_row, value, ordv, masked


def _worker_first(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from vaex_b200 import engine
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 10, n)
    v = rng.normal(0, 1, n)
    t = rng.integers(-20, 20, n).astype("i8")  # many ties inside and across shards: the global row breaks them
    i1, i2 = engine.shard_range(n, rank, world)
    idx, shapes = O.flat_indices([O.scalar(x[i1:i2], 0, 10, 40)], i2 - i1)
    for last in (False, True):
        key, row, value, ordv, masked = _local_first_state(idx, shapes[0], t[i1:i2], i1, v[i1:i2], last)
        tk, tr = torch.from_numpy(key.view(np.int64).copy()), torch.from_numpy(row.view(np.int64).copy())
        tv, to, tm = torch.from_numpy(value.view(np.int64).copy()), torch.from_numpy(ordv.copy()), torch.from_numpy(masked.copy())
        engine.all_reduce_first_tensors(tk, tr, tv, to, tm)
        if rank == 0:
            np.savez(os.path.join(out_dir, f"first_{int(last)}.npz"), value=tv.numpy().view(np.float64), order=to.numpy(), masked=tm.numpy(), row=tr.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_three_rank_first_last_reduction_matches_single_pass(tmp_path):
    """first / last across row shards: MIN all-reduces on the packed (order key, global row) state + an owner-only
    SUM of the value bits give exactly the single-pass result of the oracle, ties across shards included"""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    n, world = 5_000, 3
    mp.spawn(_worker_first, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 10, n)
    v = rng.normal(0, 1, n)
    t = rng.integers(-20, 20, n).astype("i8")
    want_first, want_last = O.binby([O.scalar(x, 0, 10, 40)], [O.agg("first", v, order=t), O.agg("last", v, order=t)], n)
    for last, want in ((0, want_first), (1, want_last)):
        z = np.load(tmp_path / f"first_{last}.npz")
        assert np.array_equal(z["masked"].astype(bool), np.ma.getmaskarray(want))
        ok = ~np.ma.getmaskarray(want)
        assert np.array_equal(z["value"][ok], np.asarray(want.data)[ok])
