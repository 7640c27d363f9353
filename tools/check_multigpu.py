#!/usr/bin/env python
"""Row-sharded multi-GPU parity check (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_multigpu.py

Every rank bins ITS row range on its GPU through the product (vaex_b200.superagg), the grids are reduced in place with
vaex_b200.engine.all_reduce (NCCL sum / min / max; first/last through the packed (order key, global row) state), and rank 0
compares every grid with the oracle's single pass over all rows.  Prints one line per aggregator and exits non-zero on a
mismatch.  (tests/ cover the same logic on CPU with gloo; this is the real-hardware twin.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from oracle import oracle as O
    from vaex_b200 import _lib, engine, superagg
    _lib.context(local)
    n = 1_000_003
    rng = np.random.default_rng(77)  # the same full columns on every rank
    x = rng.normal(0, 1, n).astype("f4")
    y = rng.normal(0, 1, n).astype("f4")
    v = rng.normal(0, 1, n)
    u = rng.integers(0, 2 ** 32 - 1, n).astype("u4")
    t = rng.integers(-1000, 1000, n).astype("i8")
    i1, i2 = engine.shard_range(n, rank, world)
    bx = superagg.BinnerScalar_float32(1, "x", -3, 3, 64)
    by = superagg.BinnerScalar_float32(1, "y", -3, 3, 64)
    grid = superagg.Grid([bx, by])
    aggs = [superagg.AggCount_float64(grid, 1, 1), superagg.AggSum_float64(grid, 1, 1), superagg.AggMin_float64(grid, 1, 1),
            superagg.AggMax_uint32(grid, 1, 1), superagg.AggFirst_float64_int64(grid, 1, 1, False), superagg.AggFirst_float64_int64(grid, 1, 1, True),
            superagg.AggFirst_float64_int64(grid, 1, 1, False)]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a[i1:i2])).cuda()  # noqa: E731
    dx, dy, dv, dt = dev(x), dev(y), dev(v), dev(t)
    du = torch.from_numpy(np.ascontiguousarray(u[i1:i2]).view("i4")).cuda()
    bx.set_data(0, dx)
    by.set_data(0, dy)
    aggs[1].set_data(0, dv)
    aggs[2].set_data(0, dv)
    aggs[3].set_data(0, du)
    for a in aggs[4:6]:
        a.set_data(0, dv, 0)
        a.set_data(0, dt, 1)
    aggs[6].set_data(0, dv, 0)  # no order column: the chunk-local row orders the rows, the global row breaks ties
    grid.bin(0, aggs, i2 - i1, row_offset=i1)
    if world > 1:
        engine.all_reduce(aggs)
    _lib.context().sync()
    got = [a.get_result() for a in aggs]
    ok = True
    if rank == 0:
        b = [O.scalar(x, -3, 3, 64), O.scalar(y, -3, 3, 64)]
        want = O.binby(b, [O.agg("count"), O.agg("sum", v), O.agg("min", v), O.agg("max", u), O.agg("first", v, order=t), O.agg("last", v, order=t),
                           O.agg("first", v)], n)
        if world > 1:
            # no order column: the reference orders by the CHUNK-LOCAL row (src/agg_first.cpp:134), so the answer depends on the
            # chunking; here one chunk per rank, i.e. the winner of a cell is the row with the smallest (row within its shard,
            # global row) — restated with numpy
            idx, shapes = O.flat_indices(b, n)
            local_row = np.concatenate([np.arange(b_ - a_) for a_, b_ in (engine.shard_range(n, r, world) for r in range(world))])
            order = np.lexsort((np.arange(n), local_row))[::-1]
            val = np.full(int(np.prod(shapes)), 99.0)
            seen = np.zeros(len(val), bool)
            val[idx[order]] = v[order]  # the best row of every cell is written last
            seen[idx] = True
            want[6] = np.ma.array(val.reshape(shapes, order="F"), mask=~seen.reshape(shapes, order="F"))
        names = ["count", "sum", "min", "max(u32)", "first(order)", "last(order)", "first(row)"]
        for name, g, w in zip(names, got, want):
            if np.ma.isMaskedArray(w):
                same = np.array_equal(np.ma.getmaskarray(g), np.ma.getmaskarray(w)) and np.array_equal(np.asarray(g.data)[~np.ma.getmaskarray(w)],
                                                                                                    np.asarray(w.data)[~np.ma.getmaskarray(w)])
            elif name == "sum":
                same = np.allclose(g, w, rtol=1e-9, atol=1e-9)
            else:
                same = np.array_equal(g, w)
            print(f"[check_multigpu] world={world} {name}: {'OK' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
    if world > 1:
        flag = torch.tensor([int(ok)], device="cuda")
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
