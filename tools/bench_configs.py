#!/usr/bin/env python
"""Times every BASELINE.json configuration on one B200 with device-resident columns (the roofline setting) and prints one
JSON object per config: rows/s, achieved algorithmic GB/s (SURVEY.md 8d bytes/row) and the fraction of the measured HBM peak.

    python tools/bench_configs.py [--rows 1e9] [--configs C1,C2,C3,C4,C5]

This is a companion to bench.py (which owns the headline line the driver parses); results are quoted in DESIGN.md.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=float, default=1e9)
    ap.add_argument("--configs", default="C1,C5,C2,C3,C4")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    from vaex_b200 import _lib, engine, superagg, superutils
    from vaex_b200.frame import Frame

    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    ctx = _lib.context(0)
    stream = engine.slot_stream(ctx, 0)
    gen = torch.Generator(device="cuda").manual_seed(42)

    def timed(fn, reps):
        fn()
        ctx.sync()
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best

    def report(name, workload, rows, bytes_per_row, ms, **extra):
        gbs = rows * bytes_per_row / (ms * 1e-3) / 1e9
        print(json.dumps(dict(config=name, workload=workload, rows=rows, ms=ms, rows_per_s=rows / (ms * 1e-3), algorithmic_bytes_per_row=bytes_per_row,
                              achieved_gbs=gbs, frac_of_measured_hbm=gbs / peak, **extra)), flush=True)

    n = int(args.rows)
    for cfg in args.configs.split(","):
        torch.cuda.empty_cache()
        if cfg == "C1":
            rows = 10_000_000
            x = torch.empty(rows, dtype=torch.float64, device="cuda").normal_(generator=gen)
            b = superagg.BinnerScalar_float64(1, "x", -3.0, 3.0, 128)
            g = superagg.Grid([b])
            a = superagg.AggCount_int64(g, 1, 1)
            b.set_data(0, x)

            def run():
                a.reset(0)
                g.bin(0, [a], rows)
            ms = timed(run, args.reps)
            assert int(a.get_result().sum()) == rows
            report("C1", "df.count(binby=x, shape=128) on 1e7 fp64 rows (shared-memory privatised)", rows, 8, ms)
        elif cfg in ("C5", "C2"):
            x = torch.empty(n, dtype=torch.float32, device="cuda").normal_(generator=gen)
            y = torch.empty(n, dtype=torch.float32, device="cuda").normal_(generator=gen)
            bx = superagg.BinnerScalar_float32(1, "x", -3.0, 3.0, 1024)
            by = superagg.BinnerScalar_float32(1, "y", -3.0, 3.0, 1024)
            g = superagg.Grid([bx, by])
            bx.set_data(0, x)
            by.set_data(0, y)
            if cfg == "C5":
                a = superagg.AggCount_int64(g, 1, 1)

                def run():
                    a.reset(0)
                    g.bin(0, [a], n)
                ms = timed(run, args.reps)
                assert int(a.get_result().sum()) == n
                report("C5/headline", "df.count(binby=[x,y], shape=1024) on fp32 rows, one GPU's shard", n, 8, ms)
            else:
                z = torch.empty(n, dtype=torch.float32, device="cuda").normal_(generator=gen)
                a = superagg.AggSum_float32(g, 1, 1)
                a.set_data(0, z, 0)

                def run():
                    a.reset(0)
                    g.bin(0, [a], n)
                ms = timed(run, args.reps)
                total = float(a.get_result().sum())
                ref = float(z.double().sum())
                assert abs(total - ref) <= 1e-6 * max(1.0, abs(ref)) + 1e-6 * n ** 0.5, (total, ref)
                report("C2", "df.sum(z, binby=[x,y], shape=1024) on fp32 rows", n, 12, ms)
                del z
            del x, y
        elif cfg == "C3":
            cols = [torch.empty(n, dtype=torch.float64, device="cuda").normal_(generator=gen) for _ in range(4)]
            bs = [superagg.BinnerScalar_float64(1, "xyz"[i], -3.0, 3.0, 256) for i in range(3)]
            g = superagg.Grid(bs)
            for b, c in zip(bs, cols):
                b.set_data(0, c)
            aggs = [superagg.AggCount_float64(g, 1, 1), superagg.AggSum_float64(g, 1, 1), superagg.AggSumMoment_float64(g, 1, 1, 2)]
            for a in aggs:
                a.set_data(0, cols[3], 0)

            def run():
                for a in aggs:
                    a.reset(0)
                g.bin(0, aggs, n)
            ms = timed(run, max(2, args.reps // 2))
            assert int(aggs[0].get_result().sum()) == n
            report("C3", "df.mean(v)+df.std(v) (count, sum, sum^2 fused) binby=[x,y,z], shape=256 on fp64 rows", n, 32, ms, grid_cells=len(g))
            del cols
        elif cfg == "C4":
            keys = torch.randint(0, 1_000_000, (n,), device="cuda", dtype=torch.int64, generator=gen) * 256 + 5
            v = torch.empty(n, dtype=torch.float64, device="cuda").normal_(generator=gen)
            torch.cuda.synchronize()  # data generation is asynchronous: keep it out of the timed region
            t0 = time.perf_counter()
            s = superutils.ordered_set_int64(7)
            s.update(keys, -1)
            nkeys = len(s)
            ctx.sync()
            t_pass1 = time.perf_counter() - t0
            # timed again on a warm (already grown) table
            t0 = time.perf_counter()
            s2 = superutils.ordered_set_int64(7)
            s2.update(keys, -1)
            assert len(s2) == nkeys
            t_pass1b = time.perf_counter() - t0
            hb = superagg.BinnerHash_int64(1, "k", s)
            g = superagg.Grid([hb])
            hb.set_data(0, keys)
            asum = superagg.AggSum_float64(g, 1, 1)
            acnt = superagg.AggCount_float64(g, 1, 1)
            for a in (asum, acnt):
                a.set_data(0, v, 0)

            def run():
                asum.reset(0)
                acnt.reset(0)
                g.bin(0, [asum, acnt], n)
            ms2 = timed(run, max(2, args.reps // 2))
            assert int(acnt.get_result().sum()) == n
            report("C4/pass1", "ordered_set_int64.update over 1e6 sparse keys (first build incl. table growth)", n, 8, t_pass1 * 1e3, unique_keys=nkeys,
                   second_build_ms=t_pass1b * 1e3)
            report("C4/pass2", "groupby sum+count through the fused hash binner (probe + 2 REDs per row)", n, 16, ms2, unique_keys=nkeys)
            report("C4/total", "df.groupby(k).agg({v:[sum,count]}) both passes", n, 24, t_pass1b * 1e3 + ms2, unique_keys=nkeys)
            del keys, v
        elif cfg == "NU":
            # per-cell nunique: 64 cells, values drawn from 1e5 / 1e7 distinct ints -> 5e6 / ~1e8-pair tables
            for card in (100_000, 10_000_000):
                gcol = torch.randint(0, 64, (n,), device="cuda", dtype=torch.int32, generator=gen)
                v = torch.randint(0, card, (n,), device="cuda", dtype=torch.int64, generator=gen)
                torch.cuda.synchronize()
                bo = superagg.BinnerOrdinal_int32(1, "g", 64, 0, False, False)
                g = superagg.Grid([bo])
                bo.set_data(0, gcol)
                a = superagg.AggNUnique_int64(g, 1, 1, False, False)
                a.set_data(0, v, 0)
                t0 = time.perf_counter()
                g.bin(0, [a], n)
                ctx.sync()
                t_first = time.perf_counter() - t0
                pairs = int(a.get_result().sum())
                t0 = time.perf_counter()
                g.bin(0, [a], n)  # same rows again: every pair is found, nothing is inserted, the table does not grow
                ctx.sync()
                t_again = time.perf_counter() - t0
                assert int(a.get_result().sum()) == pairs
                report(f"NU/{card}", "nunique(v) binby 64 ordinal cells, int64 values (first pass incl. table growth; second pass = lookups only)", n, 12,
                       t_first * 1e3, distinct_pairs=pairs, lookups_only_ms=t_again * 1e3)
                del gcol, v, a
        elif cfg == "SG":
            # sparse two-key groupby: 3000 x 3000 possible combinations, sum + count of v
            from vaex_b200.frame import Frame
            k1 = torch.randint(0, 3000, (n,), device="cuda", dtype=torch.int64, generator=gen) * 1000 + 7
            k2 = torch.randint(0, 3000, (n,), device="cuda", dtype=torch.int64, generator=gen)
            v = torch.empty(n, dtype=torch.float64, device="cuda").normal_(generator=gen)
            torch.cuda.synchronize()
            df = Frame(dict(k1=k1, k2=k2, v=v))
            t0 = time.perf_counter()
            gb = df.groupby(["k1", "k2"], combine=True)
            ctx.sync()
            t_keys = time.perf_counter() - t0
            t0 = time.perf_counter()
            out = gb.agg({"v": ["sum", "count"]})
            t_agg = time.perf_counter() - t0
            assert int(out["count"].sum()) == n
            report("SG", "df.groupby([k1,k2], combine=True).agg({v:[sum,count]}): 2 key sets + combined-code set, then fused probe pass", n, 24 + 24,
                   (t_keys + t_agg) * 1e3, groups=len(out["count"]), key_sets_ms=t_keys * 1e3, aggregate_ms=t_agg * 1e3)
            del k1, k2, v


if __name__ == "__main__":
    main()
