#!/usr/bin/env python
"""bench.py — rows/s of the headline workload (2-D 1024^2 count over fp32 x,y; limits [-3,3]) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (default N=1)
    python bench.py --impl reference --steps K --warmup W     # the reference's own CPU superagg on the host cores

One "step" = one pass of the hot path over the whole synthetic batch (rows_per_gpu rows on every rank):
reset grid -> fused binby kernel -> (N>1) NCCL all-reduce of the 1027^2 int64 grid -> D2H of the grid.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = 1024
LIMITS = (-3.0, 3.0)
BYTES_PER_ROW = 8  # algorithmic: x and y, fp32 each
METRIC = "rows/s 2D 1024^2 count on fp32 (binby x,y; limits [-3,3])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=float, default=1e9, help="rows per GPU (weak scaling)")
    ap.add_argument("--e2e-rows", type=float, default=float(1 << 28), help="rows per GPU per step of the host-buffer (e2e) leg")
    ap.add_argument("--e2e-chunk", type=float, default=float(1 << 24))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-rows", type=float, default=0, help="rows of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time the other BASELINE.json configs (sum, 3-D mean+std, groupby)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with NVML during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.err = None
        self.nv = self.h = None
        try:  # NVML is initialised BEFORE the timed region so that the first sample lands inside it
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.err = repr(e)

    def run(self):
        nv, h = self.nv, self.h
        if nv is None:
            return
        try:
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while True:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                if self.stop_flag.wait(0.005):
                    break
        except Exception as e:
            self.err = repr(e)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "error": self.err}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(x, y, nthreads):
    """The reference's own CPU path on host arrays: compiled unmodified superagg (oracle/_ref) driven by the restated
    executor chunk loop; falls back to the C port (1 thread) only if the compiled reference is absent."""
    from oracle import oracle as O, ref_driver as R
    n = len(x)
    binners = [O.scalar(x, LIMITS[0], LIMITS[1], SHAPE), O.scalar(y, LIMITS[0], LIMITS[1], SHAPE)]
    aggs = [O.agg("count")]
    if R.available():
        t0 = time.perf_counter()
        job = R.RefBinby(binners, aggs, nthreads)
        res = job.run(n)
        dt = time.perf_counter() - t0
        return res[0], dt, "reference", nthreads
    t0 = time.perf_counter()
    res = O.binby(binners, aggs, n)
    dt = time.perf_counter() - t0
    return res[0], dt, "port", 1


def cpu_calibrated_rows(nthreads, target_s=2.5):
    import numpy as np
    rng = np.random.default_rng(1)
    n = 4_000_000
    x = rng.standard_normal(n, dtype=np.float32)
    y = rng.standard_normal(n, dtype=np.float32)
    cpu_reference_run(x, y, nthreads)
    _, dt, _, _ = cpu_reference_run(x, y, nthreads)
    rate = n / dt
    rows = int(min(4e8, max(1e7, rate * target_s)))
    return rows


def gen_host(n, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n, dtype=np.float32)
    y = rng.standard_normal(n, dtype=np.float32)
    return x, y


def run_reference(args):
    """--impl reference: the reference CPU implementation, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    rows = int(args.cpu_rows) or cpu_calibrated_rows(nthreads)
    x, y = gen_host(rows, 42)
    for _ in range(args.warmup):
        cpu_reference_run(x, y, nthreads)
    t = 0.0
    kind = cores = None
    for _ in range(args.steps):
        _, dt, kind, cores = cpu_reference_run(x, y, nthreads)
        t += dt
    value = rows * args.steps / t
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"df.count(binby=[x,y], limits=[[-3,3]]*2, shape=1024) on {rows:.3g} fp32 rows per step (bounded sample), host CPU"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": kind,
                         "sample": f"{rows} rows/step, N(0,1) fp32 x,y, seed 42, executor chunking restated (1M-row chunks)", "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU: vaex_b200 has no CPU fallback (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local)
    os.environ["VAEX_B200_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    from vaex_b200 import _lib, engine, superagg

    ctx = _lib.context(local)
    rows = int(args.rows)
    gen = torch.Generator(device="cuda").manual_seed(42 + rank)
    x = torch.empty(rows, dtype=torch.float32, device="cuda").normal_(generator=gen)
    y = torch.empty(rows, dtype=torch.float32, device="cuda").normal_(generator=gen)

    bx = superagg.BinnerScalar_float32(4, "x", LIMITS[0], LIMITS[1], SHAPE)
    by = superagg.BinnerScalar_float32(4, "y", LIMITS[0], LIMITS[1], SHAPE)
    grid = superagg.Grid([bx, by])
    agg = superagg.AggCount_int64(grid, 1, 4)  # df.count() == count('*'): dtype_in int64, no data column (vaex/agg.py:254-257)
    cells = len(grid)
    host_grid = torch.empty(cells, dtype=torch.int64).pin_memory()
    stream = engine.slot_stream(ctx, 0)
    bx.set_data(0, x)
    by.set_data(0, y)

    kernel_ms = []

    def step(timed):
        agg.reset(0)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        grid.bin(0, [agg], rows)
        e1.record(stream)
        if world > 1:
            engine.all_reduce([agg], slot=0)
        agg.read_async(0, host_grid)
        if timed:
            kernel_ms.append((e0, e1))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        step(True)
    t1.record(stream)
    barrier()
    clocks = sampler.result()
    total_ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    kms = sum(a.elapsed_time(b) for a, b in kernel_ms) / len(kernel_ms)
    counted = int(host_grid.sum().item())
    assert counted == rows * world, f"row conservation failed: grid holds {counted}, expected {rows * world}"

    value = rows * world * args.steps / (total_ms * 1e-3)
    # count(*) on a 1027^2 grid takes the tile-partition path from 2^22 rows: 2 kernels per batch of <= 2^28 rows
    launches_per_step = 2 * ((rows + (1 << 28) - 1) >> 28) if rows >= (1 << 22) else 1
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_ROW * rows / (kms * 1e-3) / 1e9

    out = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"df.count(binby=[x,y], limits=[[-3,3]]*2, shape=1024) on {rows:.3g} fp32 rows per GPU, device-resident columns",
                   "rows_per_gpu": rows, "grid_cells": cells, "parallelism": f"row-shard x{world} + NCCL all-reduce of the int64 grid",
                   "l2": "inputs (8 GB/GPU) far exceed L2; no flush needed", "index_math": "fp64, bit-exact with the reference"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of the two kernels of one batch, `ncu --set full` capture
                     # profiles/r01_ncu_final_tilecount_pair.txt: 3.391 GB per 2.6e8 rows = 13.04 B/row (8 algorithmic + 2 B/row of
                     # bucket writes + 2 B/row of bucket reads + chunk padding), scaled to the rows of one step
                     "traffic": (13.04 * rows / 1e9) if rows >= (1 << 22) else None, "traffic_unit": "GB per step",
                     "peak_source": peak_src, "kernel": "k_tile_partition<float,2,TMA> + k_tile_count (csrc/tilecount.cu)" if rows >= (1 << 22) else "k_binby_fast",
                     "kernel_ms": kms, "algorithmic_bytes_per_row": BYTES_PER_ROW, "launches_per_step": launches_per_step,
                     "note": "achieved = 8 B/row x rows per step / device time of the step's binby launches (CUDA events on the launching "
                             "stream); bound by SM instruction issue in k_tile_partition, not by HBM: see DESIGN.md section 4 and profiles/"},
        "gpu_launches": args.steps * launches_per_step,
        "clocks": clocks,
    }

    # ---- BASELINE.json configs[1] (df.sum(z, binby=[x,y], shape=1024) on the same rows), reported beside the headline -----
    if rank == 0 and world == 1 and not args.no_cpu:
        z = torch.empty(rows, dtype=torch.float32, device="cuda").normal_(generator=gen)
        asum = superagg.AggSum_float32(grid, 1, 4)
        asum.set_data(0, z, 0)

        def sum_step():
            asum.reset(0)
            grid.bin(0, [asum], rows)
        sum_step()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(3):
            sum_step()
        s1.record(stream)
        ctx.sync(0)
        torch.cuda.synchronize()
        sms = s0.elapsed_time(s1) / 3
        out["also"] = {"configs[1] df.sum(z, binby=[x,y], shape=1024), fp32, device-resident": {
            "rows_per_s": rows / (sms * 1e-3), "ms_per_step": sms, "algorithmic_bytes_per_row": 12,
            "achieved_gbs": 12 * rows / (sms * 1e-3) / 1e9, "frac": 12 * rows / (sms * 1e-3) / 1e9 / peak,
            "kernel": "k_binby_fast<float,2,float> (one RED.ADD.F64 per row: L2-request bound, DESIGN.md section 4)"}}
        del z, asum

    # ---- e2e: host buffers through the C ABI, H2D inside the timed region -------------------------------------------
    if not args.no_e2e:
        erows = int(args.e2e_rows)
        chunk = int(args.e2e_chunk)
        xh = torch.empty(erows, dtype=torch.float32).pin_memory()
        yh = torch.empty(erows, dtype=torch.float32).pin_memory()
        xh.copy_(x[:erows])
        yh.copy_(y[:erows])
        xn, yn = xh.numpy(), yh.numpy()
        nslots = 4

        def e2e_step():
            agg.reset(0)
            ctx.sync(0)
            for c, i1 in enumerate(range(0, erows, chunk)):
                i2 = min(i1 + chunk, erows)
                s = c % nslots
                bx.set_data(s, xn[i1:i2])
                by.set_data(s, yn[i1:i2])
                grid.bin(s, [agg], i2 - i1, row_offset=i1, flags=_lib.FLAG_ASYNC_HOST)
            ctx.sync(-1)
            if world > 1:
                engine.all_reduce([agg], slot=0)
            agg.read_async(0, host_grid)
            ctx.sync(0)

        e2e_step()
        barrier()
        sampler2 = ClockSampler(local)
        sampler2.start()
        w0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        w = time.perf_counter() - w0
        c2 = sampler2.result()
        if c2.get("sm_mhz") is not None:  # the e2e steps are a timed region too: fold their clock samples in
            both = sorted(sampler.samples + sampler2.samples)
            out["clocks"] = {"sm_mhz": both[len(both) // 2], "sm_max_mhz": c2["sm_max_mhz"], "reasons": sorted(set(clocks["reasons"]) | set(c2["reasons"])),
                             "samples": len(both), "windows": "device-resident steps + e2e steps"}
        if world > 1:
            t = torch.tensor([w], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            w = float(t.item())
        assert int(host_grid.sum().item()) == erows * world
        out["e2e"] = {"value": erows * world * args.e2e_steps / w, "unit": "rows/s", "h2d_bytes_per_step": erows * BYTES_PER_ROW,
                      "d2h_bytes_per_step": cells * 8, "rows_per_step_per_gpu": erows, "chunk_rows": chunk, "slots": nslots,
                      "note": "pinned host columns -> b200_bin(HOST, ASYNC) on 4 slots (H2D overlapped with kernels) -> grid D2H"}
        # restore device-resident columns on slot 0
        bx.set_data(0, x)
        by.set_data(0, y)

    # ---- CPU baseline (rank 0, N=1 only): the compiled reference on the host cores, bounded sample -------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        nthreads = os.cpu_count() or 1
        crows = int(args.cpu_rows) or cpu_calibrated_rows(nthreads, target_s=4.0)
        crows = min(crows, rows)
        xc = x[:crows].cpu().numpy()
        yc = y[:crows].cpu().numpy()
        cpu_reference_run(xc, yc, nthreads)
        best = None
        cgrid = None
        for _ in range(3):
            cgrid, dt, kind, cores = cpu_reference_run(xc, yc, nthreads)
            best = dt if best is None else min(best, dt)
        # parity on the sample while we are here: the GPU grid for the same rows must be bit-identical
        agg.reset()
        bx.set_data(0, x[:crows])
        by.set_data(0, y[:crows])
        grid.bin(0, [agg], crows)
        ggrid = agg.get_result()
        out["cpu_baseline"] = {"value": crows / best, "unit": "rows/s", "cores": cores, "kind": kind, "host_cpus": os.cpu_count(),
                               "sample": f"first {crows} rows of the GPU arrays, best of 3, executor chunk loop restated (1M-row chunks), includes get_result fold",
                               "parity_on_sample": bool(np.array_equal(np.asarray(cgrid), ggrid))}

    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
