#!/usr/bin/env python
"""bench.py -- Priority-Flood fill + D8 flow accumulation on a synthetic fractal DEM.

Metric (BASELINE.json): "Priority-Flood + D8 flow-accum Mcells/s on 32768^2 DEM; HBM GB/s vs peak".
One step = FillDepressions<D8> followed by FA_D8 (unit weights) over one 32768 x 32768 float32
DEM (1.07 Gcells; 4 GiB, far larger than the 126 MB L2, so no cache flush is needed between
steps).  `value` = cells / device time with the DEM resident in HBM; `e2e` = the same two calls
through the host-pointer C ABI (pinned host buffers, H2D/D2H inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 32768] [--impl reference]

N > 1 (launched by torchrun, one rank per GPU): the raster is row-sharded; fill and accumulation
exchange one-row halos over NCCL (richdem_b200/sharded.py); strong scaling (total work fixed).
`--impl reference` times the reference's own CPU implementation (oracle/_ref when it was compiled
from /root/reference, else the C port) on a bounded sample of the same workload.
"""
from __future__ import annotations

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

METRIC = "Priority-Flood fill + D8 flow accumulation throughput"
UNIT = "Mcells/s"
ND = -9999.0
SEED = 42


def BAND_FILL_KW(row0, height):
    """Experiment hook for the row-band fill (default: none): RDB_BAND_MULTIGRID=k starts every band from the lifted
    fill of the k x k max-pooled raster, RDB_BAND_VCYCLE=n adds coarse-grid corrections."""
    k = int(os.environ.get("RDB_BAND_MULTIGRID", "0"))
    if k < 2:
        return {}
    return {"multigrid": k, "row0": row0, "height": height, "vcycle": int(os.environ.get("RDB_BAND_VCYCLE", "0"))}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index: int = 0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = max((int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(sample_n: int, steps: int, warmup: int, dem=None):
    """Times the CPU implementation (reference if compiled, else port) on a sample_n^2 DEM."""
    import oracle
    O = oracle.best()
    kind = "reference" if O.kind == "reference" else "port"
    cores = os.cpu_count() if kind == "reference" else 1
    if dem is None:
        dem = oracle.fbm_terrain(sample_n, sample_n, seed=SEED)
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        f = O.fill_depressions(dem)
        t1 = time.perf_counter()
        a = O.fa_d8(f, ND)
        t2 = time.perf_counter()
        if i >= warmup:
            times.append((t2 - t, t1 - t, t2 - t1))
        del a
    tot = sum(t[0] for t in times) / len(times)
    return {
        "value": sample_n * sample_n / tot / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
        "sample": f"{sample_n}x{sample_n} fBm DEM (seed {SEED}), FillDepressions<D8> + FA_D8, mean of {len(times)} "
                  f"run(s); fill {sum(t[1] for t in times) / len(times):.2f}s + accum "
                  f"{sum(t[2] for t in times) / len(times):.2f}s; the flood and the accumulation wavefront are "
                  f"serial in the reference (OpenMP only parallelises FM_D8)",
        "seconds_per_run": tot,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(args.size, args.cpu_sample, 4096)
    cb = cpu_reference_run(n, max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_run"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.size}x{args.size} fBm DEM, FillDepressions<D8> + FA_D8 (timed on a bounded "
                               f"{n}x{n} sample per step)", "seed": SEED},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from richdem_b200 import _lib, sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.init(local_rank)
    L = _lib.lib()
    N = args.size
    W = N
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    _lib.set_stream(stream.cuda_stream)

    r0, r1, gt, gb = sharded.local_rows(N, world, rank)
    hloc = (r1 - r0) + gt + gb
    dem0 = torch.empty((hloc, W), dtype=torch.float32, device="cuda")
    _lib.check(L.rdb200_dev_generate_fbm_f32(dem0.data_ptr(), W, hloc, r0 - gt, SEED, 12, 0.0))
    work = torch.empty_like(dem0)
    acc = torch.empty((hloc, W), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    agg = {"launches": 0, "sweep_ms": 0.0, "visits": 0, "tile_cells": 4096, "rounds": 0, "iters": 0,
           "fill_ms": 0.0, "acc_ms": 0.0, "exchange_rounds": 0}

    def one_step(record: bool):
        work.copy_(dem0)
        if world == 1:
            _lib.check(L.rdb200_dev_fill_depressions_d8_f32(work.data_ptr(), W, hloc))
            s1 = _lib.stats()
            _lib.check(L.rdb200_dev_fa_d8_f32_f64(work.data_ptr(), acc.data_ptr(), W, hloc, ND, 1))
            s2 = _lib.stats()
            if record:
                agg["launches"] += s1["kernel_launches"] + s2["kernel_launches"] + 1
                agg["sweep_ms"] += s1["ms_main_kernel"]
                agg["visits"] += s1["fill_tile_visits"]
                agg["rounds"] += s1["fill_rounds"]
                agg["iters"] += s1["fill_tile_iters"]
                agg["fill_ms"] += s1["ms_total"]
                agg["acc_ms"] += s2["ms_total"]
        else:
            ta = time.perf_counter()
            filled, rounds, st = sharded.fill_band(work, gt, gb, return_stats=True, **BAND_FILL_KW(r0 - gt, N))
            torch.cuda.synchronize()
            tb = time.perf_counter()
            res, rounds2, st2 = sharded.fa_band(filled, gt, gb, ND, dinf=False, rank_rows=(r0, r1, N), return_stats=True)
            torch.cuda.synchronize()
            tc = time.perf_counter()
            if record:
                agg["fill_ms"] += (tb - ta) * 1e3
                agg["acc_ms"] += (tc - tb) * 1e3
                agg["fill_xr"] = agg.get("fill_xr", 0) + rounds
                agg["acc_xr"] = agg.get("acc_xr", 0) + rounds2
                agg["launches"] += st["kernel_launches"] + st2["kernel_launches"] + 1
                agg["sweep_ms"] += st["ms_main_kernel"]
                agg["visits"] += st["fill_tile_visits"]
                agg["rounds"] += st["fill_rounds"]
                agg["iters"] += st["fill_tile_iters"]
                agg["exchange_rounds"] += rounds + rounds2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one_step(True)
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    ms_per_step = dev_ms / args.steps
    cells = float(N) * float(N)
    value = cells / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the reference-facing calls with HOST buffers (N=1: whole raster; N>1: per-rank band) ----
    e2e = None
    if not args.no_e2e:
        try:
            own = work[gt:gt + (r1 - r0)] if world > 1 else work
            hrows = own.shape[0] if world > 1 else hloc
            h_dem = torch.empty((hrows, W), dtype=torch.float32, pin_memory=True)
            h_acc = torch.empty((hrows, W), dtype=torch.float64, pin_memory=True)
            src = dem0[gt:gt + (r1 - r0)] if world > 1 else dem0
            h_src = torch.empty((hrows, W), dtype=torch.float32, pin_memory=True)
            h_src.copy_(src)
            torch.cuda.synchronize()
            e2e_steps = max(1, min(args.steps, args.e2e_steps))
            if world == 1:
                times = []
                for i in range(1 + e2e_steps):
                    h_dem.copy_(h_src)  # host-side reset of the in/out buffer (not timed)
                    barrier()
                    ts = time.perf_counter()
                    _lib.check(L.rdb200_fill_depressions_d8_f32(h_dem.data_ptr(), W, hrows))
                    _lib.check(L.rdb200_fa_d8_f32_f64(h_dem.data_ptr(), h_acc.data_ptr(), W, hrows, ND, 1))
                    float(h_acc[hrows // 2, W // 2])  # read the result on the host
                    te = time.perf_counter()
                    if i >= 1:
                        times.append(te - ts)
                e2e_s = sum(times) / len(times)
                e2e = {"value": cells / e2e_s / 1e6, "unit": UNIT,
                       "h2d_bytes_per_step": int(2 * cells * 4), "d2h_bytes_per_step": int(cells * 4 + cells * 8),
                       "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                       "note": "rdb200_fill_depressions_d8_f32 + rdb200_fa_d8_f32_f64 on pinned host buffers"}
            else:
                times = []
                for i in range(1 + e2e_steps):
                    barrier()
                    ts = time.perf_counter()
                    loc = torch.empty((hloc, W), dtype=torch.float32, device="cuda")
                    loc[gt:gt + (r1 - r0)].copy_(h_src, non_blocking=True)
                    # ghost rows of elevation come from the neighbours
                    sharded.exchange_rows(loc, gt, gb)
                    filled, _ = sharded.fill_band(loc, gt, gb, **BAND_FILL_KW(r0 - gt, N))
                    res, _ = sharded.fa_band(filled, gt, gb, ND, dinf=False, rank_rows=(r0, r1, N))
                    h_dem.copy_(filled[gt:gt + (r1 - r0)], non_blocking=True)
                    h_acc.copy_(res[gt:gt + (r1 - r0)], non_blocking=True)
                    barrier()
                    float(h_acc[hrows // 2, W // 2])
                    te = time.perf_counter()
                    if i >= 1:
                        times.append(te - ts)
                tt = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                e2e_s = float(tt[0])
                e2e = {"value": cells / e2e_s / 1e6, "unit": UNIT,
                       "h2d_bytes_per_step": int(cells * 4), "d2h_bytes_per_step": int(cells * 4 + cells * 8),
                       "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                       "note": "per-rank pinned host band -> sharded fill + FA_D8 -> pinned host band"}
        except Exception as exc:  # never lose the headline line because the host-side leg failed
            if world > 1:
                raise
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)[:300]}

    # ---- the other stages of the path, timed once outside the headline region (N=1) ----
    other = None
    if world == 1 and not args.no_other_stages:
        try:
            other = {}
            work.copy_(dem0)
            _lib.check(L.rdb200_dev_fill_depressions_d8_f32(work.data_ptr(), W, hloc))
            _lib.check(L.rdb200_dev_resolve_flats_epsilon_f32(work.data_ptr(), W, hloc, ND))
            s = _lib.stats()
            other["resolve_flats_ms"] = s["ms_total"]
            other["resolve_flats_bfs_levels"] = s["flat_bfs_levels"]
            other["resolve_flats_cells_raised"] = s["flat_cells_raised"]
            _lib.check(L.rdb200_dev_fa_d8_f32_f64(work.data_ptr(), acc.data_ptr(), W, hloc, ND, 1))
            other["fa_d8_after_flats_ms"] = _lib.stats()["ms_total"]
            _lib.check(L.rdb200_dev_fa_tarboton_f32_f64(work.data_ptr(), acc.data_ptr(), W, hloc, ND, 1))
            s = _lib.stats()
            other["fa_dinf_after_flats_ms"] = s["ms_total"]
            other["fa_dinf_frontier_rounds"] = s["accum_rounds"]
            dirs = torch.empty((hloc, W), dtype=torch.uint8, device="cuda")
            _lib.check(L.rdb200_dev_d8_flow_directions_f32(work.data_ptr(), dirs.data_ptr(), W, hloc, ND))
            other["d8_flow_directions_ms"] = _lib.stats()["ms_total"]
            del dirs
        except Exception as exc:
            other = {"error": repr(exc)[:300]}

    if world > 1:
        tot = torch.tensor([agg["launches"], agg["visits"], agg["sweep_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        launches_all, visits_all = int(tot[0]), int(tot[1])
        sweep_ms_mean = float(tot[2]) / world
    else:
        launches_all, visits_all, sweep_ms_mean = agg["launches"], agg["visits"], agg["sweep_ms"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    # roofline of the dominant kernel (fill_sweep_kernel): algorithmic bytes = 12 B per cell swept
    # (read W, read Z, write W -- SURVEY 8d) x cells of the tiles it visited, / its device time
    sweep_bytes = 12.0 * agg["visits"] * agg["tile_cells"]
    achieved = sweep_bytes / (agg["sweep_ms"] * 1e-3) / 1e9 if agg["sweep_ms"] > 0 else 0.0
    launches_per_step_kernel = agg["rounds"] / args.steps if args.steps else 0
    roofline = {
        "bound": "hbm", "kernel": "fill_sweep_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "peak_source": peak_src, "traffic": args.traffic,
        "algorithmic_bytes_per_launch": sweep_bytes / max(agg["rounds"], 1),
        "avg_launch_ms": agg["sweep_ms"] / max(agg["rounds"], 1),
        "sweep_share_of_step": agg["sweep_ms"] / dev_ms,
        "tile_visits_per_step": agg["visits"] / args.steps, "full_raster_equivalents_per_step":
            agg["visits"] * agg["tile_cells"] / args.steps / (cells / world),
        "in_tile_passes_per_visit": agg["iters"] / max(agg["visits"], 1),
        "sweep_launches_per_step": launches_per_step_kernel,
        "end_to_end_fill_fraction_8B_per_cell": (8.0 * cells / world) / (agg["fill_ms"] / args.steps * 1e-3) / 1e9 / peak
        if agg["fill_ms"] > 0 else None,
    }
    cb = None
    if world == 1 and not args.no_cpu_baseline:
        sn = min(N, args.cpu_sample)
        sample = torch.empty((sn, sn), dtype=torch.float32, device="cuda")
        _lib.check(L.rdb200_dev_generate_fbm_f32(sample.data_ptr(), sn, sn, 0, SEED, 12, 0.0))
        cb = cpu_reference_run(sn, 1, 0, dem=sample.cpu().numpy())
        del sample
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{N}x{N} synthetic fBm float32 DEM (seed {SEED}, 12 octaves), FillDepressions<D8> "
                               f"+ FA_D8 unit weights", "l2": "inputs (4 B/cell DEM + 8 B/cell accumulation) exceed "
                               "the 126 MB L2; no flush needed", "sharding": f"{world} row band(s)",
                   "wall_ms_per_step": wall_ms / args.steps},
        "stages_ms_per_step": {"fill": agg["fill_ms"] / args.steps, "fa_d8": agg["acc_ms"] / args.steps}
        if world == 1 else {"fill_rank0_wall": agg["fill_ms"] / args.steps, "fa_d8_rank0_wall": agg["acc_ms"] / args.steps,
                            "fill_exchange_rounds": agg.get("fill_xr", 0) / args.steps,
                            "fa_exchange_rounds": agg.get("acc_xr", 0) / args.steps},
        "other_stages_once": other,
        "roofline": roofline,
        "cpu_baseline": cb,
        "e2e": e2e,
        "gpu_launches": launches_all,
        "clocks": clk.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=8192, help="edge of the bounded CPU sample DEM")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-stages", action="store_true")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per sweep launch from the committed ncu capture (default: profiles/traffic.json)")
    args = ap.parse_args()
    if args.traffic is None:
        try:
            args.traffic = float(json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["dram_bytes_per_launch"])
        except Exception:
            args.traffic = None
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
