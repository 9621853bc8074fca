#!/usr/bin/env python
"""bench.py - N-1 power flows solved per second on a ~10k-bus grid (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic scenarios:
the full branch N-1 of the 100x100 tiled grid (src/synthetic_grids.jl topology,
10 000 buses / 29 601 branches, with the documented PV/load augmentation) per GPU.
With N > 1 ranks the batch list grows with seeded random N-2 pairs (BASELINE.json
config 5) so every rank keeps the same amount of work (weak scaling); rank r solves
slice r and the 48-byte per-scenario result records are gathered with one NCCL
all-gather.

  value  : scenarios / s, device time of sblx_run_staged (inputs resident in HBM),
           CUDA events on the engine's stream, max over ranks
  e2e    : scenarios / s through sblx_solve_outages with HOST buffers (host topology
           analysis + H2D + solve + D2H of the result arrays inside the timed region)
  roofline: dominant kernel k_front (numeric multifrontal LU), see DESIGN.md
  --impl reference : the reference's CPU path (oracle port, all host cores) on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "sparlectra.jl_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "n1_power_flows_per_second"
UNIT = "scenarios/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100)
    ap.add_argument("--cols", type=int, default=100)
    ap.add_argument("--scenarios", type=int, default=0, help="scenarios per rank per step (0 = full N-1 of the grid)")
    ap.add_argument("--max-slots", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="scenarios in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ordering", type=int, default=0)
    ap.add_argument("--relax-small", type=int, default=0)
    ap.add_argument("--relax-frac", type=float, default=0.0)
    ap.add_argument("--panel-kb", type=float, default=0.0)
    ap.add_argument("--factor-mode", type=int, default=-1)
    return ap.parse_args()


def workload(args, world):
    from sblx import grid as G
    g = G.tiled_grid(rows=args.rows, cols=args.cols, augment=True, rating_mva=100.0)
    n1 = [[k] for k in range(g.nbranch)]
    per_rank = args.scenarios if args.scenarios > 0 else len(n1)
    total = per_rank * world
    cases = list(n1[:total])
    if len(cases) < total:
        rng = np.random.Generator(np.random.PCG64(20260724))
        while len(cases) < total:
            a, b = rng.integers(0, g.nbranch, size=2)
            if a != b:
                cases.append(sorted((int(a), int(b))))
    name = f"tiled_{args.rows}x{args.cols}_aug_N-1" + ("+randomN-2" if total > len(n1) else "")
    return g, cases, per_rank, name


class ClockSampler:
    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.th = None

    def start(self):
        def run():
            q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            while not self._stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self.th = threading.Thread(target=run, daemon=True)
        self.th.start()

    def stop(self):
        self._stop.set()
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_sample_size(g, ncores):
    """Bounded CPU sample: about 15 s of work (the oracle port needs ~0.12 ms per bus per scenario on one core)."""
    per_scen_s = max(1e-3, 1.2e-4 * g.n)
    return int(min(max(ncores, 8, ncores * round(15.0 / per_scen_s)), 4096))


def cpu_reference_rate(g, vm, va, cases, nsample, nproc):
    """The reference's CPU path (oracle port) on a bounded sample, scenarios in contiguous chunks
    over `nproc` worker processes (contingency.jl:304-312 fans chunks out over threads)."""
    import multiprocessing as mp
    sample = cases[:nsample]
    chunks = [sample[i::nproc] for i in range(nproc)]
    chunks = [c for c in chunks if c]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks), initializer=_cpu_init, initargs=(g, vm, va)) as pool:
        pool.map(_cpu_chunk, [[] for _ in chunks])      # workers forked and imports done before the clock starts
        t0 = time.perf_counter()
        res = pool.map(_cpu_chunk, chunks, chunksize=1)
        dt = time.perf_counter() - t0
    its = sum(r[1] for r in res)
    return len(sample) / dt, dt, its


_CPU = {}


def _cpu_init(g, vm, va):
    os.environ["OMP_NUM_THREADS"] = "1"
    _CPU.update(g=g, vm=vm, va=va)


def _cpu_chunk(chunk):
    import sparlectra_oracle as O
    it = 0
    for c in chunk:
        r = O.runpf(_CPU["g"], _CPU["vm"], _CPU["va"], c)
        if r.converged:
            Vr = np.abs(r.V) * np.exp(1j * np.radians(np.degrees(np.angle(r.V))))
            O.contingency_metrics(_CPU["g"], Vr, r.iso, c)
        it += r.iters
    return len(chunk), it


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    from sblx import api, model as M, shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g, cases, per_rank, wname = workload(args, world)
    lo, hi = shard.shard_range(len(cases), rank, world)
    mine = cases[lo:hi]
    a0 = M.build_arrays(g)
    t0 = time.perf_counter()
    extra = {}
    if args.relax_small > 0:
        extra["relax_small"] = args.relax_small
    if args.relax_frac > 0:
        extra["relax_zero_frac"] = args.relax_frac
    if args.panel_kb > 0:
        extra["panel_smem_kb"] = args.panel_kb
    if args.factor_mode >= 0:
        extra["factor_mode"] = args.factor_mode
    eng = api.Engine(a0, device=local_rank, max_slots=args.max_slots, ordering=args.ordering, **extra)
    t_create = time.perf_counter() - t0
    Vb, conv, itb, resb, stb = eng.solve_one(None)
    if not conv:
        raise SystemExit("bench.py: base case did not converge")
    vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng.set_v0(a.v0)
    info = eng.info()
    optr, obr = eng.pack_outages(mine)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident arm (value)
    eng.stage_outages(optr, obr)
    info = eng.info()
    for _ in range(args.warmup):
        eng.run_staged()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, wall0 = 0.0, time.perf_counter()
    agg = dict(factor_ms=0.0, solve_ms=0.0, mismatch_ms=0.0, jacobian_ms=0.0, other_ms=0.0, launches=0, factor_launches=0,
               scenario_iterations=0, mismatch_evaluations=0, ticks=0)
    for _ in range(args.steps):
        eng.run_staged()
        st = eng.stats()
        dev_ms += st["total_ms"]
        for k in agg:
            agg[k] += st[k]
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    res = eng.fetch()
    nconv = int(res["converged"].sum())
    # one NCCL all-gather of the packed per-scenario records
    gathered = None
    if world > 1:
        ptr, nbytes = eng.device_summary()

        class _Raw:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        mine_t = torch.as_tensor(_Raw(), device=f"cuda:{local_rank}")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out_t = shard.gather_records(mine_t, len(cases), world, dist)
        e1.record()
        torch.cuda.synchronize()
        gathered = e0.elapsed_time(e1)
        conv_all = out_t.clone().view(torch.int32).view(-1, 12)[:, 0].sum().item()
    else:
        conv_all = nconv
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = t.tolist()

    # ---- end-to-end arm through the public C-ABI call with host buffers
    e2e_steps = max(1, min(args.steps, 2))
    eng.solve_outages((optr, obr), want_v=False, want_types=False)   # warm
    barrier()
    w0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(e2e_steps):
        eng.solve_outages((optr, obr), want_v=False, want_types=False)
        st = eng.stats()
        h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
    barrier()
    e2e_s = time.perf_counter() - w0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()
    total_scen = per_rank * world

    # ---- CPU baseline (rank 0, N=1 only): oracle port on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncores = os.cpu_count() or 1
        ns = min(len(mine), args.cpu_sample or cpu_sample_size(g, ncores))
        rate, dt, its = cpu_reference_rate(g, vm, va, mine, ns, ncores)
        cpu = {"value": rate, "unit": UNIT, "cores": ncores, "kind": "port",
               "sample": f"first {ns} N-1 scenarios of the same workload, oracle port (numpy + SuperLU, one process per core), "
                         f"{dt:.1f} s, {its} NR iterations"}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
        nnzj, nnzlu, N = info["nnz_j_scalar"], info["nnz_lu_scalar"], 2 * info["m"]
        units = agg["scenario_iterations"]
        front_bytes_unit = 8.0 * nnzj + 8.0 * nnzlu + 16.0 * N
        traffic = None
        try:   # ncu dram__bytes_read+write of all k_front launches of one tick / scenarios in that tick (same grid)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_k_front_traffic.json")))
            if args.rows == 100 and args.cols == 100:
                traffic = tj["k_front_dram_bytes_per_unit"] * units / max(1, agg["factor_launches"])
        except Exception:
            pass
        ach_front = units * front_bytes_unit / (agg["factor_ms"] * 1e-3) / 1e9 if agg["factor_ms"] > 0 else 0.0
        ach_step = units * info["algorithmic_bytes_per_iteration"] / (dev_ms * 1e-3) / 1e9 if dev_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": total_scen * args.steps / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "buses": g.n, "branches": g.nbranch, "scenarios_per_gpu_per_step": per_rank,
                       "scenarios_total_per_step": total_scen, "tol": 1e-8, "max_iter": 30, "qlimits": True,
                       "l2_policy": "per-step working set (>= 20 MB per resident scenario x slots) far exceeds the 126 MB L2",
                       "ordering": info["ordering"], "slots": info["max_slots"], "nnz_j": nnzj, "nnz_lu": nnzlu,
                       "levels": info["n_levels"], "panels": info["n_panels"]},
            "converged": int(conv_all), "iterations_per_scenario": agg["mismatch_evaluations"] / max(1, per_rank * args.steps),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "clocks": clocks,
            "e2e": {"value": total_scen * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(agg["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_front (numeric multifrontal LU, all launches of one pass)",
                         "achieved": ach_front, "peak": hbm, "unit": "GB/s", "frac": ach_front / hbm, "traffic": traffic,
                         "algorithmic_bytes_per_launch": units * front_bytes_unit / max(1, agg["factor_launches"]),
                         "traffic_note": "bytes per launch: ncu dram read+write per scenario-iteration (profiles/r1_k_front_traffic.json) x units / launches; 5.8x the algorithmic bytes = the multifrontal update-matrix arena",
                         "peak_source": peak_src, "bytes_per_unit": front_bytes_unit, "units": units,
                         "avg_launch_ms": agg["factor_ms"] / max(1, agg["factor_launches"]), "launches": agg["factor_launches"],
                         "kernel_ms": agg["factor_ms"]},
            "roofline_step": {"achieved": ach_step, "peak": hbm, "unit": "GB/s", "frac": ach_step / hbm,
                              "bytes_per_unit": info["algorithmic_bytes_per_iteration"],
                              "scenario_iterations_per_s": units / (dev_ms * 1e-3) if dev_ms > 0 else 0.0},
            "kernel_ms": {k: agg[k] for k in ("factor_ms", "solve_ms", "mismatch_ms", "jacobian_ms", "other_ms")},
            "allgather_ms": gathered, "create_s": t_create,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args, rank, world):
    """Times the reference's own CPU implementation of the path (oracle port; Julia is not in the
    image) with every host core on a bounded sample of the same workload."""
    if rank != 0:
        return
    import sparlectra_oracle as O
    g, cases, per_rank, wname = workload(args, 1)
    vm, va, flat, base = O.solve_base(g)
    ncores = os.cpu_count() or 1
    ns = min(len(cases), args.cpu_sample or cpu_sample_size(g, ncores))
    rates = []
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_rate(g, vm, va, cases, ns, ncores)
    t_all = 0.0
    for _ in range(max(1, min(args.steps, 2))):
        rate, dt, its = cpu_reference_rate(g, vm, va, cases, ns, ncores)
        rates.append(rate)
        t_all += dt
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates),
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t_all / len(rates), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "buses": g.n, "branches": g.nbranch},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port",
                             "sample": f"{ns} N-1 scenarios per step, oracle port (numpy + SuperLU), one process per core"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
