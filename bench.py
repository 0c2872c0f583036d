#!/usr/bin/env python
"""Benchmark of the per-step simulation hot path (BASELINE.json: entity-updates/sec, device-timed).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload):
  N = 1  forest_fire 4096 x 4096, 1 replica (BASELINE.json configs[1]; examples/forest_fire.rs with the
         three systems + the FireStats aggregate series every step);
  N > 1  the row-split forest of configs[1]: 16384 cells per column, 2048 columns per GPU, one-column halo
         exchange per step fused into the step kernel (peer memory over NVLink; NCCL send/recv as fallback;
         N = 8 is exactly 16384 x 16384) — weak scaling.
A "step" is one simulation step of the whole grid.  `value` is cell-updates/s with the state resident in
HBM, every timed step started with a cold L2 (a 512 MB buffer is overwritten in between, excluded from
the timing; with N > 1 an untimed all-reduce then lines the ranks up again), timed with a CUDA event pair
around every step on the engine's stream (one incerto_sim_run_cold call), max over ranks.  `e2e` is the same metric
through the C ABI with host buffers: spawn from host arrays (H2D), run, read the FireStats series back.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x1CE27002025
METRIC = "entity-updates/sec (device-timed)"
UNIT = "cell-updates/s"
E2E_STEPS = 500  # SIMULATION_STEPS of examples/forest_fire.rs:32


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_forest(ib, F, W, H, device, rank=0, world=1, series=1):
    b = ib.SimulationBuilder(seed=SEED, device=device)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    if world > 1:
        b.set_row_partition(rank, world)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    b.add_entity_spawner(ib.DeviceSpawner(F.SPAWN_FOREST_GRID, F.SpawnForestParams(pos.handle, cell.handle, W, H, 0.7, 3, 3)))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
    if series:
        b.record_aggregate_time_series(ib.FireStatsOf(cell), series)
    return b.build(), pos, cell


def e2e_job(ib, F, np, W, H, device, host_state, host_pos, steps):
    """spawn from HOST buffers -> run -> read the series back; returns (seconds, h2d bytes, d2h bytes)."""
    t0 = time.perf_counter()
    b = ib.SimulationBuilder(seed=SEED, device=device)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    n = host_state.shape[0]
    b.add_entity_spawner(lambda s: s.spawn_batch(n, {pos: host_pos, cell: host_state}))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
    b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
    sim = b.build(borrow_spawn_buffers=True)   # the spawner's staging arrays are read in place (no host-side copy)
    sim.run(steps)
    ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))
    fs = sim.sample_aggregate(ib.FireStatsOf(cell))
    assert ts.len() == steps and fs.total_cells == n
    dt = time.perf_counter() - t0
    sim.destroy()
    # the dense prefix keeps positions implicit on the device: only the cell states (and the 3 overlay
    # positions) cross the bus; the series comes back as steps * 5 u64 + their counts
    return dt, int(W * H + 3 + 3 * 4), int(steps * 48 + 40)


def cpu_baseline_sample(threads: int, seconds_target: float = 12.0):
    """The oracle (C++ restatement of the Bevy systems: hash-map grid, one serial pass per system) timed on
    the host cores on a bounded sample of the workload."""
    from oracle.lib import load
    lib = load()
    W = H = 1024
    t0 = time.perf_counter()
    lib.oracle_forest_bench(SEED, W, H, 2, threads)
    per2 = time.perf_counter() - t0
    steps = max(4, min(200, int(seconds_target / max(per2 / 2, 1e-3))))
    t0 = time.perf_counter()
    done = lib.oracle_forest_bench(SEED, W, H, steps, threads)
    dt = time.perf_counter() - t0
    return done / dt, f"forest_fire {W}x{H}, {steps} steps x {threads} replica(s) incl. spawn + grid build, {dt:.1f} s", steps, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself (Rust on
    Bevy) cannot be built here, so this is the oracle port on all host threads (independent replicas)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    from oracle.lib import load
    lib = load()
    W = H = 1024
    for _ in range(max(args.warmup, 0) and 1):
        lib.oracle_forest_bench(SEED, 256, 256, 2, cores)
    t0 = time.perf_counter()
    done = lib.oracle_forest_bench(SEED, W, H, args.steps, cores)
    dt = time.perf_counter() - t0
    value = done / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "forest_fire (examples/forest_fire.rs), CPU port of the Bevy systems",
                   "sample": f"{W}x{H} grid x {cores} independent replicas, {args.steps} steps each, spawn + grid build included"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"forest_fire {W}x{H} x {cores} replicas x {args.steps} steps ({dt:.1f} s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = Rust/Bevy, not buildable here (no cargo/rustc); this is oracle/ (kind 'port') on all host threads",
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import incerto_b200 as ib          # fails loudly if the CUDA engine is missing
    from incerto_b200 import _ffi as F

    if world == 1:
        W, H = 4096, 4096
        workload = "forest_fire 4096x4096, 1 replica (BASELINE.json configs[1])"
    else:
        W, H = 2048 * world, 16384
        workload = (f"forest_fire {W}x{H} row-split over {world} GPUs (2048 columns of 16384 cells per GPU, "
                    "1-column halo exchange per step; configs[1], N = 8 is 16384^2)")
    sim, pos, cell = build_forest(ib, F, W, H, local_rank, rank, world, series=1)
    if world > 1:
        uid = [ib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.attach_comm(uid[0], rank, world)
    cells = W * H

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up
    sim.run_cold(args.warmup)
    barrier()

    # ---- timed: K steps, each from a cold L2 (flush kernel in between, not timed), all enqueued back to back on
    # the engine's stream with a CUDA event pair around every step; the device time is the sum over the steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sim.run_cold(args.steps)
    total_ms, launches = sim.last_run_stats()
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    kernel_ms = total_ms / args.steps   # this rank's own average step (the roofline of its kernel)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = cells * args.steps / (total_ms * 1e-3)

    # ---- the same steps left L2-resident (what a user's run(n) sees): one call, graph replay
    warm_steps = max(args.steps, 64) if world == 1 else args.steps
    sim.run(warm_steps)
    warm_ms, _ = sim.last_run_stats()
    if dist is not None:
        t = torch.tensor([warm_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        warm_ms = float(t.item())
    value_warm = cells * warm_steps / (warm_ms * 1e-3)

    # ---- roofline of the dominant kernel (forest_step): algorithmic 2 B per cell-update
    peak, peak_src = measured_peak_gbs()
    per_rank_cells = cells / world
    achieved = 2.0 * per_rank_cells / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "forest_step_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch_4096")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "forest_step_kernel", "algorithmic_bytes_per_unit": 2,
                "units_per_launch": per_rank_cells, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "l2_resident": {"achieved": 2.0 * per_rank_cells * warm_steps / (warm_ms * 1e-3) / 1e9,
                                "note": "same kernel, state left in L2 between steps (CUDA-graph replay); not an HBM figure"}}

    halo = None
    if world > 1:
        names = {k["name"] for k in sim.kernel_stats() if k["launches"]}
        halo = ("fused into the step kernel: boundary warps store into the neighbours' slabs over NVLink (CUDA IPC), "
                "mailbox stamps" if "forest_step_fused_halo" in names
                else "peer-memory stores over NVLink (CUDA IPC) + mailbox stamps, separate kernels" if "forest_halo_push_peer" in names
                else "NCCL send/recv")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "systems": "fire_spread + burn_progression + regrowth + FireStats series/1",
                   "l2": "flushed between timed steps (512 MB written, untimed)", "seed": hex(SEED),
                   "parallelism": "1 GPU" if world == 1 else f"row partition x{world}, halo transport: {halo}"},
        "roofline": roofline,
        "value_l2_resident": value_warm,
        "gpu_launches": int(launches),
    }
    if rank == 0:
        line["clocks"] = clocks

    # ---- e2e through the C ABI with host buffers (rank 0's share; replicas of it on the other ranks)
    if not args.no_e2e:
        eW, eH = (4096, 4096)
        xs, ys = np.meshgrid(np.arange(eW, dtype=np.int16), np.arange(eH, dtype=np.int16), indexing="ij")
        host_pos = np.concatenate([np.stack([xs.ravel(), ys.ravel()], 1), np.array([[7, 9], [2000, 2000], [4095, 0]], np.int16)])
        rng = np.random.default_rng(SEED & 0xFFFF)
        host_state = np.where(rng.random(eW * eH) < 0.7, F.CELL_HEALTHY, F.CELL_EMPTY).astype(np.uint8)
        host_state = np.concatenate([host_state, np.array([3, 3, 3], np.uint8)])
        e2e_job(ib, F, np, eW, eH, local_rank, host_state, host_pos, 32)   # warm-up
        barrier()
        dt, h2d, d2h = e2e_job(ib, F, np, eW, eH, local_rank, host_state, host_pos, E2E_STEPS)
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        line["e2e"] = {"value": world * eW * eH * E2E_STEPS / dt, "unit": UNIT,
                       "h2d_bytes_per_step": h2d / E2E_STEPS, "d2h_bytes_per_step": d2h / E2E_STEPS,
                       "job": f"spawn {eW}x{eH}+3 entities from host arrays, run({E2E_STEPS}), read the FireStats series"
                              + ("" if world == 1 else f" (one independent job per GPU, x{world})"),
                       "seconds": dt}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v1, sample1, _, _ = cpu_baseline_sample(1)
        line["cpu_baseline"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample1,
                                "note": "Bevy runs these three conflicting systems serially: 1 core is what the reference uses"}
    if rank == 0:
        print(json.dumps(line))
    sim.destroy()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
