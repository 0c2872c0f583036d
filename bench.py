#!/usr/bin/env python
"""Benchmark of the per-step simulation hot path (BASELINE.json: entity-updates/sec, device-timed).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Headline (`metric`, `value`, `roofline`, `e2e` - unchanged since round 1):
  N = 1  forest_fire 4096 x 4096, 1 replica (BASELINE.json configs[1]; examples/forest_fire.rs with the
         three systems + the FireStats aggregate series every step);
  N > 1  the row-split forest of configs[1]: 16384 cells per column, 2048 columns per GPU, one-column halo
         exchange per step fused into the step kernel (peer memory over NVLink; NCCL send/recv as fallback;
         N = 8 is exactly 16384 x 16384) - weak scaling.
A "step" is one simulation step of the whole grid.  `value` is cell-updates/s with the state resident in
HBM, every timed step started with a cold L2 (a 512 MB buffer is overwritten and half of it read back in between -
the L2 then holds clean foreign lines -, excluded from the timing; with N > 1 an untimed all-reduce then lines the ranks up again), timed with a CUDA event pair
around every step on the engine's stream (one incerto_sim_run_cold call), max over ranks.  `e2e` is the same metric
through the C ABI with host buffers: spawn from host arrays (H2D), run, read the FireStats series back -
best and median of 7 jobs (after 2 warm-up jobs) with a phase breakdown; for N > 1 the job is the row-split grid, every rank spawning its
own slab from host arrays.

Extra keys (round 2, VERDICT r01 "Next round" 1, 4, 5):
  `workloads`   every other BASELINE config on this GPU (N > 1: one world per GPU, value summed): coin_toss 2^28,
                forest_fire 16384^2, pandemic 10 M, pandemic_spatial 50 M, traders 1 M x 100, air_traffic_3d 5 M -
                device-timed run(n) (CUDA events on the engine's stream) and, where the state fits the L2, a cold-L2
                figure next to it; algorithmic bytes per update are the engine's own per-launch accounting of the
                SHIPPED layout (csrc: KTimer), fractions are against MEASURED_PEAKS.json;
  `replicas`    configs[2] / [4]: independent Monte Carlo replicas of pandemic 10 M and traders 1 M, a fixed number
                per GPU (weak scaling), two in flight per GPU, statistics all-reduced over NCCL;
  `partition_parity`, `replica_parity`, `global_aggregate_parity` (N > 1)  untimed correctness checks of the
                multi-GPU paths (row-split grid = undivided run; sharded + all-reduced replica statistics = one GPU
                running them all; cross-replica Median / Percentile by distributed radix select = numpy on the
                gathered columns; `air_partition_parity`: one airspace split by x with migrating aircraft = the
                undivided world); `air_one_world` (N > 1): the 5 M-aircraft airspace as ONE world over the N GPUs;
  `same_slab_1gpu_us`, `fixed_cost_us` (N = 1)  the N > 1 slab (2048 x 16384) alone on one GPU, and the cold floor
                of a 64 x 512 grid (launch + cold code / constant bank + event pair).
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
E2E_JOBS = 7
E2E_WARMUP_JOBS = 2   # the first jobs of a process also pay for what earlier sections left behind (graph / buffer teardown)
L2_BYTES = 126e6
WORKLOAD_N1 = "forest_fire 4096x4096, 1 replica (BASELINE.json configs[1])"


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic() -> dict:
    """DRAM bytes (read + write) per launch of every step kernel, from the committed `ncu --set full` captures of
    this round (profiles/r02_dram_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep files)."""
    p = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


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


class Ranks:
    """torch.distributed plumbing of the bench: barrier, max / sum over ranks, object gather."""

    def __init__(self, torch, dist, rank, world):
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def _reduce(self, v: float, op) -> float:
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, v: float) -> float:
        return self._reduce(v, self.dist.ReduceOp.SUM if self.dist else None)

    def gather(self, obj) -> list:
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def share_unique_id(self, ib) -> bytes:
        box = [ib.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(box, src=0)
        return bytes(box[0])


# ------------------------------------------------------------------------------------------------ e2e (host buffers)
def host_forest(np, F, W, H, x0, x1, overlays):
    """Host arrays of the columns [x0, x1) of a W x H forest (x-major, as forest_fire.rs:239-257 spawns them) followed
    by the overlay igniters: what a spawner closure would hand to `spawn_batch`."""
    xs, ys = np.meshgrid(np.arange(x0, x1, dtype=np.int16), np.arange(H, dtype=np.int16), indexing="ij")
    pos = np.concatenate([np.stack([xs.ravel(), ys.ravel()], 1), np.asarray(overlays, np.int16).reshape(-1, 2)])
    rng = np.random.default_rng((SEED & 0xFFFF) + x0)
    st = np.where(rng.random((x1 - x0) * H) < 0.7, F.CELL_HEALTHY, F.CELL_EMPTY).astype(np.uint8)
    st = np.concatenate([st, np.full(len(overlays), 3, np.uint8)])
    return np.ascontiguousarray(pos), np.ascontiguousarray(st)


def e2e_job(ib, F, W, H, device, host_pos, host_state, steps, ranks=None, uid=None, total_cells=None):
    """spawn from HOST buffers -> run -> read the series back; returns the phase times in seconds."""
    t0 = time.perf_counter()
    b = ib.SimulationBuilder(seed=SEED, device=device)
    pos = b.component("GridPosition2D", F.I16, lanes=2)
    cell = b.component("ForestCell.state", F.U8)
    if ranks is not None and ranks.world > 1:
        b.set_row_partition(ranks.rank, ranks.world)
    b.add_spatial_grid_2d(pos, cell, bounds=((0, 0), (W - 1, H - 1)))
    n = host_state.shape[0]
    b.add_entity_spawner(lambda s: s.spawn_batch(n, {pos: host_pos, cell: host_state}))
    b.add_systems(ib.FireSpread(pos, cell, 0.6, 3), ib.BurnProgression(pos, cell), ib.Regrowth(pos, cell, 0.001))
    b.record_aggregate_time_series(ib.FireStatsOf(cell), 1)
    t1 = time.perf_counter()
    sim = b.build(borrow_spawn_buffers=True)   # the spawner's staging arrays are read in place (no host-side copy)
    if uid is not None:
        sim.attach_comm(uid, ranks.rank, ranks.world)
    t2 = time.perf_counter()
    sim.run(steps)
    t3 = time.perf_counter()
    ts = sim.get_aggregate_time_series(ib.FireStatsOf(cell))     # partitioned: all-reduced over the ranks when read
    fs = sim.sample_aggregate(ib.FireStatsOf(cell))
    t4 = time.perf_counter()
    assert ts.len() == steps
    if total_cells is not None:
        v = ts.values()[-1]
        assert v.healthy_count + v.burning_count + v.burned_count + v.empty_count == total_cells, "FireStats of the job do not add up"
    elif ranks is None or ranks.world == 1:
        assert fs.total_cells == n
    sim.destroy()
    t5 = time.perf_counter()
    return {"total_s": t5 - t0, "builder_ms": (t1 - t0) * 1e3, "build_upload_ms": (t2 - t1) * 1e3,
            "run_ms": (t3 - t2) * 1e3, "readback_ms": (t4 - t3) * 1e3, "destroy_ms": (t5 - t4) * 1e3}


def run_e2e(ib, F, np, ranks, local_rank):
    world = ranks.world
    overlays = [[7, 9], [2000, 2000], [2047, 0]]
    if world == 1:
        W, H = 4096, 4096
        host_pos, host_state = host_forest(np, F, W, H, 0, W, overlays)
        h2d = W * H + 3 + 3 * 4
        job = f"spawn {W}x{H}+3 entities from host arrays, run({E2E_STEPS}), read the FireStats series"
    else:
        # the row-split job: every rank spawns the columns it needs (its slab + one halo column per interior side)
        # from host arrays; W = 2048 columns per GPU (x must fit the i16 position lanes at N = 8), H = 8192 cells
        W, H = 2048 * world, 8192
        b, e = ib.partition_range(W, ranks.rank, world)
        lo, hi = max(b - 1, 0), min(e + 1, W)
        host_pos, host_state = host_forest(np, F, W, H, lo, hi, overlays)
        h2d = (hi - lo) * H + 3 + 3 * 4
        job = (f"row-split {W}x{H}+3 forest over {world} GPUs: every rank spawns its slab (+ halo columns) from host arrays, "
               f"attaches the communicator, run({E2E_STEPS}), reads the all-reduced FireStats series")
    total_cells = W * H + 3
    # the cell states are what a job uploads: staged in PINNED host memory (a host that feeds a GPU allocates its staging
    # buffers that way); the positions are only read by the host-side dense-grid check
    try:
        import torch
        pinned = torch.empty(host_state.shape[0], dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[:] = host_state
        host_state, pinned_note = pinned.numpy(), "pinned"
    except Exception as ex_:  # no pinned allocation: pageable arrays work too
        pinned, pinned_note = None, f"pageable ({type(ex_).__name__})"
    jobs = []
    # one NCCL unique id for all jobs: the engine caches the communicator per id, so only the warm-up job pays for
    # ncclCommInitRank (a host that runs job after job creates its communicator once)
    uid = ranks.share_unique_id(ib) if world > 1 else None
    for j in range(E2E_JOBS + E2E_WARMUP_JOBS):       # the first jobs are warm-up
        ranks.barrier()
        r = e2e_job(ib, F, W, H, local_rank, host_pos, host_state, 32 if j == 0 else E2E_STEPS, ranks, uid, total_cells)
        r["total_s"] = ranks.max(r["total_s"])
        if j >= E2E_WARMUP_JOBS:
            jobs.append(r)
    tot = sorted(x["total_s"] for x in jobs)
    best = min(jobs, key=lambda x: x["total_s"])
    cells = W * H
    return {"value": cells * E2E_STEPS / tot[0], "unit": UNIT,
            "median": cells * E2E_STEPS / statistics.median(tot), "jobs": E2E_JOBS,
            "h2d_bytes_per_step": h2d / E2E_STEPS, "d2h_bytes_per_step": (E2E_STEPS * 48 + 40) / E2E_STEPS,
            "job": job, "host_buffers": pinned_note, "seconds": tot[0], "seconds_median": statistics.median(tot), "seconds_all": tot,
            "breakdown_best_job_ms": {k: round(v, 3) for k, v in best.items() if k.endswith("_ms")},
            "note": "value = best of the jobs, median beside it; phases are host wall-clock of this rank "
                    "(build_upload = incerto_builder_build: host-side dense-grid check beside cudaMalloc + H2D"
                    + (" + attach (cached NCCL communicator, CUDA IPC mapping of the neighbours' slabs)" if world > 1 else "") + "; run = 500 steps incl. synchronisation)"}


# --------------------------------------------------------------------------------------- the other BASELINE configs
def kernel_table(sim, base=None):
    out = {}
    for k in sim.kernel_stats():
        b = (base or {}).get(k["name"], (0, 0.0, 0.0))
        out[k["name"]] = (k["launches"] - b[0], k["algorithmic_bytes"] - b[1], k["device_ms"] - b[2])
    return out


def time_workload(name, make, steps, warm, ranks, peak, traffic, cold_state_bytes=None, note=None):
    """run(steps) on a warmed-up simulation: device time of the run (CUDA events on the engine's stream), the engine's
    algorithmic-byte accounting over the same launches, then a short profiled run for the per-kernel split."""
    sim, h = make()
    n = h.n
    sim.run(warm)
    base = {k: (v[0], v[1], v[2]) for k, v in kernel_table(sim).items()}
    ranks.barrier()
    sim.run(steps)
    ms, launches = sim.last_run_stats()
    tab = kernel_table(sim, base)
    alg_bytes = sum(v[1] for v in tab.values())
    ms_max = ranks.max(ms)
    out = {"entities": n, "steps": steps, "us_per_step": ms_max / steps * 1e3,
           "updates_per_s": ranks.sum(n * steps) / (ms_max * 1e-3),
           "algorithmic_bytes_per_update": alg_bytes / (n * steps), "launches_per_step": launches / steps}
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    out["achieved_GBps_per_gpu"] = gbs
    out["frac"] = gbs / peak
    # cold-L2 figure where the state would otherwise sit in the 126 MB L2 between steps
    if cold_state_bytes is not None and cold_state_bytes < 2 * L2_BYTES:
        ranks.barrier()
        csteps = max(20, min(steps, 40))
        sim.run_cold(csteps)
        cms, _ = sim.last_run_stats()
        cms = ranks.max(cms)
        out["cold_l2"] = {"us_per_step": cms / csteps * 1e3, "steps": csteps,
                          "frac": alg_bytes / (n * steps) * n / (cms / csteps * 1e-3) / 1e9 / peak,
                          "state_bytes": cold_state_bytes}
    # per-kernel split (events around every launch serialise the stream: not used for the figures above)
    sim.set_profiling(True)
    base = {k: (v[0], v[1], v[2]) for k, v in kernel_table(sim).items()}
    psteps = 15 if name == "pandemic_spatial" else 10     # one full quarantine wave
    sim.run(psteps)
    kern = {}
    dram = 0.0
    have_all = True
    for kname, (dl, db, dms) in kernel_table(sim, base).items():
        if dl <= 0 or dms <= 0:
            continue
        kern[kname] = {"us_per_launch": dms / dl * 1e3, "launches_per_step": dl / psteps,
                       "algorithmic_MB_per_launch": db / dl / 1e6, "frac": (db / (dms * 1e-3) / 1e9 / peak) if db else None}
        t = traffic.get(name, {}).get(kname)
        if t is not None:
            kern[kname]["dram_bytes_per_launch_ncu"] = t
            dram += t * dl / psteps
        elif db:
            have_all = False
    sim.set_profiling(False)
    out["kernels"] = kern
    out["dram_bytes_per_step"] = dram if (dram and have_all) else None
    if note:
        out["note"] = note
    sim.destroy()
    return out


def run_workloads(ib, ex, ranks, local_rank, peak, steps_scale=1.0):
    traffic = ncu_traffic()
    r = ranks.rank
    S = lambda k: max(20, int(k * steps_scale))
    W = {}
    W["coin_toss"] = time_workload(
        "coin_toss", lambda: ex.coin_toss(1 << 28, replica=r, device=local_rank), S(20), 5, ranks, peak, traffic,
        note="2^28 entities, CoinOdds aggregate series every step (forces one launch per step; a plain run(n) fuses steps)")
    W["forest_fire_16384"] = time_workload(
        "forest_fire_16384", lambda: ex.forest_fire(16384, 16384, replica=r, device=local_rank), S(50), 5, ranks, peak, traffic,
        note="16384^2 on one GPU: state 2 x 268 MB >> L2, no flush needed")
    W["pandemic"] = time_workload(
        "pandemic", lambda: ex.pandemic(10_000_000, 5000, replica=r, device=local_rank), S(100), 10, ranks, peak, traffic,
        cold_state_bytes=10_000_000 * 5 + 5000 * 5000 * 5,
        note="10 M agents on 5000^2 cells, one replica; the 50 MB of state + per-cell tables fit the L2 "
             "(as they do between the steps of each of the 1024 replicas of configs[2])")
    W["pandemic_spatial"] = time_workload(
        "pandemic_spatial", lambda: ex.pandemic_spatial(50_000_000, 6325, replica=r, device=local_rank), S(105), 30, ranks, peak, traffic,
        note="50 M people on 6325^2 cells; 105 steps = 7 quarantine waves of 15 steps (two heavy steps per wave)")
    W["traders"] = time_workload(
        "traders", lambda: ex.traders(1_000_000, 100, replica=r, device=local_rank), S(50), 100, ranks, peak, traffic,
        cold_state_bytes=1_000_000 * (56 + 16),
        note="1 M traders x 100 stocks, one replica; integer-pipe (Philox) bound, not HBM bound")
    W["air_traffic_3d"] = time_workload(
        "air_traffic_3d", lambda: ex.air_traffic_3d(5_000_000, 11547, 10, replica=r, device=local_rank), S(50), 10, ranks, peak, traffic,
        cold_state_bytes=5_000_000 * 10 + 5_000_000 * 8,
        note="5 M aircraft in 11547^2 x 10 cells")
    return W


def run_replicas(ib, ex, np, ranks, local_rank, peak):
    """configs[2] / [4a]: replicas r = rank, rank + N, ... run back to back by incerto_sim_reset, two in flight per GPU
    (one engine handle = one stream each); the per-replica statistics are summed over the ranks by the engine's
    NCCL all-reduce."""
    out = {}
    world, rank = ranks.world, ranks.rank
    for wl, per_gpu, steps, make, bpu in (
            ("pandemic", 64, 100, lambda rep: ex.pandemic(10_000_000, 5000, replica=rep, device=local_rank), 10),
            ("traders", 16, 50, lambda rep: ex.traders(1_000_000, 100, replica=rep, device=local_rank), 124)):
        conc = 2
        sims = [make(rank) for _ in range(conc)]
        n = sims[0][1].n
        if world > 1:
            sims[0][0].attach_comm(ranks.share_unique_id(ib), rank, world)
        for sim, _ in sims:
            sim.run(5)
        replicas = per_gpu * world
        mine = list(range(rank, replicas, world))
        u64 = np.zeros(2, np.uint64)
        f64 = np.zeros(1, np.float64)
        ranks.barrier()
        ms = 0.0
        for w0 in range(0, len(mine), conc):
            wave = mine[w0:w0 + conc]
            for (sim, _), rep in zip(sims, wave):
                sim.reset(rep)
            for (sim, _), rep in zip(sims, wave):
                sim.run_async(steps)
            for (sim, _), rep in zip(sims, wave):
                sim.synchronize()
            ms += max(sims[0][0].span_ms(sims[k][0]) for k in range(len(wave)))
            for (sim, h), rep in zip(sims, wave):
                if wl == "pandemic":
                    u64 += np.array([sim.count(ib.With(h.person)), sim.count(ib.With(h.infected))], np.uint64)
                else:
                    f64[0] += sim.sample_aggregate(ib.Mean(h.tc.net_worth))
        ms = ranks.max(ms)
        if world > 1:
            u64 = sims[0][0].allreduce_u64(u64, 0)
            f64 = sims[0][0].allreduce_f64(f64, 0)
        value = n * steps * replicas / (ms * 1e-3)
        o = {"updates_per_s": value, "replicas": replicas, "replicas_per_gpu": per_gpu, "in_flight_per_gpu": conc,
             "entities_per_replica": n, "steps_per_replica": steps, "ms_max_over_ranks": ms, "scaling": "weak",
             "us_per_replica_step": ms * 1e3 / (len(mine) * steps),
             "frac_per_gpu": value * bpu / world / 1e9 / peak, "algorithmic_bytes_per_update": bpu}
        if wl == "pandemic":
            o["survivors_all_replicas"], o["infected_all_replicas"] = int(u64[0]), int(u64[1])
        else:
            o["mean_net_worth_over_replicas"] = float(f64[0]) / replicas
        out[wl] = o
        for sim, _ in sims:
            sim.destroy()
    return out


# ------------------------------------------------------------------------------------------- multi-GPU parity checks
def check_partition_parity(ib, ex, np, ranks, local_rank):
    """A row-split grid small enough for the undivided run on rank 0 (512 N columns x 1024 cells, 64 steps, dense
    enough to burn across every cut): the slabs and the all-reduced FireStats series must equal the undivided run."""
    world, rank = ranks.world, ranks.rank
    W, H, steps = 512 * world, 1024, 64
    kw = dict(density=0.9, p_regrow=0.002, fires=16 * world)
    sim, h = ex.forest_fire(W, H, device=local_rank, partition=(rank, world), **kw)
    sim.attach_comm(ranks.share_unique_id(ib), rank, world)
    sim.run(steps)
    b, e = ib.partition_range(W, rank, world)
    slab = sim.read_component(h.cell)[:(e - b) * H].copy()
    ts = sim.get_aggregate_time_series(h.stats)                     # collective: every rank reads it
    ser = np.array([[v.healthy_count, v.burning_count, v.burned_count, v.empty_count, v.total_cells] for v in ts.values()], np.int64)
    slabs = ranks.gather(slab)
    ok = True
    if rank == 0:
        whole, hw = ex.forest_fire(W, H, device=local_rank, **kw)
        whole.run(steps)
        want = whole.read_component(hw.cell)[:W * H]
        ts0 = whole.get_aggregate_time_series(hw.stats)
        ser0 = np.array([[v.healthy_count, v.burning_count, v.burned_count, v.empty_count, v.total_cells] for v in ts0.values()], np.int64)
        ok = bool(np.array_equal(np.concatenate(slabs), want)) and bool(np.array_equal(ser, ser0)) and int(ser0[-1, 2]) > 0
        whole.destroy()
    sim.destroy()
    return bool(ranks.max(0.0 if ok else 1.0) == 0.0)


def check_replica_parity(ib, ex, np, ranks, local_rank):
    """2 N small replicas (pandemic 200 k x 20 steps, traders 20 k x 20) dealt to the ranks, statistics all-reduced by
    the engine over NCCL, against rank 0 running all of them alone."""
    world, rank = ranks.world, ranks.rank
    R = 2 * world
    sim, h = ex.pandemic(200_000, 707, replica=rank, device=local_rank)
    sim.attach_comm(ranks.share_unique_id(ib), rank, world)
    tsim, th = ex.traders(20_000, 100, replica=rank, device=local_rank)

    def one(rep):
        sim.reset(rep)
        sim.run(20)
        tsim.reset(rep)
        tsim.run(20)
        return (np.array([sim.count(ib.With(h.person)), sim.count(ib.With(h.infected))], np.uint64),
                np.array([tsim.sample_aggregate(ib.Mean(th.tc.net_worth))], np.float64))
    u, f = np.zeros(2, np.uint64), np.zeros(1, np.float64)
    for rep in range(rank, R, world):
        a, b = one(rep)
        u, f = u + a, f + b
    u = sim.allreduce_u64(u, 0)
    f = sim.allreduce_f64(f, 0)
    ok = True
    if rank == 0:
        u0, f0 = np.zeros(2, np.uint64), np.zeros(1, np.float64)
        for rep in range(R):
            a, b = one(rep)
            u0, f0 = u0 + a, f0 + b
        ok = bool(np.array_equal(u, u0)) and abs(float(f[0]) - float(f0[0])) <= 1e-12 * abs(float(f0[0]))
    sim.destroy()
    tsim.destroy()
    return bool(ranks.max(0.0 if ok else 1.0) == 0.0)


def check_air_partition_parity(ib, ex, np, ranks, local_rank):
    """ONE air_traffic_3d world (400 k aircraft at the reference density) split by x over the N GPUs -
    aircraft migrate between the ranks, ghosts cross the cuts, all over NCCL - stepped 40 times: every aircraft's
    position (gathered by uid) and the all-reduced conflict series must equal the undivided world on rank 0."""
    world, rank = ranks.world, ranks.rank
    n, size, steps = 400_000, 3266, 40
    sim, h = ex.air_traffic_3d(n, size, device=local_rank, partition=(rank, world))
    sim.attach_comm(ranks.share_unique_id(ib), rank, world)
    sim.run(steps)
    pos, uids = sim.read_component(h.position, with_uids=True)
    ts = sim.get_aggregate_time_series(h.conflicts)                  # collective: summed over the ranks, then halved
    ser = np.array(ts.values(), dtype=np.int64)
    parts = ranks.gather((uids, pos))
    ok = True
    migrated = 0
    if rank == 0:
        whole, hw = ex.air_traffic_3d(n, size, device=local_rank)
        whole.run(steps)
        want = whole.read_component(hw.position)
        ser0 = np.array(whole.get_aggregate_time_series(hw.conflicts).values(), dtype=np.int64)
        got = np.full_like(want, -1)
        seen = np.zeros(n, np.int32)
        for u, p in parts:
            got[u] = p
            seen[u] += 1
        ok = bool((seen == 1).all()) and bool(np.array_equal(got, want)) and bool(np.array_equal(ser, ser0)) and int(ser0.max()) > 0
        whole.destroy()
    sim.destroy()
    return bool(ranks.max(0.0 if ok else 1.0) == 0.0)


def run_air_one_world(ib, ex, ranks, local_rank, peak):
    """The 5 M-aircraft airspace of configs[4b] as ONE world over the N GPUs (strong scaling of a single world; every rank
    keeps the full table and steps the aircraft whose x lies in its slab)."""
    world, rank = ranks.world, ranks.rank
    n, size, steps = 5_000_000, 11547, 50
    sim, h = ex.air_traffic_3d(n, size, device=local_rank, partition=(rank, world))
    sim.attach_comm(ranks.share_unique_id(ib), rank, world)
    sim.run(10)
    ranks.barrier()
    sim.run(steps)
    ms, launches = sim.last_run_stats()
    ms = ranks.max(ms)
    conflicts = sim.sample_aggregate_global(h.conflicts)
    owned = ranks.sum(float(sim.sample_aggregate(ib.Count(h.target))))
    sim.destroy()
    return {"entities": n, "steps": steps, "us_per_step": ms / steps * 1e3, "updates_per_s": n * steps / (ms * 1e-3),
            "launches_per_step": launches / steps, "conflicts_last_step": int(conflicts), "aircraft_owned_by_all_ranks": int(owned),
            "note": "one 11547^2 x 10 airspace split by x; per step: ghosts + candidates + step kernels on the owned rows, one "
                    "fixed-size NCCL send/recv pair per neighbour (migrated aircraft + boundary-column ghosts), unpack"}


def check_global_aggregate_parity(ib, ex, np, ranks, local_rank):
    """SURVEY §8 (f) 3, cross-replica quantiles: every rank holds one traders replica (200 k traders, 30 steps); Median,
    Percentile<90>, Minimum, Maximum, Count and Mean of the net worth over the traders of ALL ranks through
    incerto_sim_sample_aggregate_global (distributed radix select: per digit pass one NCCL all-reduce of the 256-bin
    histogram) against numpy on the gathered columns."""
    world, rank = ranks.world, ranks.rank
    sim, h = ex.traders(200_000 + 1000 * rank, 100, replica=rank, device=local_rank)   # ragged rank sizes
    sim.attach_comm(ranks.share_unique_id(ib), rank, world)
    sim.run(30)
    nw = h.tc.net_worth
    got = {"count": sim.sample_aggregate_global(ib.Count(nw)), "min": sim.sample_aggregate_global(ib.Minimum(nw)),
           "max": sim.sample_aggregate_global(ib.Maximum(nw)), "median": sim.sample_aggregate_global(ib.Median(nw)),
           "p90": sim.sample_aggregate_global(ib.Percentile(nw, 90)), "mean": sim.sample_aggregate_global(ib.Mean(nw))}
    cols = ranks.gather(sim.read_component(nw))
    ok = True
    if rank == 0:
        allv = np.sort(np.concatenate(cols))
        n = len(allv)
        ok = (int(got["count"]) == n and got["min"] == allv[0] and got["max"] == allv[-1] and got["median"] == allv[n // 2]
              and got["p90"] == allv[int(np.floor(90 / 100.0 * n))] and abs(got["mean"] - float(allv.mean())) <= 1e-9 * abs(float(allv.mean())))
    sim.destroy()
    return bool(ranks.max(0.0 if ok else 1.0) == 0.0)


# --------------------------------------------------------------------------------------------------- CPU legs
def cpu_baseline_sample(threads: int, seconds_target: float = 12.0):
    """The oracle (C++ restatement of the Bevy systems: hash-map grid, one serial pass per system) timed on
    the host cores on a bounded sample of the workload."""
    from oracle.lib import load
    lib = load()
    W = H = 1024
    t0 = time.perf_counter()
    lib.oracle_forest_bench(SEED, W, H, 2, threads)
    t2 = time.perf_counter() - t0
    t0 = time.perf_counter()
    lib.oracle_forest_bench(SEED, W, H, 10, threads)
    t10 = time.perf_counter() - t0
    per_step = max((t10 - t2) / 8.0, 1e-4)          # spawn + hash-grid build are in both: the difference is 8 steps
    steps = max(20, min(600, int((seconds_target - t2) / per_step)))
    t0 = time.perf_counter()
    done = lib.oracle_forest_bench(SEED, W, H, steps, threads)
    dt = time.perf_counter() - t0
    return done / dt, f"forest_fire {W}x{H}, {steps} steps x {threads} replica(s) incl. spawn + grid build, {dt:.1f} s", steps, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the GPU arm's own configuration.  The
    reference itself (Rust on Bevy) cannot be built here, so this is the oracle port (`kind: port`).  `value` is the
    4096 x 4096 grid, 1 replica, one host thread - the three forest systems all take ForestCell mutably, so Bevy runs
    them (and each system's loop) serially: one core is all the reference can use on this configuration.  The generous
    multi-core figure (independent 1024^2 replicas on every host thread) is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    from oracle.forest import ForestOracle
    from oracle.lib import load
    lib = load()
    W = H = 4096
    o = ForestOracle(SEED, 0, W, H).spawn()
    warm = max(1, min(args.warmup, 2))      # the first step builds the hash-map SpatialGrid (PreUpdate of step 1)
    o.run(warm)
    # bounded: at most ~2 minutes of timed steps (a step of this grid is ~2.5 s on one core); ms_per_step is per
    # simulated step either way
    t0 = time.perf_counter()
    o.run(1)
    t1 = time.perf_counter() - t0
    timed = max(1, min(args.steps, int(120.0 / max(t1, 1e-3))))
    if timed > 1:
        o.run(timed - 1)
    dt = time.perf_counter() - t0
    value = W * H * timed / dt
    del o
    t0 = time.perf_counter()
    done = lib.oracle_forest_bench(SEED, 1024, 1024, 10, cores)
    dtm = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / timed * 1e3, "timed_steps": timed, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_N1, "systems": "fire_spread + burn_progression + regrowth + FireStats series/1",
                   "seed": hex(SEED), "parallelism": "1 host thread (Bevy serialises the three conflicting systems)",
                   "sample": f"{W}x{H} grid, 1 replica, {warm} untimed + {timed} timed steps"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"forest_fire {W}x{H} x 1 replica x {timed} steps ({dt:.1f} s)"},
        "all_host_threads": {"value": done / dtm, "unit": UNIT, "cores": cores,
                             "sample": f"forest_fire 1024x1024 x {cores} independent replicas x 10 steps incl. spawn + grid build ({dtm:.1f} s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = Rust/Bevy, not buildable here (no cargo/rustc); this is oracle/ (kind 'port')",
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------------- main
def cold_forest_us(ex, W, H, steps, device):
    sim, _ = ex.forest_fire(W, H, device=device)
    sim.run_cold(5)
    sim.run_cold(steps)
    ms, _ = sim.last_run_stats()
    sim.run(64)
    sim.run(64)
    wms, _ = sim.last_run_stats()
    sim.destroy()
    return ms / steps * 1e3, wms / 64 * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true")
    ap.add_argument("--no-replicas", action="store_true")
    ap.add_argument("--only-headline", action="store_true", help="skip every extra section (ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.only_headline:
        args.no_cpu_baseline = args.no_e2e = args.no_workloads = args.no_replicas = True
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
    ranks = Ranks(torch, dist, rank, world)

    import incerto_b200 as ib          # fails loudly if the CUDA engine is missing
    from incerto_b200 import _ffi as F
    from incerto_b200 import examples as ex

    if world == 1:
        W, H = 4096, 4096
        workload = WORKLOAD_N1
    else:
        W, H = 2048 * world, 16384
        workload = (f"forest_fire {W}x{H} row-split over {world} GPUs (2048 columns of 16384 cells per GPU, "
                    "1-column halo exchange per step; configs[1], N = 8 is 16384^2)")
    sim, hd = ex.forest_fire(W, H, device=local_rank, partition=(rank, world))
    if world > 1:
        sim.attach_comm(ranks.share_unique_id(ib), rank, world)
    cells = W * H

    # ---- warm-up
    sim.run_cold(args.warmup)
    ranks.barrier()

    # ---- timed: K steps, each from a cold L2 (flush kernel in between, not timed), all enqueued back to back on
    # the engine's stream with a CUDA event pair around every step; the device time is the sum over the steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sim.run_cold(args.steps)
    total_ms, launches = sim.last_run_stats()
    ranks.barrier()
    clocks = sampler.stop() if rank == 0 else {}
    kernel_ms = total_ms / args.steps   # this rank's own average step (the roofline of its kernel)
    total_ms = ranks.max(total_ms)
    value = cells * args.steps / (total_ms * 1e-3)

    # ---- the same steps left L2-resident (what a user's run(n) sees): one call, graph replay
    warm_steps = max(args.steps, 64) if world == 1 else args.steps
    sim.run(warm_steps)
    warm_ms, _ = sim.last_run_stats()
    warm_ms = ranks.max(warm_ms)
    value_warm = cells * warm_steps / (warm_ms * 1e-3)

    # ---- roofline of the dominant kernel (forest_step): algorithmic 2 B per cell-update
    peak, peak_src = measured_peak_gbs()
    per_rank_cells = cells / world
    achieved = 2.0 * per_rank_cells / (kernel_ms * 1e-3) / 1e9
    traffic = ncu_traffic().get("forest_fire_4096", {}).get("forest_step") if world == 1 else None
    traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch under `ncu --set full` (profiles/r02_ncu_forestbench.txt): "
                    "the 16.8 MB input buffer is read from DRAM; the 16.8 MB written stay in the 126 MB L2 until later launches evict "
                    "them, so the write-back of a launch's output falls outside its own ncu window (at 16384^2, where the output "
                    "cannot stay, ncu counts read + write = 495 MB for 537 MB algorithmic)")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note if traffic else None,
                "kernel": "forest_step_kernel", "algorithmic_bytes_per_unit": 2,
                "units_per_launch": per_rank_cells, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "l2_resident": {"achieved": 2.0 * per_rank_cells * warm_steps / (warm_ms * 1e-3) / 1e9,
                                "note": "same kernel, state left in L2 between steps (CUDA-graph replay); not an HBM figure"}}

    halo = None
    if world > 1:
        names = {k["name"] for k in sim.kernel_stats() if k["launches"]}
        halo = ("fused into the step kernel: boundary CTAs pull the neighbours' boundary columns over NVLink (CUDA IPC), "
                "mailbox stamps" if "forest_step_fused_halo" in names
                else "peer-memory stores over NVLink (CUDA IPC) + mailbox stamps, separate kernels" if "forest_halo_push_peer" in names
                else "NCCL send/recv")
    sim.destroy()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "systems": "fire_spread + burn_progression + regrowth + FireStats series/1",
                   "l2": "flushed between timed steps (512 MB written, then 256 MB of it read back so that the L2 holds clean lines; untimed)", "seed": hex(SEED),
                   "parallelism": "1 GPU" if world == 1 else f"row partition x{world}, halo transport: {halo}"},
        "roofline": roofline,
        "value_l2_resident": value_warm,
        "gpu_launches": int(launches),
    }
    if rank == 0:
        line["clocks"] = clocks

    # ---- N = 1: the N > 1 slab alone on this GPU (so weak-scaling efficiency is computable from driver data) and
    # the fixed cost of a cold step
    if world == 1 and not args.only_headline:
        slab_cold, slab_warm = cold_forest_us(ex, 2048, 16384, max(args.steps, 20), local_rank)
        line["same_slab_1gpu_us"] = slab_cold
        line["same_slab_1gpu_l2_resident_us"] = slab_warm
        fixed_cold, fixed_warm = cold_forest_us(ex, 64, 512, max(args.steps, 20), local_rank)
        line["fixed_cost_us"] = fixed_cold
        line["fixed_cost_l2_resident_us"] = fixed_warm
        line["value_above_fixed_cost"] = {
            "us": total_ms / args.steps * 1e3 - fixed_cold,
            "GBps": 2.0 * cells / ((total_ms / args.steps * 1e3 - fixed_cold) * 1e-6) / 1e9,
            "note": "what the 16.8 M cells cost beyond a 64 x 512 grid's cold step (launch, cold code / constant bank, event pair)"}

    # ---- N = 1: the same 4096^2 grid as 12 independent replicas in flight (one engine handle = one stream each,
    # run_async): 12 x 33.5 MB of state do not fit the L2, so every step streams its cells from HBM, but launch
    # latency, cold code and the event pair of a lone cold step are overlapped / amortised - the Monte Carlo use of
    # the same kernel, next to the single-replica headline above
    if world == 1 and not args.only_headline:
        R, rsteps = 12, 64
        sims = [ex.forest_fire(4096, 4096, replica=r, device=local_rank)[0] for r in range(R)]
        for sm in sims:
            sm.run(16)
        for sm in sims:
            sm.synchronize()
        for sm in sims:
            sm.run_async(rsteps)
        for sm in sims:
            sm.synchronize()
        span = max(sims[0].span_ms(sm) for sm in sims)
        gbs = 2.0 * 4096 * 4096 * R * rsteps / (span * 1e-3) / 1e9
        line["replica_stream"] = {"replicas_in_flight": R, "steps_per_replica": rsteps, "span_ms": span,
                                  "cell_updates_per_s": 4096 * 4096 * R * rsteps / (span * 1e-3),
                                  "us_per_replica_step": span * 1e3 / (R * rsteps), "achieved_GBps": gbs, "frac": gbs / peak,
                                  "note": "12 independent 4096^2 replicas (403 MB of state > L2) stepped concurrently, CUDA-graph "
                                          "replay per handle; device span from the first handle's start to the last one's end"}
        for sm in sims:
            sm.destroy()

    # ---- multi-GPU correctness (untimed)
    if world > 1 and not args.only_headline:
        line["partition_parity"] = check_partition_parity(ib, ex, np, ranks, local_rank)
        line["replica_parity"] = check_replica_parity(ib, ex, np, ranks, local_rank)
        line["global_aggregate_parity"] = check_global_aggregate_parity(ib, ex, np, ranks, local_rank)
        line["air_partition_parity"] = check_air_partition_parity(ib, ex, np, ranks, local_rank)

    # ---- e2e through the C ABI with host buffers
    if not args.no_e2e:
        line["e2e"] = run_e2e(ib, F, np, ranks, local_rank)

    # ---- every other BASELINE config + the replica-sharded Monte Carlo mode
    if not args.no_workloads:
        line["workloads"] = run_workloads(ib, ex, ranks, local_rank, peak)
    if not args.no_replicas:
        line["replicas"] = run_replicas(ib, ex, np, ranks, local_rank, peak)
    if world > 1 and not args.no_workloads:
        line["air_one_world"] = run_air_one_world(ib, ex, ranks, local_rank, peak)

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v1, sample1, _, _ = cpu_baseline_sample(1)
        line["cpu_baseline"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample1,
                                "note": "Bevy runs these three conflicting systems serially: 1 core is what the reference uses"}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
