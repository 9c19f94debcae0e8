#!/usr/bin/env python3
"""bench.py -- collision-checked states/sec (BASELINE.json metric) on config C2:
Panda 7-DoF, 1 M random states per GPU vs a 500 k-point cloud scene, 2 m^3 field at 1 cm (200^3 voxels).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--states n] [--no-edt]

One "step" = one pass of the hot path over one batch: FK + self-field + world-field + intra-group sphere checks of
every state of the batch -> validity bitmap (PlanningScene::checkCollision semantics).  N > 1 (torchrun, one rank
per GPU): the states are sharded, the fields replicated, and each step ends with an NCCL all-gather of the bitmap
slices; `value` is the whole-job aggregate.  Prints ONE JSON line on rank 0.

Timed regions
  value  device-resident inputs; CUDA events on the launching stream around each step; L2 flushed between steps.
  e2e    the reference-facing C-ABI call b200sv_check_batch with HOST buffers (pinned): H2D of the states, kernels,
         D2H of the bitmap, all inside the timed region (wall clock around the synchronous call).
  cpu_baseline / --impl reference: the oracle port of the reference's CPU path (oracle/csrc/oracle.c, lean flavour:
         sphere centres + centre voxel only) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C2"
METRIC = "collision_checked_states_per_sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--states", type=int, default=0, help="states per GPU per step (default: the config's 1 000 000)")
    ap.add_argument("--no-edt", action="store_true", help="skip the C4 EDT rebuild measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity check against the oracle")
    ap.add_argument("--fp64", action="store_true", help="force every state through the FP64 kernel")
    ap.add_argument("--workload", default=WORKLOAD, choices=["C2", "C3", "C3D", "C3W"],
                    help="C2 (default, the config the metric is quoted on), C3 (dual arm, 10 M states over all ranks) C3D (C3 with the "
                         "dual arm at 280 spheres) or C3W (C3 with the whole_body group: planar base, 18 variables)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_env(cfg, cloud):
    from oracle.collision_model import OracleEnv
    from oracle.robot_model import RobotModel
    res = os.path.join(ROOT, "moveit2_b200", "resources")
    files = {"panda": ("panda_spheres.urdf", "panda.srdf"), "dualarm": ("dualarm_spheres.urdf", "dualarm.srdf"),
             "dualarm_dense": ("dualarm_dense_spheres.urdf", "dualarm.srdf"),
             "panda_primitives": ("panda_primitives.urdf", "panda.srdf")}[cfg["robot"]]
    with open(os.path.join(res, files[0])) as f:
        urdf = f.read()
    with open(os.path.join(res, files[1])) as f:
        srdf = f.read()
    env = OracleEnv(RobotModel(urdf, srdf), cfg["group"], cfg["size"], cfg["origin"], cfg["resolution"],
                    cfg["max_distance"])
    t = time.perf_counter()
    env.world.add_points(cloud)
    env.world_build_s = time.perf_counter() - t  # the reference's bucket-queue propagation (oracle port, one thread)
    return env


def cpu_rate(env, q, seconds=10.0):
    """states/s of the oracle port on all host threads, on a bounded sample sized for ~`seconds` of wall time."""
    n0 = min(len(q), 20000)
    t = time.perf_counter()
    _, used = env.check(q[:n0], 7, faithful=False, threads=all_host_threads())
    r0 = n0 / (time.perf_counter() - t)
    n = int(min(len(q), max(n0, r0 * seconds)))
    t = time.perf_counter()
    col, used = env.check(q[:n], 7, faithful=False, threads=all_host_threads())
    dt = time.perf_counter() - t
    return n / dt, used, n, col


def workload_name(name, cfg, n_vars, n_spheres, n_states, n_points, cells):
    """The same string in both arms: what is checked, not how."""
    return ("%s: %s %s (%d DoF, %d spheres), %d states per GPU per step, %d-point clutter cloud, %g x %g x %g m field @ %g m "
            "(%d x %d x %d voxels), checks world+self+intra" %
            (name, cfg["robot"], cfg["group"], n_vars, n_spheres, n_states, n_points, cfg["size"][0], cfg["size"][1],
             cfg["size"][2], cfg["resolution"], cells[0], cells[1], cells[2]))


def all_host_threads():
    """All the host threads the process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would otherwise
    silently turn the CPU arm into a single-thread baseline."""
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path, restated (oracle port; the reference cannot be
    compiled in this image), all host threads, each step a bounded sample of the C2 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("ORACLE_MARCH", "native")
    from moveit2_b200 import workloads
    import oracle
    oracle.build(force=True)
    cfg = workloads.CONFIGS[args.workload]
    cloud = workloads.make_cloud(args.workload)
    env = oracle_env(cfg, cloud)
    lo, hi = env.joint_bounds()
    n_full = args.states or cfg["states"]["n"]
    sample = min(n_full, 200_000)
    q = workloads.random_states(cfg["states"]["seed"], sample, lo, hi)
    host_threads = all_host_threads()
    for _ in range(max(args.warmup, 1)):
        env.check(q[:20000], 7, threads=host_threads)
    t = time.perf_counter()
    used = 1
    for _ in range(args.steps):
        _, used = env.check(q, 7, faithful=False, threads=host_threads)
    dt = time.perf_counter() - t
    rate = sample * args.steps / dt
    line = {"metric": METRIC, "value": rate, "unit": "states/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.workload, cfg, q.shape[1], env.n_spheres, n_full, len(cloud),
                                                 env.world.num_cells),
                       "sample": "each step checks the first %d states of the batch" % sample},
            "cpu_baseline": {"value": rate, "unit": "states/s", "cores": used, "kind": "port",
                             "sample": "%d states per step, lean flavour (centre voxel only), OpenMP" % sample},
            "e2e": {"value": rate, "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else any library prints to fd 1 (NCCL's
    version banner, build chatter) was diverted to stderr at start-up."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import moveit2_b200 as mb
    from moveit2_b200 import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    cfg = workloads.CONFIGS[args.workload]
    # C3 is a fixed 10 M-state batch cut over the ranks (whole 32-state tiles per rank); C2 is 1 M states per rank
    n = args.states or (cfg["states"]["n"] if args.workload == "C2" else cfg["states"]["n"] // world // 32 * 32)
    eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                           resolution=cfg["resolution"], max_distance=cfg["max_distance"], device=local)
    V, S = eng.n_group_vars, eng.n_spheres
    cloud = workloads.make_cloud(args.workload)                # replicated scene: every rank builds the same field
    eng.field_add_points(cloud)
    edt_ms_c2 = eng.stats()["edt_ms"]
    lo, hi = eng.group_bounds()
    q_host = workloads.random_states(cfg["states"]["seed"] + 1000 * rank, n, lo, hi)   # this rank's shard
    flags = mb.CHECK_ALL | (mb.CHECK_FP64 if args.fp64 else 0)

    # ---- device-resident step ---------------------------------------------------------------------------
    from moveit2_b200.sharding import ShardedStateValidity
    d_q = torch.from_numpy(q_host).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def check_local(q_local, bitmap):
        eng.check_batch_device(q_local.data_ptr(), q_local.shape[0], flags, bitmap.data_ptr(), stream.cuda_stream)

    # weak scaling: n states per rank; the global batch is world * n states, sharded contiguously (n is a multiple of
    # 32 or the single-GPU tail), bitmap slices all-gathered (NCCL over NVLink) at the end of every step
    sv = ShardedStateValidity(check_local, world * n, dev)
    assert sv.hi - sv.lo == n or world == 1
    d_bm = sv.local

    def step():
        sv.check(d_q)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                                   # L2 flush between timed steps (not timed)
        starts[i].record(stream)
        step()
        stops[i].record(stream)
    barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    st = eng.stats()
    valid_frac = float(np.unpackbits(d_bm.cpu().numpy(), bitorder="little")[:n].mean())

    # ---- end to end through the C-ABI with host buffers ----------------------------------------------------------
    q_pin = torch.from_numpy(q_host).pin_memory()
    bm_pin = torch.zeros((n + 7) // 8, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        eng.check_batch_ptr(q_pin.data_ptr(), n, flags, bm_pin.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.check_batch_ptr(q_pin.data_ptr(), n, flags, bm_pin.data_ptr())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    e2e_kernel_ms = eng.stats()["check_ms"]
    assert np.array_equal(bm_pin.numpy(), d_bm.cpu().numpy()[: (n + 7) // 8]), "host and device entry points disagree"

    # ---- max over ranks ----------------------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * n * args.steps / (dev_ms * 1e-3)
    e2e = world * n * args.steps / e2e_s
    peak, peak_src = load_peaks()
    # algorithmic bytes per state: 8 V (f64 joint values in) + 32 S F (one 32-byte sector per
    # sphere-centre voxel read) + 1/8 (validity bit out).  F = 1: world and self d^2 of a voxel are interleaved in ONE
    # compact field, so a single sector read serves both checks.
    b_state = 8 * V + 32 * S * 1 + 0.125
    per_gpu_rate = n * args.steps / (dev_ms * 1e-3)
    achieved = per_gpu_rate * b_state / 1e9
    gather_gbs = eng.gather_roofline(1 << 28, 3)
    gather_achieved = per_gpu_rate * 32 * S * 1 / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "states/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if args.fp64 else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, cfg, V, S, n, len(cloud), eng.num_cells),
                   "field_layout": "world+self uint%d d^2 interleaved, %.0f MB" %
                                   (8 * eng.info.field_bytes_per_voxel,
                                    2e-6 * eng.info.field_bytes_per_voxel * eng.num_cells[0] * eng.num_cells[1] * eng.num_cells[2]),
                   "l2": "flushed between timed steps (256 MiB write)", "fp32_pass_with_fp64_recheck": not args.fp64,
                   "valid_fraction": valid_frac, "rechecked_states_per_step": st["n_rechecked"],
                   "sharding": "states sharded contiguously (moveit2_b200/sharding.py), fields replicated, one NCCL "
                               "all_gather of the bitmap slices per step"
                   if world > 1 else "single GPU"},
        "e2e": {"value": e2e, "unit": "states/s", "h2d_bytes_per_step": n * V * 8, "d2h_bytes_per_step": (n + 7) // 8,
                "ms_per_step": 1e3 * e2e_s / args.steps, "kernel_ms_inside": e2e_kernel_ms},
        "gpu_launches": 2 * args.steps if not args.fp64 else args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch over 1 M states, from the committed
                     # ncu --set full capture (profiles/r1_fast_v13_summary.txt: 65.58 MB + 3.92 MB), scaled to n: the
                     # 16 MB field and the transform scratch are L2-resident, DRAM sees the states streaming in once
                     "traffic": 69.06 * n if args.workload == "C2" else None,
                     "traffic_source": "profiles/r1_fast_v14_summary.txt (ncu --set full, 1 launch)",
                     "peak_source": peak_src, "algorithmic_bytes_per_state": b_state,
                     "kernel": "fastCheckKernel<uint8_t,true> (+ the exact recheck launch recheckWarpKernel<uint8_t>)"},
        "gather_roofline": {"achieved": gather_achieved, "peak": gather_gbs, "unit": "GB/s (32 B sectors)",
                            "frac": gather_achieved / gather_gbs,
                            "how": "b200sv_gather_roofline: random 32 B sector reads over the same 16 MB field footprint"},
        "clocks": clocks,
        "edt_ms_c2_field": edt_ms_c2,
    }

    # ---- parity inside the run + CPU baseline (rank 0, bounded sample) -------------------------------------------------
    if not args.no_parity:
        os.environ.setdefault("ORACLE_MARCH", "native")
        import oracle
        oracle.build(force=True)
        env = oracle_env(cfg, cloud)
        g, o = eng.field_get_d2(), env.world.d2()
        line["parity"] = {"edt_voxels": int(g.size), "edt_voxels_differing_from_reference_propagation": int((g != o).sum()),
                          "edt_gpu_above_reference": int((g > o).sum())}
        edt_cpu = {"value": g.size / env.world_build_s, "unit": "voxels/s", "cores": 1, "kind": "port",
                   "sample": "the %s field: %d points -> %d voxels by addPointsToField / propagatePositive (%.2f s; the "
                             "propagation is serial by construction)" % (args.workload, len(cloud), g.size, env.world_build_s)}
        rate, used, ns, col = cpu_rate(env, q_host)
        line["cpu_baseline"] = {"value": rate, "unit": "states/s", "cores": used, "kind": "port",
                                "sample": "first %d states of the step's batch, oracle lean flavour, OpenMP" % ns}
        # the faithful flavour (what a CollisionEnvDistanceField call really does per state: GroupStateRepresentation copy,
        # every interior point re-posed, 7-voxel reads) on a short sample, for context
        nf = min(len(q_host), 20000)
        t0 = time.perf_counter()
        env.check(q_host[:nf], 7, faithful=True, threads=all_host_threads())
        line["cpu_baseline"]["faithful_flavour_states_per_s"] = nf / (time.perf_counter() - t0)
        # validity on identical fields: inject the reference's own propagation result
        eng.field_set_d2(o, mb.FIELD_WORLD)
        eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
        valid = eng.check_batch(q_host[:ns], flags)
        mism = valid == col.astype(bool)
        line["parity"].update({"validity_states_compared": int(ns), "validity_mismatches": int(mism.sum())})
        eng.field_add_points(np.zeros((0, 3)))   # rebuild own field
    else:
        line["cpu_baseline"] = None
        edt_cpu = None

    # ---- EDT voxels/s on C4 -----------------------------------------------------------------------------------------------
    if not args.no_edt:
        c4 = workloads.CONFIGS["C4"]
        e4 = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=c4["size"], origin=c4["origin"],
                                              resolution=c4["resolution"], max_distance=c4["max_distance"], device=local)
        pts = workloads.make_cloud("C4")
        d_pts = torch.from_numpy(pts).to(dev)
        nv = e4.num_cells[0] * e4.num_cells[1] * e4.num_cells[2]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for i in range(8):
            e4.field_reset()
            flush.fill_(i)
            ev0.record(stream)
            e4.field_add_points_device(d_pts.data_ptr(), len(pts), mb.FIELD_WORLD, stream.cuda_stream)
            ev1.record(stream)
            torch.cuda.synchronize()
            times.append(ev0.elapsed_time(ev1))
        ms = float(np.median(times[3:]))
        b_edt = 24 * len(pts) + 4 * nv
        # end to end through b200sv_field_add_points with HOST points: pinned (the contract's e2e), and pageable for context
        pts_pin = torch.from_numpy(pts).pin_memory()
        e2e_runs = []
        for _ in range(3):
            e4.field_reset()
            t0 = time.perf_counter()
            e4.field_add_points(pts_pin.numpy())
            e2e_runs.append(time.perf_counter() - t0)
        e2e_edt = float(np.median(e2e_runs))
        e4.field_reset()
        t0 = time.perf_counter()
        e4.field_add_points(pts)
        e2e_edt_pageable = time.perf_counter() - t0
        line["edt"] = {"metric": "edt_voxels_per_sec", "workload": "C4: 5M points -> 400x400x200 voxels @ 1 cm, max 25 cells",
                       "value": nv / (ms * 1e-3), "unit": "voxels/s", "ms": ms,
                       "e2e_ms_host_points": 1e3 * e2e_edt, "e2e_ms_pageable_host_points": 1e3 * e2e_edt_pageable,
                       "e2e_voxels_per_sec": nv / e2e_edt, "h2d_bytes": 24 * len(pts), "cpu_baseline": edt_cpu,
                       "roofline": {"bound": "hbm", "achieved": b_edt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": b_edt / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": b_edt}}
        e4.close()
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
