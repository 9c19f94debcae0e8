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

Besides the contract's keys the line carries: `parity` (10 M states vs the oracle on identical fields for the headline
config and for C3D, the count with each side's OWN fields, EDT voxel counts, and `reference_mode`: the fields rebuilt by the
device's reproduction of the reference's propagation -- voxels and verdicts differing from the oracle's, expected 0), `edt` (C4 rebuild, its own roofline),
`fk` (batched FK alone), `e2e_f32` / `e2e_edges` / `latency` (the other reference-facing entry points end to end),
`roofline_l1_gather` (the gather microbenchmark denominator) and, at N > 1, `c3_strong` (BASELINE.json configs[2]: the
10 M-state dual-arm batch sharded over the ranks, with the same batch on ONE rank timed in the same run).
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
    ap.add_argument("--no-c3", action="store_true", help="skip the C3D legs (parity at N = 1, strong scaling at N > 1)")
    ap.add_argument("--parity-states", type=int, default=10_000_000,
                    help="states compared with the oracle per parity leg (chunks of 1 M, bounded by a time budget)")
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


def make_config(workload):
    """`config` is byte-identical in both arms: what is checked and how the timed steps are separated, nothing run-dependent
    (those go under `run`)."""
    return {"workload": workload, "l2": "flushed between timed steps (256 MiB write); inputs resident in HBM for `value`"}


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
            "config": make_config(workload_name(args.workload, cfg, q.shape[1], env.n_spheres, n_full, len(cloud),
                                                env.world.num_cells)),
            "cpu_baseline": {"value": rate, "unit": "states/s", "cores": used, "kind": "port",
                             "sample": "each step checks the first %d states of the batch, lean flavour (centre voxel "
                                       "only), OpenMP" % sample},
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


def parity_leg(mb, workloads, name, eng, cloud, n_want, flags, budget_s=25.0):
    """Validity vs the oracle on IDENTICAL fields (the reference's own propagation result injected into the engine), in
    chunks of 1 M fresh states until n_want states or the time budget; then the same states with each side's OWN fields."""
    cfg = workloads.CONFIGS[name]
    env = oracle_env(cfg, cloud)
    own_w, own_s = eng.field_get_d2(mb.FIELD_WORLD), eng.field_get_d2(mb.FIELD_SELF)
    o_w, o_s = env.world.d2(), env.self_field.d2()
    out = {"edt_voxels": int(own_w.size), "edt_voxels_differing_from_reference_propagation": int((own_w != o_w).sum()),
           "edt_gpu_above_reference": int((own_w > o_w).sum()),
           "edt_occupancy_equal": bool(np.array_equal(own_w == 0, o_w == 0))}
    eng.field_set_d2(o_w, mb.FIELD_WORLD)
    eng.field_set_d2(o_s, mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    t0 = time.perf_counter()
    done = mism = rechecked = n_valid = 0
    cpu_s = 0.0
    chunks = []
    while done < n_want and (time.perf_counter() - t0 < budget_s or done == 0):
        m = min(1_000_000, n_want - done)
        q = workloads.random_states(900_000 + len(chunks), m, lo, hi)   # fresh states: not the timed batch
        t1 = time.perf_counter()
        col, used = env.check(q, 7, faithful=False, threads=all_host_threads())
        cpu_s += time.perf_counter() - t1
        valid = eng.check_batch(q, flags)
        rechecked += int(eng.stats()["n_rechecked"])
        mism += int((valid == col.astype(bool)).sum())
        n_valid += int(valid.sum())
        chunks.append((q, col) if len(chunks) < 2 else None)
        done += m
    out.update({"validity_states_compared": int(done), "validity_mismatches": int(mism),
                "states_decided_by_fp64_recheck": int(rechecked), "valid_fraction": n_valid / max(done, 1),
                "oracle_states_per_s": done / cpu_s, "oracle_threads": int(used)})
    # each side with its OWN field: the exact EDT is never above the reference's propagation, so the GPU can only report
    # MORE collisions; counted on the first chunks
    eng.field_set_d2(own_w, mb.FIELD_WORLD)
    eng.field_set_d2(own_s, mb.FIELD_SELF)
    own_diff = own_more = n_own = 0
    for item in chunks:
        if item is None:
            continue
        q, col = item
        valid = eng.check_batch(q, flags)
        d = valid == col.astype(bool)
        own_diff += int(d.sum())
        own_more += int((d & ~valid).sum())
        n_own += len(q)
    out.update({"validity_states_compared_with_own_fields": int(n_own),
                "validity_states_differing_with_own_fields": int(own_diff),
                "of_which_gpu_reports_collision": int(own_more)})
    # B200SV_PROPAGATION_REFERENCE: both fields rebuilt by the device's reproduction of the reference's bucket-queue
    # propagation from the same point arrays -- every voxel must equal the oracle's, and so must every state's verdict
    # with each side's own fields
    try:
        eng.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_WORLD)
        eng.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_SELF)
        ref_ms_all = []
        for _ in range(3):   # timing: the first build also allocates the mode's buffers
            eng.field_reset(mb.FIELD_WORLD)
            eng.field_add_points(cloud, mb.FIELD_WORLD)
            ref_ms_all.append(float(eng.stats()["edt_ms"]))
        ref_ms = min(ref_ms_all[1:])
        # parity: a FRESH field, as the oracle's is (reset() keeps what a build left in the queue, in the reference too, so
        # the rebuilds above are not what the oracle built once)
        eng.field_set_propagation(mb.PROPAGATION_EXACT, mb.FIELD_WORLD)
        eng.field_set_propagation(mb.PROPAGATION_REFERENCE, mb.FIELD_WORLD)
        eng.field_add_points(cloud, mb.FIELD_WORLD)
        r_w, r_s = eng.field_get_d2(mb.FIELD_WORLD), eng.field_get_d2(mb.FIELD_SELF)
        ref_diff = ref_n = 0
        for item in chunks:
            if item is None:
                continue
            q, col = item
            ref_diff += int((eng.check_batch(q, flags) == col.astype(bool)).sum())
            ref_n += len(q)
        st = eng.field_propagation_stats(mb.FIELD_WORLD)
        out["reference_mode"] = {
            "what": "b200sv_field_set_propagation(B200SV_PROPAGATION_REFERENCE): the reference's propagation reproduced on the "
                    "device, fields built from the same point arrays as the oracle's",
            "world_voxels_differing_from_reference_propagation": int((r_w != o_w).sum()),
            "self_voxels_differing_from_reference_propagation": int((r_s != o_s).sum()),
            "world_build_ms": ref_ms, "world_build_ms_first_call": ref_ms_all[0],
            "oracle_world_build_s": float(env.world_build_s),
            "rounds": st["rounds"], "hazard_rounds": st["hazard_rounds"], "queue_entries": st["queue_entries"],
            "validity_states_compared_with_own_fields": int(ref_n),
            "validity_states_differing_with_own_fields": int(ref_diff)}
    finally:  # back to the exact transform and this run's own fields
        eng.field_set_propagation(mb.PROPAGATION_EXACT, mb.FIELD_WORLD)
        eng.field_set_propagation(mb.PROPAGATION_EXACT, mb.FIELD_SELF)
        eng.field_set_d2(own_w, mb.FIELD_WORLD)
        eng.field_set_d2(own_s, mb.FIELD_SELF)
    return out, env


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
    from moveit2_b200.sharding import ShardedStateValidity, broadcast_field

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_engine(name):
        """Engine for a config; the world field is built ONCE (rank 0) and broadcast over NCCL to the other ranks."""
        cfg = workloads.CONFIGS[name]
        eng = mb.StateValidityEngine.for_robot(cfg["robot"], cfg["group"], size=cfg["size"], origin=cfg["origin"],
                                               resolution=cfg["resolution"], max_distance=cfg["max_distance"], device=local)
        cloud = workloads.make_cloud(name)
        bytes_bc = 0
        if rank == 0:
            eng.field_add_points(cloud)
        if world > 1:
            bytes_bc = broadcast_field(eng, mb.FIELD_WORLD, dev)
        return cfg, eng, cloud, bytes_bc

    def timed_steps(eng, d_q, n_global, flags, steps, warmup, overlap):
        """K steps of the sharded check; device time by CUDA events on the launching stream, L2 flushed between steps."""
        def check_local(q_local, bitmap):
            eng.check_batch_device(q_local.data_ptr(), q_local.shape[0], flags, bitmap.data_ptr(), stream.cuda_stream)
        sv = ShardedStateValidity(check_local, n_global, dev, overlap=overlap)
        for _ in range(max(warmup, 3)):
            flush.fill_(1)
            sv.check(d_q)
        sv.wait()
        barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.fill_(i & 0xFF)                               # L2 flush between timed steps (not timed)
            starts[i].record(stream)
            sv.check(d_q)
            if i == steps - 1:
                sv.wait()                                       # the last batch's gather is inside the timed region
            stops[i].record(stream)
        barrier()
        return sum(s.elapsed_time(e) for s, e in zip(starts, stops)), sv

    cfg, eng, cloud, field_bc_bytes = make_engine(args.workload)
    # C3* are a fixed 10 M-state batch cut over the ranks (whole 32-state tiles per rank); C2 is 1 M states per rank
    n = args.states or (cfg["states"]["n"] if args.workload == "C2" else cfg["states"]["n"] // world // 32 * 32)
    V, S = eng.n_group_vars, eng.n_spheres
    edt_ms_c2 = eng.stats()["edt_ms"]
    lo, hi = eng.group_bounds()
    q_host = workloads.random_states(cfg["states"]["seed"] + 1000 * rank, n, lo, hi)   # this rank's shard
    flags = mb.CHECK_ALL | (mb.CHECK_FP64 if args.fp64 else 0)

    # ---- device-resident steps -------------------------------------------------------------------------------------
    d_q = torch.from_numpy(q_host).to(dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, sv = timed_steps(eng, d_q, world * n, flags, args.steps, args.warmup, overlap=True)
    assert sv.hi - sv.lo == n or world == 1
    d_bm = sv.local
    st = eng.stats()
    valid_frac = float(np.unpackbits(d_bm.cpu().numpy(), bitorder="little")[:n].mean())

    # ---- end to end through the C-ABI with host buffers ----------------------------------------------------------------
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
    # float32 states through b200sv_check_batch_f32 (not the reference's layout; half the H2D bytes)
    qf_pin = torch.from_numpy(q_host.astype(np.float32)).pin_memory()
    bmf_pin = torch.zeros((n + 7) // 8, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        eng.check_batch_f32_ptr(qf_pin.data_ptr(), n, flags, bmf_pin.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.check_batch_f32_ptr(qf_pin.data_ptr(), n, flags, bmf_pin.data_ptr())
    torch.cuda.synchronize()
    e2e_f32_s = time.perf_counter() - t0

    # ---- max over ranks ----------------------------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_s, e2e_f32_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s, e2e_f32_s = float(t[0]), float(t[1]), float(t[2])

    # ---- N > 1: BASELINE.json configs[2], the 10 M-state PR2-class dual-arm batch (C3D), strong scaling ----------------------
    c3 = None
    if world > 1 and not args.no_c3:
        del d_q
        cfg3, eng3, cloud3, _ = make_engine("C3D")
        n3_total = cfg3["states"]["n"] // (32 * world) * (32 * world)
        n3 = n3_total // world
        lo3, hi3 = eng3.group_bounds()
        k3 = max(3, min(args.steps, 10))
        d_q3 = torch.from_numpy(workloads.random_states(cfg3["states"]["seed"] + 1000 * rank, n3, lo3, hi3)).to(dev)
        ms3, _ = timed_steps(eng3, d_q3, n3_total, mb.CHECK_ALL, k3, 3, overlap=True)
        t3 = torch.tensor([ms3], dtype=torch.float64, device=dev)
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        ms3 = float(t3[0]) / k3
        del d_q3
        c3 = {"workload": "C3D: %s %s (%d DoF, %d spheres), ONE batch of %d states sharded over %d GPUs, %d-point clutter "
                          "cloud, checks world+self+intra" % (cfg3["robot"], cfg3["group"], eng3.n_group_vars, eng3.n_spheres,
                                                              n3_total, world, len(cloud3)),
              "scaling": "strong", "ms_per_batch": ms3, "value": n3_total / (ms3 * 1e-3), "unit": "states/s", "steps": k3}
        if rank == 0:   # the same batch on ONE GPU, same run, same box: the denominator of the strong-scaling efficiency
            world_keep = world
            d_q1 = torch.from_numpy(workloads.random_states(cfg3["states"]["seed"], n3_total, lo3, hi3)).to(dev)
            bm1 = torch.zeros(n3_total // 8, dtype=torch.uint8, device=dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * k3)]
            for i in range(2 + k3):
                flush.fill_(i)
                if i >= 2:
                    ev[2 * (i - 2)].record(stream)
                eng3.check_batch_device(d_q1.data_ptr(), n3_total, mb.CHECK_ALL, bm1.data_ptr(), stream.cuda_stream)
                if i >= 2:
                    ev[2 * (i - 2) + 1].record(stream)
            torch.cuda.synchronize()
            ms1 = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(k3)) / k3
            c3.update({"ms_per_batch_one_gpu": ms1, "value_one_gpu": n3_total / (ms1 * 1e-3),
                       "speedup_vs_one_gpu": ms1 / ms3, "efficiency_vs_one_gpu": ms1 / ms3 / world_keep})
            del d_q1
        barrier()
        eng3.close()
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
        "config": make_config(workload_name(args.workload, cfg, V, S, n, len(cloud), eng.num_cells)),
        "run": {"field_layout": "world+self uint%d d^2 interleaved, %.0f MB" %
                                (8 * eng.info.field_bytes_per_voxel,
                                 2e-6 * eng.info.field_bytes_per_voxel * eng.num_cells[0] * eng.num_cells[1] * eng.num_cells[2]),
                "fp32_pass_with_fp64_recheck": not args.fp64, "valid_fraction": valid_frac,
                "rechecked_states_per_step": st["n_rechecked"],
                "sharding": ("states sharded contiguously (moveit2_b200/sharding.py); the world field is built on rank 0 and "
                             "broadcast by NCCL (%d bytes); one NCCL all_gather of the bitmap slices per step, issued on a "
                             "side stream so it overlaps the next step's kernels (the last one is inside the timed region)"
                             % field_bc_bytes) if world > 1 else "single GPU"},
        "e2e": {"value": e2e, "unit": "states/s", "h2d_bytes_per_step": n * V * 8, "d2h_bytes_per_step": (n + 7) // 8,
                "ms_per_step": 1e3 * e2e_s / args.steps, "kernel_ms_inside": e2e_kernel_ms},
        "e2e_f32": {"value": world * n * args.steps / e2e_f32_s, "unit": "states/s", "h2d_bytes_per_step": n * V * 4,
                    "d2h_bytes_per_step": (n + 7) // 8, "ms_per_step": 1e3 * e2e_f32_s / args.steps,
                    "entry": "b200sv_check_batch_f32 (float32 states, widened on the device; not the reference's layout)"},
        "gpu_launches": 2 * args.steps if not args.fp64 else args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch over 1 M states, from the committed
                     # ncu --set full capture, scaled to n: the 16 MB field and the transform scratch are L2-resident,
                     # DRAM sees the states streaming in once
                     "traffic": 70.21 * n if args.workload == "C2" else None,
                     "traffic_source": "profiles/r2_fast_v15_summary.txt (ncu --set full, 1 launch: 65.58 MB read + 4.63 MB written)",
                     "peak_source": peak_src, "algorithmic_bytes_per_state": b_state,
                     "binding": "NOT DRAM: the field is L2-resident (ncu: dram 2 %, L2 hit 93 %); the kernel is bound by "
                                "instruction issue (67 %) and the L1TEX data pipe (66 %) -- frac is algorithmic sector "
                                "bytes over the HBM copy peak, the north star's figure is roofline_l1_gather",
                     "kernel": "fastCheckKernel<uint8_t,true> (+ the exact recheck launch recheckWarpKernel<uint8_t>)"},
        "roofline_l1_gather": {"bound": "l1tex-gather", "achieved": gather_achieved, "peak": gather_gbs,
                               "unit": "GB/s (32 B sectors)", "frac": gather_achieved / gather_gbs,
                               "how": "b200sv_gather_roofline: random 32 B sector reads over the same field footprint "
                                      "(xorshift + multiply-shift addressing, ncu: profiles/r2_gather_summary.txt)"},
        "clocks": clocks,
        "edt_ms_c2_field": edt_ms_c2,
    }
    line["gather_roofline"] = line["roofline_l1_gather"]   # round-1 name of the same object
    if c3 is not None:
        line["c3_strong"] = c3

    # ---- the other reference-facing entry points, end to end (rank 0) ------------------------------------------------
    # batched FK alone (b200sv_fk_batch_device): HBM-write bound, 128 B per link per state
    nf = min(n, 500_000)
    d_qf = torch.from_numpy(q_host[:nf]).to(dev)
    d_T = torch.empty((nf, eng.n_links, 16), dtype=torch.float64, device=dev)
    fk = {}
    for prec in (32, 64):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for i in range(6):
            flush.fill_(i)
            ev0.record(stream)
            eng.fk_batch_device(d_qf.data_ptr(), nf, prec, d_T.data_ptr(), stream.cuda_stream)
            ev1.record(stream)
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        ms = float(np.median(ts[2:]))
        b_fk = 8 * V + 128 * eng.n_links
        fk["fp%d" % prec] = {"states_per_s": nf / (ms * 1e-3), "ms": ms,
                             "roofline": {"bound": "hbm", "achieved": nf * b_fk / (ms * 1e-3) / 1e9, "peak": peak,
                                          "unit": "GB/s", "frac": nf * b_fk / (ms * 1e-3) / 1e9 / peak,
                                          "algorithmic_bytes_per_state": b_fk}}
    fk["workload"] = "%d states, all %d model links as 4x4 f64 out (RobotState::updateLinkTransforms)" % (nf, eng.n_links)
    line["fk"] = fk
    del d_T, d_qf
    # batched edge validation: 2 states in, ~34 checked on the device
    ne = 100_000
    q_pin_np = q_pin.numpy()   # the edges' end states come from pinned host memory, like the e2e leg's states
    ea, eb = q_pin_np[:ne], q_pin_np[ne:2 * ne]
    seg = 0.01 * float(np.sum(hi - lo))
    for _ in range(2):
        eng.check_motions(ea, eb, seg)
    t0 = time.perf_counter()
    for _ in range(3):
        _, _, n_exp = eng.check_motions(ea, eb, seg)
    dt = (time.perf_counter() - t0) / 3
    line["e2e_edges"] = {"edges_per_s": ne / dt, "states_per_s": n_exp / dt, "edges": ne, "states_checked": int(n_exp),
                         "h2d_bytes_per_call": 2 * ne * V * 8, "d2h_bytes_per_call": ne // 8 + 4 * ne, "ms_per_call": 1e3 * dt,
                         "entry": "b200sv_check_motions (pinned host edges in, validity bitmap + first invalid index out)"}
    # single-state latency through the host entry point (OMPL's per-state isValid through the plugin)
    lat = {}
    for nb in (1, 32):
        for _ in range(50):
            eng.check_batch_bitmap(q_host[:nb], flags)
        ts = []
        bmp = np.zeros(8, np.uint8)
        qq = np.ascontiguousarray(q_host[:nb])
        for _ in range(500):
            t1 = time.perf_counter()
            eng.check_batch_ptr(qq.ctypes.data, nb, flags, bmp.ctypes.data)
            ts.append(time.perf_counter() - t1)
        lat["n=%d" % nb] = {"median_us": 1e6 * float(np.median(ts)), "p90_us": 1e6 * float(np.percentile(ts, 90))}
    # a new state every call, back to back (a planner's per-state isValid): the resident kernel answers without a launch
    ts = []
    for i in range(1000):
        qq = q_host[i:i + 1]
        t1 = time.perf_counter()
        eng.check_batch_ptr(qq.ctypes.data, 1, flags, bmp.ctypes.data)
        ts.append(time.perf_counter() - t1)
    lat["n=1, a new state every call"] = {"median_us": 1e6 * float(np.median(ts)), "p90_us": 1e6 * float(np.percentile(ts, 90))}
    # ONE edge per b200sv_check_motions call (an OMPL MotionValidator::checkMotion): one launch, the batch sized on the host
    ts, n_exp1 = [], 0
    for i in range(300):
        t1 = time.perf_counter()
        _, _, n1 = eng.check_motions(q_host[i:i + 1], q_host[1000 + i:1001 + i], seg)
        ts.append(time.perf_counter() - t1)
        n_exp1 += n1
    lat["one edge per check_motions call"] = {"median_us": 1e6 * float(np.median(ts[20:])), "p90_us": 1e6 * float(np.percentile(ts[20:], 90)),
                                              "states_per_edge": n_exp1 / 300.0}
    lat["how"] = ("wall clock around b200sv_check_batch / b200sv_check_motions from Python ctypes and numpy (adds 1-3 us; "
                  "tools/c_latency.c and tools/c_edge_latency.c measure from C), pageable host buffers")
    line["latency"] = lat

    # ---- parity inside the run + CPU baseline (rank 0, bounded) ----------------------------------------------------------
    if not args.no_parity:
        os.environ.setdefault("ORACLE_MARCH", "native")
        import oracle
        oracle.build(force=True)
        par, env = parity_leg(mb, workloads, args.workload, eng, cloud, args.parity_states, flags)
        line["parity"] = par
        edt_cpu = {"value": par["edt_voxels"] / env.world_build_s, "unit": "voxels/s", "cores": 1, "kind": "port",
                   "sample": "the %s field: %d points -> %d voxels by addPointsToField / propagatePositive (%.2f s; the "
                             "propagation is serial by construction)" % (args.workload, len(cloud), par["edt_voxels"],
                                                                         env.world_build_s)}
        rate, used, ns, col = cpu_rate(env, q_host)
        line["cpu_baseline"] = {"value": rate, "unit": "states/s", "cores": used, "kind": "port",
                                "sample": "first %d states of the step's batch, oracle lean flavour, OpenMP" % ns}
        # the faithful flavour (what a CollisionEnvDistanceField call really does per state: GroupStateRepresentation copy,
        # every interior point re-posed, 7-voxel reads) on a short sample, for context
        nfa = min(len(q_host), 20000)
        t0 = time.perf_counter()
        env.check(q_host[:nfa], 7, faithful=True, threads=all_host_threads())
        line["cpu_baseline"]["faithful_flavour_states_per_s"] = nfa / (time.perf_counter() - t0)
        del env
        if args.workload != "C3D" and not args.no_c3:   # the PR2-class robot of configs[2], same bar
            cfg3 = workloads.CONFIGS["C3D"]
            eng3 = mb.StateValidityEngine.for_robot(cfg3["robot"], cfg3["group"], size=cfg3["size"], origin=cfg3["origin"],
                                                    resolution=cfg3["resolution"], max_distance=cfg3["max_distance"],
                                                    device=local)
            cloud3 = workloads.make_cloud("C3D")
            eng3.field_add_points(cloud3)
            par3, _ = parity_leg(mb, workloads, "C3D", eng3, cloud3, args.parity_states, mb.CHECK_ALL)
            par3["workload"] = "C3D: dualarm_dense arms (%d DoF, %d spheres), the C3 scene" % (eng3.n_group_vars, eng3.n_spheres)
            line["parity_c3d"] = par3
            eng3.close()
    else:
        line["cpu_baseline"] = None
        edt_cpu = None

    # ---- the cloud SURVEY.md 8d wrote down literally (half uniform in the volume): a reported side case ----------------------
    if not args.no_parity and args.workload == "C2":
        cfg_s = workloads.CONFIGS["C2S"]
        eng_s = mb.StateValidityEngine.for_robot(cfg_s["robot"], cfg_s["group"], size=cfg_s["size"], origin=cfg_s["origin"],
                                                 resolution=cfg_s["resolution"], max_distance=cfg_s["max_distance"], device=local)
        cloud_s = workloads.make_cloud("C2S")
        eng_s.field_add_points(cloud_s)
        ns = 1_000_000
        qs_host = workloads.random_states(cfg_s["states"]["seed"], ns, lo, hi)
        d_qs = torch.from_numpy(qs_host).to(dev)
        bm_s = torch.zeros(4 * ((ns + 31) // 32), dtype=torch.uint8, device=dev)
        ts = []
        for i in range(8):
            flush.fill_(i)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            eng_s.check_batch_device(d_qs.data_ptr(), ns, flags, bm_s.data_ptr(), stream.cuda_stream)
            ev1.record(stream)
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        ms_s = float(np.median(ts[3:]))
        env_s = oracle_env(cfg_s, cloud_s)
        eng_s.field_set_d2(env_s.world.d2(), mb.FIELD_WORLD)
        eng_s.field_set_d2(env_s.self_field.d2(), mb.FIELD_SELF)
        nsp = 200_000
        col_s, _ = env_s.check(qs_host[:nsp], 7, faithful=False, threads=all_host_threads())
        val_s = eng_s.check_batch(qs_host[:nsp], flags)
        line["survey_c2"] = {"workload": "C2 with SURVEY.md 8d's literal cloud: 250 000 points uniform in the 8 m^3 volume + 250 000 "
                                         "on 20 random boxes, no keep-out",
                             "value": ns / (ms_s * 1e-3), "unit": "states/s", "ms_per_step": ms_s,
                             "valid_fraction": float(val_s.mean()), "validity_states_compared": nsp,
                             "validity_mismatches": int((val_s == col_s.astype(bool)).sum()),
                             "note": "at this density nearly every random arm state collides and stops gathering early: "
                                     "a degenerate workload for a validity kernel, which is why the headline uses surface "
                                     "clutter (moveit2_b200/workloads.py)"}
        del d_qs, env_s
        eng_s.close()

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
        # a checksum of the field just built, and of the same points inserted in another order: the result is a function
        # of the occupancy SET (tests/test_gpu_round2.py compares all 32 M voxels with the exact transform)
        g4 = e4.field_get_d2()
        checksum = int(g4.astype(np.int64).sum())
        occupied = int((g4 == 0).sum())
        del g4
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
                       "d2_checksum": checksum, "occupied_voxels": occupied,
                       "roofline": {"bound": "hbm", "achieved": b_edt / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": b_edt / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": b_edt}}
        e4.close()
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
