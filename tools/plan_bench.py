#!/usr/bin/env python3
"""Config C5 (BASELINE.json): RRTConnect for the Panda arm in the cluttered C2 scene, planning-time distribution over
seeded queries, edges validated in batches through b200sv_check_motions.  The reference arm runs the SAME planner with
the oracle's CPU restatement of the validity path as the validator (all host threads).  Not a roofline item.

  python tools/plan_bench.py [--queries 50] [--batch 32] [--reference-queries 10] > gpurun_out/plan_c5.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleMotionValidator:
    """Same interface as moveit2_b200.ompl_interface.MotionValidator, decisions by the CPU oracle (reference arm)."""

    def __init__(self, env, space):
        self.env, self.space = env, space
        self.n_edges = self.n_states = self.n_calls = 0

    def checkMotions(self, s1, s2):
        from oracle import motion
        L = self.space.getLongestValidSegmentLength()
        valid, first, n = motion.check_motions(self.env, np.atleast_2d(s1), np.atleast_2d(s2), L, self.space.continuous,
                                               self.space.factor)
        self.n_edges += len(valid)
        self.n_states += n
        self.n_calls += 1
        nd = np.maximum(np.ceil(self.space.distance(np.atleast_2d(s1), np.atleast_2d(s2)) / L), 1.0)
        return valid, np.where(valid, 1.0, (np.maximum(first, 1) - 1) / nd)


def box_clutter(rng, n_boxes, home_centres, home_radii):
    """The reference's own speed-test recipe (compare_collision_speed_checking_fcl_bullet.cpp:83-160): boxes with edges
    U(0.05, 0.2) m at x, y in U(-1, 1), z in U(0, 1), random orientation, rejected if they touch the robot at its home pose
    (here: any home-pose collision sphere within 3 cm of the box)."""
    boxes = []
    while len(boxes) < n_boxes:
        dims = rng.uniform(0.05, 0.2, 3)
        xyz = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0, 1)])
        qv = rng.normal(size=4)
        w, x, y, z = qv / np.linalg.norm(qv)
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        local = (home_centres - xyz) @ R  # centres in the box frame
        gap = np.linalg.norm(local - np.clip(local, -dims / 2, dims / 2), axis=1) - home_radii
        if gap.min() < 0.03:
            continue
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, xyz
        boxes.append((dims, T))
    return boxes


def eng_resolution(eng):
    from moveit2_b200 import workloads
    return workloads.CONFIGS["C2"]["resolution"]


def summarise(times, ok):
    t = np.asarray(times)
    return {"queries": len(t), "solved": int(np.sum(ok)), "median_ms": float(np.median(t) * 1e3),
            "mean_ms": float(np.mean(t) * 1e3), "p90_ms": float(np.percentile(t, 90) * 1e3), "max_ms": float(np.max(t) * 1e3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=50)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reference-queries", type=int, default=10)
    ap.add_argument("--fraction", type=float, default=0.01, help="longest_valid_segment_fraction (OMPL default 0.01)")
    ap.add_argument("--scene", default="cloud", choices=["cloud", "boxes"],
                    help="cloud: the C2 clutter cloud; boxes: the reference's speed-test recipe, ingested as shapes")
    ap.add_argument("--boxes", type=int, default=100)
    args = ap.parse_args()
    os.environ.setdefault("ORACLE_MARCH", "native")
    import oracle
    oracle.build()
    import moveit2_b200 as mb
    from moveit2_b200.ompl_interface import ModelBasedStateSpace, MotionValidator, StateValidityChecker
    from moveit2_b200.planning import rrt_connect
    from util import config_pair
    eng, env, cloud, _ = config_pair("C2", n_states=1)
    if args.scene == "cloud":
        eng.field_add_points(cloud)
        env.world.add_points(cloud)
        scene = "C2 scene (500k-point clutter cloud"
    else:
        from oracle.collision_model import find_internal_points_body
        home = np.array([0.0, -0.785, 0.0, -2.356, 0.0, 1.571, 0.785])  # panda "ready" pose
        radii = env.arr["sphere_radius"][:env.n_spheres]
        for dims, T in box_clutter(np.random.default_rng(51), args.boxes, env.sphere_centers(home).reshape(-1, 3), radii):
            # the way CollisionEnvDistanceField ingests world objects: body-frame lattice, then the pose
            eng.field_add_shape(mb.SHAPE_BOX, dims, T, lattice_frame=mb.LATTICE_BODY)
            env.world.add_points(find_internal_points_body(4, dims, T, eng_resolution(eng), "body"))
        assert np.array_equal(eng.field_get_d2() == 0, env.world.d2() == 0), "device and restated box decomposition differ"
        scene = "%d random boxes (reference speed-test recipe, decomposed on the device" % args.boxes
    # both arms decide on identical fields: the reference's own propagation result
    eng.field_set_d2(env.world.d2(), mb.FIELD_WORLD)
    eng.field_set_d2(env.self_field.d2(), mb.FIELD_SELF)
    lo, hi = eng.group_bounds()
    space = ModelBasedStateSpace(lo, hi, longest_valid_segment_fraction=args.fraction)
    svc = StateValidityChecker(eng, space)
    rng = np.random.default_rng(50)
    cand = space.sampleUniform(rng, 16 * args.queries + 64)
    cand = cand[svc.isValidBatch(cand)]
    queries = [(cand[2 * i], cand[2 * i + 1]) for i in range(args.queries)]
    out = {"workload": "C5: Panda panda_arm, %s, 2 m^3 field @ 1 cm), RRTConnect, "
                       "batch %d edges per validator call, longest_valid_segment_fraction %g" % (scene, args.batch, args.fraction)}
    for name, validator, nq in (("b200", MotionValidator(eng, space), args.queries),
                                ("reference_cpu_port", OracleMotionValidator(env, space), args.reference_queries)):
        times, ok, paths = [], [], []
        rrt_connect(space, validator, queries[0][0], queries[0][1], np.random.default_rng(0), batch=args.batch)  # warm-up
        validator.n_edges = validator.n_states = validator.n_calls = 0
        for i in range(nq):
            path, st = rrt_connect(space, validator, queries[i][0], queries[i][1], np.random.default_rng(1000 + i),
                                   batch=args.batch)
            times.append(st["time_s"])
            ok.append(path is not None)
            paths.append(path)
        s = summarise(times, ok)
        s.update({"edges_checked": validator.n_edges, "states_checked": validator.n_states,
                  "validator_calls": validator.n_calls})
        out[name] = s
        if name == "b200":
            # every edge of every returned path re-validated by the oracle: the plans are collision-free for the reference
            from oracle import motion
            bad = 0
            for p in paths:
                if p is None:
                    continue
                a, b = np.stack(p[:-1]), np.stack(p[1:])
                v, _, _ = motion.check_motions(env, a, b, space.getLongestValidSegmentLength())
                bad += int((~v).sum())
            s["path_edges_rejected_by_oracle"] = bad
    print(json.dumps(out))


if __name__ == "__main__":
    main()
