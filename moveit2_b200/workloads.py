"""Synthetic, seeded inputs for the BASELINE.json configs.  numpy only.

Deviation from SURVEY.md 8d, on purpose: the survey proposed obstacle points *uniform in the field volume*
(2 000 points in 1 m^3 for C1, 250 000 in 8 m^3 for C2).  At those densities every random arm state collides
(mean nearest-obstacle distance 4 cm / 2 cm vs. 6 cm link spheres) and a validity kernel with early exit would be
benchmarked on a degenerate all-colliding workload.  Sensor clouds are surfaces, not volumes, so the clouds
here are points on the surfaces of random boxes ("clutter", after the reference's own speed-test recipe:
moveit_ros/planning/planning_components_tools/src/compare_collision_speed_checking_fcl_bullet.cpp:83-160,
boxes rejected when they intrude on the robot's base column) plus a floor patch.  Roughly half the states
collide; the exact fraction is printed by bench.py.
"""
import numpy as np


def box_surface_points(rng, center, half, rot_z, n):
    """n points uniform on the surface of a box (area-weighted faces), rotated about z, translated."""
    hx, hy, hz = half
    areas = np.array([hy * hz, hy * hz, hx * hz, hx * hz, hx * hy, hx * hy])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    u = rng.uniform(-1.0, 1.0, size=(n, 3)) * np.array(half)
    ax = face // 2
    sign = np.where(face % 2 == 0, -1.0, 1.0)
    u[np.arange(n), ax] = sign * np.array(half)[ax]
    c, s = np.cos(rot_z), np.sin(rot_z)
    R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    return u @ R.T + np.array(center)


def clutter_cloud(seed, n_points, n_boxes, lo, hi, keepout_radius=0.28, edge=(0.05, 0.2), floor=True):
    """Point cloud on `n_boxes` random boxes inside [lo, hi], none closer than keepout_radius to the z axis."""
    rng = np.random.default_rng(seed)
    lo, hi = np.array(lo, float), np.array(hi, float)
    boxes = []
    while len(boxes) < n_boxes:
        c = rng.uniform(lo, hi)
        half = rng.uniform(edge[0], edge[1], size=3) / 2.0
        if np.hypot(c[0], c[1]) - np.linalg.norm(half[:2]) < keepout_radius:
            continue
        boxes.append((c, half, rng.uniform(0, np.pi)))
    n_floor = n_points // 10 if floor else 0
    per = (n_points - n_floor) // n_boxes
    pts = [box_surface_points(rng, c, h, a, per) for (c, h, a) in boxes]
    rest = n_points - n_floor - per * n_boxes
    if rest:
        pts.append(box_surface_points(rng, boxes[0][0], boxes[0][1], boxes[0][2], rest))
    if n_floor:
        # a floor ring well below the arm's shoulder: z = lo_z + one cell, outside the keep-out disc
        r = np.sqrt(rng.uniform((keepout_radius + 0.25) ** 2, min(hi[0], hi[1]) ** 2, size=n_floor))
        t = rng.uniform(0, 2 * np.pi, size=n_floor)
        pts.append(np.stack([r * np.cos(t), r * np.sin(t), np.full(n_floor, lo[2] + 0.015)], axis=1))
    return np.ascontiguousarray(np.concatenate(pts, axis=0), dtype=np.float64)


def random_states(seed, n, lo, hi):
    """Joint vectors uniform inside the bounds MoveIt samples (RevoluteJointModel::getVariableRandomPositions)."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.uniform(lo, hi, size=(n, len(lo))), dtype=np.float64)


CONFIGS = {
    # name: field size (m), field origin (centre, CollisionEnvDistanceField convention corner = origin - size/2),
    #       resolution, max propagation distance, cloud (seed, n_points, n_boxes, lo, hi), states (seed, n)
    "C1": dict(robot="panda", group="panda_arm", size=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.5), resolution=0.02,
               max_distance=0.25, cloud=dict(seed=1, n_points=2000, n_boxes=6, lo=(-0.45, -0.45, 0.05),
                                             hi=(0.45, 0.45, 0.95), keepout_radius=0.2),
               states=dict(seed=0, n=10_000)),
    "C2": dict(robot="panda", group="panda_arm", size=(2.0, 2.0, 2.0), origin=(0.0, 0.0, 0.5), resolution=0.01,
               max_distance=0.25, cloud=dict(seed=3, n_points=500_000, n_boxes=24, lo=(-0.9, -0.9, -0.3),
                                             hi=(0.9, 0.9, 1.3), keepout_radius=0.3),
               states=dict(seed=2, n=1_000_000)),
    "C3": dict(robot="dualarm", group="arms", size=(3.0, 3.0, 2.5), origin=(0.0, 0.0, 1.0), resolution=0.02,
               max_distance=0.25, cloud=dict(seed=6, n_points=200_000, n_boxes=30, lo=(-1.3, -1.3, 0.1),
                                             hi=(1.3, 1.3, 2.0), keepout_radius=0.75),
               states=dict(seed=4, n=10_000_000)),
    # C3D: the C3 scene with the dual arm at the sphere count of a PR2-class decomposition (280 spheres on 28 links,
    # ~1 700 cluster pairs: tools/gen_robots.py gen_dualarm_dense)
    "C3D": dict(robot="dualarm_dense", group="arms", size=(3.0, 3.0, 2.5), origin=(0.0, 0.0, 1.0), resolution=0.02,
                max_distance=0.25, cloud=dict(seed=6, n_points=200_000, n_boxes=30, lo=(-1.3, -1.3, 0.1),
                                              hi=(1.3, 1.3, 2.0), keepout_radius=0.75),
                states=dict(seed=4, n=10_000_000)),
    # C3W: the C3 scene, the dual arm's whole_body group (planar base + torso + both arms: 18 variables, 93 spheres)
    "C3W": dict(robot="dualarm", group="whole_body", size=(3.0, 3.0, 2.5), origin=(0.0, 0.0, 1.0), resolution=0.02,
                max_distance=0.25, cloud=dict(seed=6, n_points=200_000, n_boxes=30, lo=(-1.3, -1.3, 0.1),
                                              hi=(1.3, 1.3, 2.0), keepout_radius=0.75),
                states=dict(seed=4, n=10_000_000)),
    # C2S: C2 with the cloud SURVEY.md 8d wrote down literally -- half the points uniform in the field VOLUME, half on the
    # surfaces of 20 random boxes (no keep-out around the robot, no floor).  At that density nearly every random arm state
    # collides (see the module docstring), so it is a reported side case (bench key `survey_c2`), not the headline.
    "C2S": dict(robot="panda", group="panda_arm", size=(2.0, 2.0, 2.0), origin=(0.0, 0.0, 0.5), resolution=0.01,
                max_distance=0.25, cloud=dict(seed=3, n_points=500_000, n_boxes=20, lo=(-0.9, -0.9, -0.4),
                                              hi=(0.9, 0.9, 1.4), keepout_radius=0.0, uniform_fraction=0.5),
                states=dict(seed=2, n=1_000_000)),
    # C4: EDT rebuild only -- 5 M points uniform in a 4 x 4 x 2 m box, 400 x 400 x 200 voxels at 1 cm
    "C4": dict(size=(4.0, 4.0, 2.0), origin=(0.0, 0.0, 1.0), resolution=0.01, max_distance=0.25,
               cloud=dict(seed=5, n_points=5_000_000)),
}


def uniform_cloud(seed, n_points, size, origin):
    rng = np.random.default_rng(seed)
    lo = np.array(origin) - 0.5 * np.array(size)
    hi = np.array(origin) + 0.5 * np.array(size)
    return np.ascontiguousarray(rng.uniform(lo, hi, size=(n_points, 3)), dtype=np.float64)


def make_cloud(name):
    cfg = CONFIGS[name]
    c = cfg["cloud"]
    if "n_boxes" not in c:
        return uniform_cloud(c["seed"], c["n_points"], cfg["size"], cfg["origin"])
    if c.get("uniform_fraction"):
        n_uni = int(c["n_points"] * c["uniform_fraction"])
        uni = uniform_cloud(c["seed"], n_uni, cfg["size"], cfg["origin"])
        box = clutter_cloud(c["seed"] + 1, c["n_points"] - n_uni, c["n_boxes"], c["lo"], c["hi"], c.get("keepout_radius", 0.0),
                            floor=False)
        return np.ascontiguousarray(np.concatenate([uni, box], axis=0))
    return clutter_cloud(c["seed"], c["n_points"], c["n_boxes"], c["lo"], c["hi"], c.get("keepout_radius", 0.28))
