"""Shared helpers of the GPU parity tests: build the product engine and the oracle env on identical inputs."""
import os

import numpy as np

import moveit2_b200 as mb
from moveit2_b200 import workloads
from oracle.collision_model import OracleEnv
from oracle.robot_model import RobotModel

RES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "moveit2_b200", "resources")
ROBOT_FILES = {"panda": ("panda_spheres.urdf", "panda.srdf"), "dualarm": ("dualarm_spheres.urdf", "dualarm.srdf"),
             "dualarm_dense": ("dualarm_dense_spheres.urdf", "dualarm.srdf"),
             "panda_primitives": ("panda_primitives.urdf", "panda.srdf")}


def read(*p):
    with open(os.path.join(*p)) as f:
        return f.read()


def make_pair(robot, group, size, origin, resolution, max_distance, tolerance=0.0, device=0, signed=False):
    u, s = ROBOT_FILES[robot]
    eng = mb.StateValidityEngine.from_urdf(read(RES, u), read(RES, s), group, size, origin, resolution, max_distance,
                                           tolerance, device=device, signed=signed)
    env = OracleEnv(RobotModel(read(RES, u), read(RES, s)), group, size, origin, resolution, max_distance, tolerance,
                    use_signed_distance_field=signed)
    return eng, env


def config_pair(name, n_states=None, device=0):
    cfg = workloads.CONFIGS[name]
    eng, env = make_pair(cfg["robot"], cfg["group"], cfg["size"], cfg["origin"], cfg["resolution"],
                         cfg["max_distance"], device=device)
    cloud = workloads.make_cloud(name)
    lo, hi = eng.group_bounds()
    n = cfg["states"]["n"] if n_states is None else n_states
    q = workloads.random_states(cfg["states"]["seed"], n, lo, hi)
    return eng, env, cloud, q


def unpack(bitmap, n):
    return np.unpackbits(np.asarray(bitmap, np.uint8), bitorder="little")[:n].astype(bool)
