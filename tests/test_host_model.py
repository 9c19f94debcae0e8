"""CPU tests of the product's HOST side (no GPU compute): the C-ABI library loads and exports what the header
declares, refuses to compute without a device, and its C++ URDF/SRDF model builder agrees with the oracle's
independent Python restatement of RobotModel / JointModelGroup / DistanceFieldCacheEntry."""
import os
import re

import numpy as np
import pytest

import moveit2_b200 as mb
from moveit2_b200 import _lib
from oracle.collision_model import OracleEnv
from oracle.robot_model import RobotModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(*p):
    with open(os.path.join(*p)) as f:
        return f.read()


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    header = read(ROOT, "include", "b200sv.h")
    declared = sorted(set(re.findall(r"\b(b200sv_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no prototypes found in include/b200sv.h"
    for name in declared:
        assert hasattr(L, name), "libb200sv.so does not export %s" % name
    assert sorted(L._b200sv_symbols) == declared  # the ctypes binding covers the whole header


def test_no_device_means_error_not_fallback(resources):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(mb.B200svError) as e:
        mb.StateValidityEngine.for_robot("panda", "panda_arm", size=(1, 1, 1), origin=(0, 0, 0.5), resolution=0.02,
                                         max_distance=0.25, device=0)
    assert e.value.code == _lib.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_host_only_context_refuses_compute():
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=(1, 1, 1), origin=(0, 0, 0.5),
                                           resolution=0.02, max_distance=0.25, device=mb.DEVICE_NONE)
    assert eng.num_cells == (50, 50, 50) and eng.max_distance_sq == 169
    q = np.zeros((4, 7))
    for call in (lambda: eng.check_batch(q), lambda: eng.fk_batch(q), lambda: eng.field_add_points(np.zeros((1, 3))),
                 lambda: eng.field_get_d2(), lambda: eng.gather_roofline(1 << 20, 1),
                 lambda: eng.check_motions(q[:2], q[2:], 0.1), lambda: eng.collision_gradients(q),
                 lambda: eng.check_contacts(q, max_contacts=2), lambda: eng.field_write_df(),
                 lambda: eng.field_read_df(b"resolution: 0.02\n"), lambda: eng.field_add_sphere((0, 0, 0.5), 0.1),
                 lambda: eng.field_update_points(np.zeros((1, 3)), np.ones((1, 3)))):
        with pytest.raises(mb.B200svError) as e:
            call()
        assert e.value.code == _lib.ERR_NO_DEVICE


def test_bad_inputs_are_errors():
    with pytest.raises(mb.B200svError) as e:
        mb.StateValidityEngine.from_urdf("<robot><link name='a'/>", "", "g", (1, 1, 1), (0, 0, 0), 0.02, 0.25,
                                         device=mb.DEVICE_NONE)
    assert e.value.code == _lib.ERR_PARSE
    with pytest.raises(mb.B200svError) as e:  # unknown group: the reference warns "No group %s" (:716-720)
        mb.StateValidityEngine.for_robot("panda", "nope", size=(1, 1, 1), origin=(0, 0, 0.5), resolution=0.02,
                                         max_distance=0.25, device=mb.DEVICE_NONE)
    assert e.value.code == _lib.ERR_INVALID
    with pytest.raises(mb.B200svError) as e:  # 250 propagation cells: beyond the EDT's 16-bit lanes
        mb.StateValidityEngine.for_robot("panda", "panda_arm", size=(0.2, 0.2, 0.2), origin=(0, 0, 0.5), resolution=0.001,
                                         max_distance=0.25, device=mb.DEVICE_NONE)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    # signed fields: the thresholds of the integer predicate exist unless tolerance > radius + resolution
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=(1, 1, 1), origin=(0, 0, 0.5), resolution=0.02,
                                           max_distance=0.25, device=mb.DEVICE_NONE, signed=True)
    eng.close()
    with pytest.raises(mb.B200svError) as e:
        mb.StateValidityEngine.for_robot("panda", "panda_arm", size=(1, 1, 1), origin=(0, 0, 0.5), resolution=0.02,
                                         max_distance=0.25, tolerance=0.2, device=mb.DEVICE_NONE, signed=True)
    assert e.value.code == _lib.ERR_UNSUPPORTED


CASES = [("panda", "panda_spheres.urdf", "panda.srdf", "panda_arm"),
         ("panda", "panda_spheres.urdf", "panda.srdf", "panda_arm_hand"),
         ("panda", "panda_spheres.urdf", "panda.srdf", "hand"),
         ("dualarm", "dualarm_spheres.urdf", "dualarm.srdf", "arms"),
         ("dualarm", "dualarm_spheres.urdf", "dualarm.srdf", "left_arm"),
         ("dualarm", "dualarm_spheres.urdf", "dualarm.srdf", "whole_body"),
         ("panda_primitives", "panda_primitives.urdf", "panda.srdf", "panda_arm"),
         ("panda_primitives", "panda_primitives.urdf", "panda.srdf", "panda_arm_hand")]


@pytest.mark.parametrize("robot,urdf,srdf,group", CASES)
def test_cpp_model_builder_matches_oracle(resources, robot, urdf, srdf, group):
    res, maxd = 0.02, 0.25
    size, origin = (1.5, 1.5, 1.5), (0.0, 0.0, 0.6)
    eng = mb.StateValidityEngine.for_robot(robot, group, size=size, origin=origin, resolution=res, max_distance=maxd,
                                           device=mb.DEVICE_NONE)
    m = RobotModel(read(resources, urdf), read(resources, srdf))
    env = OracleEnv(m, group, size, origin, res, maxd)
    ra, a = eng.robot_arrays(), env.arr
    assert eng.link_names() == [l.name for l in m.links]
    for k in ("parent", "joint_type", "first_var", "group_var_index"):
        assert np.array_equal(ra[k], a[k]), k
    for k, ok in (("joint_axis", "axis"), ("joint_origin", "origin")):
        assert np.array_equal(ra[k], a[ok]), k  # bit-identical doubles: same formulas, same order
    assert np.array_equal(ra["base_positions"], env.base_positions)
    ma = eng.model_arrays()
    assert np.array_equal(ma["glink_index"], a["glink_index"])
    assert np.array_equal(ma["sphere_start"], a["sphere_start"])
    assert np.array_equal(ma["self_enabled"].astype(bool), a["self_enabled"].astype(bool))
    n = eng.n_glinks
    iu = np.triu_indices(n, 1)
    geom = a["has_geometry"].astype(bool)
    both = geom[iu[0]] & geom[iu[1]]
    assert np.array_equal(ma["intra_enabled"][iu][both], a["intra_enabled"].reshape(n, n)[iu][both])
    ns = eng.n_spheres
    assert ns == env.n_spheres
    assert np.array_equal(ma["sphere_xyzr"][:, :3].reshape(-1), a["sphere_rel"][: 3 * ns])
    assert np.array_equal(ma["sphere_xyzr"][:, 3], a["sphere_radius"][:ns])
    assert np.allclose(ma["bounding_sphere"][:, :3].reshape(-1), a["bs_center"][: 3 * n], atol=1e-15)
    assert np.allclose(ma["bounding_sphere"][:, 3], a["bs_radius"][:n], atol=1e-15)
    lo, hi = eng.group_bounds()
    olo, ohi = env.joint_bounds()
    assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)
    sp = eng.self_points()
    assert sp.shape == env.self_points.shape and np.array_equal(sp, env.self_points)


def test_integer_thresholds_equal_reference_predicate():
    """collide <=> d2 < T must reproduce (max > dist) && (radius - dist > tol) with dist = sqrt(d2) * res
    (collision_distance_field_types.cpp:229, propagation_distance_field.cpp:99-101) for every d2."""
    for res, maxd, tol in ((0.02, 0.25, 0.0), (0.01, 0.25, 0.0), (0.01, 0.25, 0.004), (0.025, 0.3, 0.0)):
        eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=(1, 1, 1), origin=(0, 0, 0.5),
                                               resolution=res, max_distance=maxd, tolerance=tol,
                                               device=mb.DEVICE_NONE)
        thr = eng.sphere_thresholds()
        radii = eng.model_arrays()["sphere_xyzr"][:, 3]
        d2 = np.arange(eng.max_distance_sq + 1)
        dist = np.sqrt(d2.astype(np.float64)) * res
        for r, t in zip(radii, thr):
            ref = (maxd > dist) & (r - dist > tol)
            assert np.array_equal(ref, d2 < t), (res, r, t)
        assert eng.info.field_bytes_per_voxel == 1


def test_reference_interface_mirror_gates_on_group():
    eng = mb.StateValidityEngine.for_robot("panda", "panda_arm", size=(1, 1, 1), origin=(0, 0, 0.5),
                                           resolution=0.02, max_distance=0.25, device=mb.DEVICE_NONE)
    env = mb.CollisionEnvB200(eng, "panda_arm")
    res = mb.CollisionResult()
    env.checkCollision(mb.CollisionRequest(group_name="other"), res, np.zeros(7))  # unknown group: nothing checked
    assert not res.collision
    assert env.distanceSelf() == 0.0 and env.distanceRobot() == 0.0
    bm = env.checkCollisionBatch(mb.CollisionRequest(group_name="other"), np.zeros((10, 7)))
    assert bm.tolist() == [0xFF, 0xFF]


def test_header_is_plain_c_and_links(resources, tmp_path):
    """include/b200sv.h compiles as C99 (-pedantic -Werror) and a C program drives the library through it alone."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    _lib.load()  # fails loudly if the library is not built
    exe = str(tmp_path / "c_abi_check")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "tests", "c_abi_check.c"), "-L", libdir, "-lb200sv", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, os.path.join(resources, "panda_spheres.urdf"), os.path.join(resources, "panda.srdf")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c abi ok" in out.stdout and "no CPU fallback" in out.stdout


def test_multi_shape_attached_bodies_carry_the_whole_body_cull():
    """The reference culls a multi-shape attached body as a whole (any shape's bounding sphere near => all its spheres are
    tested, collision_env_distance_field.cpp:455-507); the ABI carries one entry per shape plus body ids.  Ordinary bodies:
    the library proves per pair that the cull passes whenever spheres are in reach and needs nothing more; bounding radii of
    a metre (where the reference's squared-distance-vs-radius-sum quirk can cull spheres in reach) get the other shapes'
    bounding spheres chained to the pair (GPU parity: test_a_body_of_several_shapes_is_culled_as_a_whole).  Shapes of one
    body must ride on one link.  Host-only contexts: no GPU needed."""
    import moveit2_b200 as mb
    from oracle.collision_model import OracleEnv
    from oracle.robot_model import RobotModel
    res = os.path.join(ROOT, "moveit2_b200", "resources")
    with open(os.path.join(res, "panda_spheres.urdf")) as f:
        urdf = f.read()
    with open(os.path.join(res, "panda.srdf")) as f:
        srdf = f.read()

    def build(length):
        # a rod of two shapes, each `length` long: bounding radius length / 2, spheres out to its ends
        env = OracleEnv(RobotModel(urdf, srdf), "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.02, 0.25)
        step = length / 6.0
        rod = [((0.0, 0.0, 0.20 + step * (k + 0.5)), 0.025) for k in range(12)]
        bodies = [dict(link="panda_hand", name="rod", spheres=rod[:6], bs_center=(0, 0, 0.20 + 0.5 * length),
                       bs_radius=0.5 * length, touch_links={"panda_hand"}),
                  dict(link="panda_hand", name="rod", spheres=rod[6:], bs_center=(0, 0, 0.20 + 1.5 * length),
                       bs_radius=0.5 * length, touch_links={"panda_hand"})]
        env.attach_bodies(bodies)
        robot, model = env.flat_arrays()
        n0 = len(env.group.updated_links)
        assert list(model["body_id"]) == [-1] * n0 + [0, 0]
        return robot, model

    for length in (0.3, 2.0):
        robot, model = build(length)
        eng = mb.StateValidityEngine.from_arrays(robot, model, (1, 1, 1), (0, 0, 0.5), 0.02, 0.25, device=mb.DEVICE_NONE)
        assert eng.n_glinks == len(model["glink_index"])
        eng.close()
    model["glink_index"] = model["glink_index"].copy()
    model["glink_index"][-1] = model["glink_index"][-3]   # the second shape on another link than the first
    with pytest.raises(mb.B200svError) as e:
        mb.StateValidityEngine.from_arrays(robot, model, (1, 1, 1), (0, 0, 0.5), 0.02, 0.25, device=mb.DEVICE_NONE)
    assert e.value.code == -1 and "same link" in str(e.value)
