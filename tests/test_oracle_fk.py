"""Pins the ORACLE's robot-model / FK restatement on the reference's golden values
(moveit_core/robot_state/test/robot_state_test.cpp: OneRobot fixture :250-434, FK :476-599; SimpleRobot :177-245)."""
import math
import os

import numpy as np

from oracle.collision_model import OracleEnv
from oracle.robot_model import RobotModel, fk_all, quat_from_rpy, set_group_positions


def load(golden):
    return RobotModel(open(os.path.join(golden, "one_robot.urdf")).read(),
                      open(os.path.join(golden, "one_robot.srdf")).read())


def rot_to_quat(R):
    w = math.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
    return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])


def test_groups_and_order(golden):
    m = load(golden)
    g = m.groups
    # robot_state_test.cpp:490-505
    assert g["base_from_joints"].joint_names == ["base_joint", "joint_a", "joint_c"]
    assert g["base_from_base_to_tip"].joint_names == ["base_joint", "joint_a", "joint_b"]
    assert g["base_with_subgroups"].joint_names == ["base_joint", "joint_a", "joint_b", "joint_c"]
    assert g["mim_joints"].joint_names == ["mim_f", "joint_f"]
    assert "base_with_bad_subgroups" not in g
    assert g["base_from_joints"].link_names == ["base_link", "link_a", "link_c"]
    assert g["base_from_base_to_tip"].link_names == ["base_link", "link_a", "link_b"]
    assert g["base_with_subgroups"].link_names == ["base_link", "link_a", "link_b", "link_c"]
    upd = ["base_link", "link_a", "link_b", "link_c", "link_d", "link_e", "link/with/slash"]
    for n in ("base_from_joints", "base_from_base_to_tip", "base_with_subgroups"):
        assert g[n].updated_link_names == upd
    assert m.n_vars == 7  # :509
    assert g["mim_joints"].variable_count == 2  # :575


def test_fk_golden(golden):
    m = load(golden)
    v = {"base_joint/x": 0, "base_joint/y": 1, "base_joint/theta": 2, "joint_a": 3, "joint_c": 4, "mim_f": 5,
         "joint_f": 6}
    p = m.default_positions()
    p[v["base_joint/x"]], p[v["base_joint/y"]], p[v["base_joint/theta"]] = 1.0, 1.0, 0.5
    p[v["joint_a"]], p[v["joint_c"]] = -0.5, 0.08
    T = fk_all(m, p)
    li = m.link_index
    # robot_state_test.cpp:521-544 (1e-5 tolerances)
    exp = {"base_link": ((1, 1, 0), (0, 0, 0.247404, 0.968912)), "link_a": ((1, 1, 0), (0, 0, 0, 1)),
           "link_b": ((1, 1.5, 0), (0, -0.2084598, 0, 0.97803091)), "link_c": ((1.08, 1.4, 0), (0, 0, 0, 1))}
    for name, (t, q) in exp.items():
        M = T[li[name]]
        assert np.allclose(M[:3, 3], t, atol=1e-5), name
        assert np.allclose(rot_to_quat(M[:3, :3]), q, atol=1e-5), name


def test_mimic_golden(golden):
    m = load(golden)
    li = m.link_index
    p = m.default_positions()
    T = fk_all(m, p)  # :562-566
    assert np.allclose(T[li["link_c"]][:3, 3], (0.0, 0.4, 0), atol=1e-9)
    assert np.allclose(T[li["link_d"]][:3, 3], (0.2, 0.5, 0), atol=1e-9)
    assert np.allclose(T[li["link_e"]][:3, 3], (0.3, 0.6, 0), atol=1e-9)
    # setJointGroupPositions(g_mim, {mim_f (overwritten), joint_f = 1.0}) :591-598
    g = m.groups["mim_joints"]
    set_group_positions(m, g, p, [123.0, 1.0])
    assert p[5] == 1.6 and p[6] == 1.0
    T = fk_all(m, p)
    assert np.allclose(T[li["link_c"]][:3, 3], (0.0, 0.4, 0), atol=1e-9)
    assert np.allclose(T[li["link_d"]][:3, 3], (1.7, 0.5, 0), atol=1e-9)
    assert np.allclose(T[li["link_e"]][:3, 3], (2.8, 0.6, 0), atol=1e-9)


def test_simple_robot_planar():
    # LoadingAndFK.SimpleRobot (robot_state_test.cpp:177-245): planar virtual joint, FK of base_link
    urdf = '<robot name="myrobot"><link name="base_link"><collision><geometry><box size="1 2 1"/></geometry>' \
           '</collision></link></robot>'
    srdf = '<robot name="myrobot"><virtual_joint name="base_joint" child_link="base_link" ' \
           'parent_frame="odom_combined" type="planar"/><group name="base"><joint name="base_joint"/></group></robot>'
    m = RobotModel(urdf, srdf)
    assert m.n_vars == 3 and len(m.joints) == 1
    p = m.default_positions()
    p[:] = (10.0, 8.0, 0.1)
    assert np.allclose(fk_all(m, p)[0][:3, 3], (10, 8, 0))
    p[:] = (5.0, 4.0, 0.0)
    assert np.allclose(fk_all(m, p)[0][:3, 3], (5, 4, 0))


def test_c_oracle_fk_matches_numpy(resources):
    """The C restatement (used for everything large) agrees with the numpy one to the last bits."""
    for urdf, srdf, group in (("panda_spheres.urdf", "panda.srdf", "panda_arm"),
                              ("dualarm_spheres.urdf", "dualarm.srdf", "whole_body")):
        m = RobotModel(open(os.path.join(resources, urdf)).read(), open(os.path.join(resources, srdf)).read())
        env = OracleEnv(m, group, (1, 1, 1), (0, 0, 0.5), 0.05, 0.25)
        lo, hi = env.joint_bounds()
        q = np.random.default_rng(0).uniform(lo, hi, size=(50, len(lo)))
        Tc = env.fk(q)
        for i in range(len(q)):
            p = set_group_positions(m, env.group, env.base_positions.copy(), q[i])
            assert np.allclose(Tc[i], fk_all(m, p), atol=1e-14, rtol=0)


def test_panda_structure(resources):
    m = RobotModel(open(os.path.join(resources, "panda_spheres.urdf")).read(),
                   open(os.path.join(resources, "panda.srdf")).read())
    g = m.groups["panda_arm"]
    assert m.n_vars == 16 and g.variable_index_list == list(range(7, 14))
    assert g.updated_link_names == ["panda_link%d" % i for i in range(1, 9)] + ["panda_hand", "panda_leftfinger",
                                                                               "panda_rightfinger"]
    # ready pose sanity: hand above the table, in front of the base
    env = OracleEnv(m, "panda_arm", (1, 1, 1), (0, 0, 0.5), 0.05, 0.25)
    T = env.fk(np.array([[0, -0.785, 0, -2.356, 0, 1.571, 0.785]]))[0]
    hand = T[m.link_index["panda_hand"]][:3, 3]
    assert 0.25 < hand[0] < 0.4 and abs(hand[1]) < 1e-3 and 0.5 < hand[2] < 0.7


def test_quat_from_rpy_normalised():
    q = quat_from_rpy(0.3, -0.42, 1.2)
    assert abs(sum(v * v for v in q) - 1.0) < 1e-15
