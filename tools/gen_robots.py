#!/usr/bin/env python3
"""Generate the benchmark robot descriptions (URDF + SRDF) used by tests and bench.py.

Why generated: the reference's Panda fixture (moveit_py/test/unit/fixtures/panda.urdf) references
collision *meshes* that live in an external package (moveit_resources) and the sphere decomposition of
a mesh goes through un-vendored geometric_shapes code.  Following the survey's
recommendation the benchmark robot keeps the Panda kinematic parameters (joint origins, axes, limits --
Franka's published numbers, the same the fixture carries at panda.urdf:31-222) but gives every link
`<collision><geometry><sphere/>` primitives: for a SPHERE shape the reference emits exactly one collision
sphere at the collision-origin translation with the raw radius
(collision_distance_field_types.cpp:63-67), so reference and this repo derive identical sphere sets.

Outputs (committed):
  moveit2_b200/resources/panda_spheres.urdf, panda.srdf      Panda 7-DoF + hand (configs C1, C2)
  moveit2_b200/resources/dualarm_spheres.urdf, dualarm.srdf  PR2-like dual arm on a lifting torso (C3)
"""
import math
import os
import re

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "moveit2_b200", "resources")

HALF_PI = "1.57079632679"

# name, type, parent, child, xyz, rpy, axis, (lower, upper), (soft_lower, soft_upper) | None, mimic
PANDA_JOINTS = [
    ("panda_joint1", "revolute", "panda_link0", "panda_link1", "0 0 0.333", "0 0 0", "0 0 1",
     (-2.9671, 2.9671), (-2.8973, 2.8973), None),
    ("panda_joint2", "revolute", "panda_link1", "panda_link2", "0 0 0", "-%s 0 0" % HALF_PI, "0 0 1",
     (-1.8326, 1.8326), (-1.7628, 1.7628), None),
    ("panda_joint3", "revolute", "panda_link2", "panda_link3", "0 -0.316 0", "%s 0 0" % HALF_PI, "0 0 1",
     (-2.9671, 2.9671), (-2.8973, 2.8973), None),
    ("panda_joint4", "revolute", "panda_link3", "panda_link4", "0.0825 0 0", "%s 0 0" % HALF_PI, "0 0 1",
     (-3.1416, 0.0873), (-3.0718, 0.0175), None),
    ("panda_joint5", "revolute", "panda_link4", "panda_link5", "-0.0825 0.384 0", "-%s 0 0" % HALF_PI, "0 0 1",
     (-2.9671, 2.9671), (-2.8973, 2.8973), None),
    ("panda_joint6", "revolute", "panda_link5", "panda_link6", "0 0 0", "%s 0 0" % HALF_PI, "0 0 1",
     (-0.0873, 3.8223), (-0.0175, 3.7525), None),
    ("panda_joint7", "revolute", "panda_link6", "panda_link7", "0.088 0 0", "%s 0 0" % HALF_PI, "0 0 1",
     (-2.9671, 2.9671), (-2.8973, 2.8973), None),
    ("panda_joint8", "fixed", "panda_link7", "panda_link8", "0 0 0.107", "0 0 0", None, None, None, None),
    ("panda_hand_joint", "fixed", "panda_link8", "panda_hand", "0 0 0", "0 0 -0.785398163397", None, None, None,
     None),
    ("panda_finger_joint1", "prismatic", "panda_hand", "panda_leftfinger", "0 0 0.0584", "0 0 0", "0 1 0",
     (0.0, 0.04), None, None),
    ("panda_finger_joint2", "prismatic", "panda_hand", "panda_rightfinger", "0 0 0.0584", "0 0 0", "0 -1 0",
     (0.0, 0.04), None, "panda_finger_joint1"),
]

# Sphere model (link frame, metres): our own hand-placed decomposition -- a benchmark input, not
# reference data.  Radii are multiples of 5 mm so the set of distinct thresholds stays small.
PANDA_SPHERES = {
    # link0 is outside the panda_arm group: its interior points become the SELF field.  The reference's
    # self-field check has no ACM between group links and that field (collision_env_distance_field.cpp:935
    # "TODO - deal with AllowedCollisionMatrix"), so link1's spheres must stay clear of link0's volume or
    # every state is in self collision.
    "panda_link0": [(0.0, 0.0, 0.04, 0.07), (-0.07, 0.0, 0.04, 0.07), (-0.13, 0.0, 0.04, 0.06)],
    "panda_link1": [(0.0, -0.08, 0.0, 0.06), (0.0, -0.03, 0.0, 0.06), (0.0, 0.0, -0.08, 0.06),
                    (0.0, 0.0, -0.13, 0.06)],
    "panda_link2": [(0.0, 0.0, 0.03, 0.06), (0.0, 0.0, 0.08, 0.06), (0.0, -0.12, 0.0, 0.06),
                    (0.0, -0.17, 0.0, 0.06)],
    "panda_link3": [(0.0, 0.0, -0.06, 0.05), (0.0, 0.0, -0.10, 0.06), (0.08, 0.06, 0.0, 0.055),
                    (0.08, 0.02, 0.0, 0.055)],
    "panda_link4": [(0.0, 0.0, 0.02, 0.055), (0.0, 0.0, 0.06, 0.055), (-0.08, 0.095, 0.0, 0.06),
                    (-0.08, 0.06, 0.0, 0.055)],
    "panda_link5": [(0.0, 0.055, 0.0, 0.06), (0.0, 0.075, 0.0, 0.06), (0.0, 0.0, -0.22, 0.06),
                    (0.0, 0.05, -0.18, 0.05), (0.01, 0.08, -0.14, 0.025), (0.01, 0.085, -0.11, 0.025),
                    (0.01, 0.09, -0.08, 0.025), (0.01, 0.095, -0.05, 0.025), (-0.01, 0.08, -0.14, 0.025),
                    (-0.01, 0.085, -0.11, 0.025), (-0.01, 0.09, -0.08, 0.025), (-0.01, 0.095, -0.05, 0.025)],
    "panda_link6": [(0.0, 0.0, 0.0, 0.06), (0.08, 0.03, 0.0, 0.06), (0.08, -0.01, 0.0, 0.06)],
    "panda_link7": [(0.0, 0.0, 0.07, 0.05), (0.02, 0.04, 0.08, 0.025), (0.04, 0.02, 0.08, 0.025),
                    (0.04, 0.06, 0.085, 0.02), (0.06, 0.04, 0.085, 0.02)],
    "panda_hand": [(0.0, y, z, r) for (z, r) in ((0.01, 0.03), (0.03, 0.025), (0.05, 0.025))
                   for y in (-0.075, -0.045, -0.015, 0.015, 0.045, 0.075)],
    "panda_leftfinger": [(0.0, 0.01, 0.043, 0.01), (0.0, 0.02, 0.015, 0.01)],
    "panda_rightfinger": [(0.0, -0.01, 0.043, 0.01), (0.0, -0.02, 0.015, 0.01)],
}

# panda.srdf:51-69, 80-82, 101-112 of the reference fixture: the pairs MoveIt's setup assistant disables.
PANDA_DISABLED = [
    ("panda_link0", "panda_link1", "Adjacent"), ("panda_link0", "panda_link2", "Never"),
    ("panda_link0", "panda_link3", "Never"), ("panda_link0", "panda_link4", "Never"),
    ("panda_link1", "panda_link2", "Adjacent"), ("panda_link1", "panda_link3", "Never"),
    ("panda_link1", "panda_link4", "Never"), ("panda_link2", "panda_link3", "Adjacent"),
    ("panda_link2", "panda_link4", "Never"), ("panda_link2", "panda_link6", "Never"),
    ("panda_link3", "panda_link4", "Adjacent"), ("panda_link3", "panda_link5", "Never"),
    ("panda_link3", "panda_link6", "Never"), ("panda_link3", "panda_link7", "Never"),
    ("panda_link4", "panda_link5", "Adjacent"), ("panda_link4", "panda_link6", "Never"),
    ("panda_link4", "panda_link7", "Never"), ("panda_link5", "panda_link6", "Adjacent"),
    ("panda_link6", "panda_link7", "Adjacent"),
    ("panda_hand", "panda_leftfinger", "Adjacent"), ("panda_hand", "panda_rightfinger", "Adjacent"),
    ("panda_leftfinger", "panda_rightfinger", "Default"),
    ("panda_hand", "panda_link3", "Never"), ("panda_hand", "panda_link4", "Never"),
    ("panda_hand", "panda_link6", "Never"), ("panda_hand", "panda_link7", "Adjacent"),
    ("panda_leftfinger", "panda_link3", "Never"), ("panda_leftfinger", "panda_link4", "Never"),
    ("panda_leftfinger", "panda_link6", "Never"), ("panda_leftfinger", "panda_link7", "Never"),
    ("panda_link3", "panda_rightfinger", "Never"), ("panda_link4", "panda_rightfinger", "Never"),
    ("panda_link6", "panda_rightfinger", "Never"), ("panda_link7", "panda_rightfinger", "Never"),
]


def link_xml(name, spheres):
    out = ['  <link name="%s">' % name]
    for (x, y, z, r) in spheres:
        out.append('    <collision><origin rpy="0 0 0" xyz="%r %r %r"/><geometry><sphere radius="%r"/></geometry>'
                   '</collision>' % (x, y, z, r))
    out.append("  </link>")
    return out


def joint_xml(j):
    name, typ, parent, child, xyz, rpy, axis, lim, soft, mimic = j
    out = ['  <joint name="%s" type="%s">' % (name, typ)]
    if soft:
        out.append('    <safety_controller k_position="100.0" k_velocity="40.0" soft_lower_limit="%r" '
                   'soft_upper_limit="%r"/>' % soft)
    out.append('    <origin rpy="%s" xyz="%s"/>' % (rpy, xyz))
    out.append('    <parent link="%s"/>' % parent)
    out.append('    <child link="%s"/>' % child)
    if axis:
        out.append('    <axis xyz="%s"/>' % axis)
    if lim:
        out.append('    <limit effort="50" lower="%r" upper="%r" velocity="2.0"/>' % lim)
    if mimic:
        if isinstance(mimic, tuple):
            out.append('    <mimic joint="%s" multiplier="%r" offset="%r"/>' % mimic)
        else:
            out.append('    <mimic joint="%s"/>' % mimic)
    out.append("  </joint>")
    return out


def write_robot(fname, robot_name, links, joints, spheres):
    lines = ['<?xml version="1.0"?>',
             "<!-- generated by tools/gen_robots.py; sphere collision geometry is a benchmark input -->",
             '<robot name="%s">' % robot_name]
    for l in links:
        lines += link_xml(l, spheres.get(l, []))
    for j in joints:
        lines += joint_xml(j)
    lines.append("</robot>")
    with open(os.path.join(OUT, fname), "w") as f:
        f.write("\n".join(lines) + "\n")


def gen_panda():
    links = ["panda_link%d" % i for i in range(9)] + ["panda_hand", "panda_leftfinger", "panda_rightfinger"]
    write_robot("panda_spheres.urdf", "panda", links, PANDA_JOINTS, PANDA_SPHERES)
    s = ['<?xml version="1.0"?>', '<robot name="panda">',
         '  <group name="panda_arm"><chain base_link="panda_link0" tip_link="panda_link8"/></group>',
         '  <group name="hand"><link name="panda_hand"/><link name="panda_leftfinger"/>'
         '<link name="panda_rightfinger"/><joint name="panda_finger_joint1"/>'
         '<passive_joint name="panda_finger_joint2"/></group>',
         '  <group name="panda_arm_hand"><group name="panda_arm"/><group name="hand"/></group>',
         '  <virtual_joint child_link="panda_link0" name="virtual_joint" parent_frame="world" type="floating"/>',
         '  <end_effector group="hand" name="hand" parent_group="panda_arm" parent_link="panda_link8"/>']
    for a, b, why in PANDA_DISABLED:
        s.append('  <disable_collisions link1="%s" link2="%s" reason="%s"/>' % (a, b, why))
    s.append("</robot>")
    with open(os.path.join(OUT, "panda.srdf"), "w") as f:
        f.write("\n".join(s) + "\n")


# ----------------------------------------------------------------------------------------------------
# PR2-like dual-arm robot (config C3).  The real PR2 description lives in the external
# moveit_resources_pr2_description package (absent here), so this is a synthetic tree
# sized like it: planar base, prismatic torso lift, pan/tilt head, two 7-DoF arms, and grippers whose
# fingers mimic one driving joint each (the PR2's gripper linkage is a mimic chain too).
def seg_spheres(p0, p1, r, n):
    return [tuple(round(p0[k] + (p1[k] - p0[k]) * i / max(n - 1, 1), 4) for k in range(3)) + (r,)
            for i in range(n)]


def gen_panda_primitives():
    """panda_primitives.urdf: the Panda kinematics with <cylinder> / <box> collision geometry on some links (link0, outside the
    group, feeds the self field), spheres elsewhere -- exercises BodyDecomposition's bounding-cylinder spheres and the interior
    lattice of non-sphere bodies on the URDF path."""
    with open(os.path.join(OUT, "panda_spheres.urdf")) as f:
        u = f.read()

    def replace_link(urdf, name, collisions):
        pat = re.compile(r'(<link name="%s">)(.*?)(</link>)' % name, re.S)
        body = "\n" + "".join('    <collision><origin rpy="%s" xyz="%s"/><geometry>%s</geometry></collision>\n' % c
                              for c in collisions) + "  "
        return pat.sub(lambda m: m.group(1) + body + m.group(3), urdf)

    u = replace_link(u, "panda_link0", [("0 0 0", "-0.04 0.0 0.07", '<box size="0.22 0.19 0.13"/>')])
    u = replace_link(u, "panda_link2", [("0 0 0", "0.0 0.0 0.02", '<cylinder radius="0.055" length="0.24"/>')])
    u = replace_link(u, "panda_link4", [("1.5707963 0 0.3", "-0.04 0.02 0.0", '<cylinder radius="0.05" length="0.21"/>')])
    u = replace_link(u, "panda_link5", [("0.1 0.2 0", "0.0 0.04 -0.18", '<box size="0.09 0.26 0.11"/>'),
                                        ("0 0 0", "0.0 0.06 -0.03", '<sphere radius="0.055"/>')])
    u = replace_link(u, "panda_link6", [("0 1.5707963 0", "0.04 0.0 0.0", '<cylinder radius="0.045" length="0.17"/>')])
    u = replace_link(u, "panda_hand", [("0 0 0", "0.0 0.0 0.03", '<box size="0.05 0.2 0.07"/>')])
    with open(os.path.join(OUT, "panda_primitives.urdf"), "w") as f:
        f.write(u)


def gen_dualarm_dense(copies=4):
    """dualarm_dense_spheres.urdf: the dual arm with every collision sphere replaced by `copies` slightly smaller ones
    spread around it -- 280 spheres on 28 links, the sphere count of a PR2-class decomposition (workload C3D)."""
    offs = [(0.0, 0.0, 0.0), (0.015, 0.0, 0.0), (-0.015, 0.0, 0.0), (0.0, 0.015, 0.0), (0.0, -0.015, 0.0), (0.0, 0.0, 0.015)]
    with open(os.path.join(OUT, "dualarm_spheres.urdf")) as f:
        urdf = f.read()

    def rep(m):
        x, y, z = (float(v) for v in m.group(1).split())
        r = float(m.group(2))
        return "".join('<collision><origin rpy="0 0 0" xyz="%r %r %r"/><geometry><sphere radius="%r"/></geometry></collision>'
                       % (x + o[0], y + o[1], z + o[2], r - 0.01) for o in offs[:copies])

    dense = re.sub(r'<collision><origin rpy="0 0 0" xyz="([^"]+)"/><geometry><sphere radius="([^"]+)"/></geometry></collision>',
                   rep, urdf)
    with open(os.path.join(OUT, "dualarm_dense_spheres.urdf"), "w") as f:
        f.write(dense)


def gen_dualarm():
    links = ["base_link", "torso_lift_link", "head_pan_link", "head_tilt_link"]
    joints = [
        ("torso_lift_joint", "prismatic", "base_link", "torso_lift_link", "-0.05 0 0.74", "0 0 0", "0 0 1",
         (0.0, 0.33), (0.0115, 0.325), None),
        ("head_pan_joint", "revolute", "torso_lift_link", "head_pan_link", "-0.017 0 0.381", "0 0 0", "0 0 1",
         (-3.007, 3.007), (-2.857, 2.857), None),
        ("head_tilt_joint", "revolute", "head_pan_link", "head_tilt_link", "0.068 0 0", "0 0 0", "0 1 0",
         (-0.471, 1.396), (-0.3712, 1.29626), None),
    ]
    spheres = {
        "base_link": seg_spheres((-0.2, -0.2, 0.15), (0.2, -0.2, 0.15), 0.15, 3)
        + seg_spheres((-0.2, 0.0, 0.15), (0.2, 0.0, 0.15), 0.15, 3)
        + seg_spheres((-0.2, 0.2, 0.15), (0.2, 0.2, 0.15), 0.15, 3)
        + seg_spheres((-0.22, 0.0, 0.4), (-0.22, 0.0, 0.7), 0.15, 3),
        # the torso is outside the arm groups => part of the SELF field; keep it clear of the shoulder columns
        "torso_lift_link": seg_spheres((-0.3, -0.15, 0.05), (-0.3, 0.15, 0.05), 0.12, 3)
        + seg_spheres((-0.3, -0.15, 0.3), (-0.3, 0.15, 0.3), 0.12, 3),
        "head_pan_link": [(0.0, 0.0, 0.05, 0.08)],
        "head_tilt_link": seg_spheres((0.05, -0.12, 0.1), (0.05, 0.12, 0.1), 0.1, 4),
    }
    disabled = [("base_link", "torso_lift_link", "Adjacent"), ("torso_lift_link", "head_pan_link", "Adjacent"),
                ("head_pan_link", "head_tilt_link", "Adjacent"), ("torso_lift_link", "head_tilt_link", "Never"),
                ("base_link", "head_pan_link", "Never"), ("base_link", "head_tilt_link", "Never")]
    for side, sy in (("l", 1.0), ("r", -1.0)):
        p = side + "_"
        arm = [
            (p + "shoulder_pan_joint", "revolute", "torso_lift_link", p + "shoulder_pan_link",
             "0 %r 0" % (0.188 * sy), "0 0 0", "0 0 1",
             ((-0.7146, 2.2854) if side == "l" else (-2.2854, 0.7146)),
             ((-0.5646, 2.1354) if side == "l" else (-2.1354, 0.5646)), None),
            (p + "shoulder_lift_joint", "revolute", p + "shoulder_pan_link", p + "shoulder_lift_link",
             "0.1 0 0", "0 0 0", "0 1 0", (-0.5236, 1.3963), (-0.3536, 1.2963), None),
            (p + "upper_arm_roll_joint", "revolute", p + "shoulder_lift_link", p + "upper_arm_roll_link",
             "0 0 0", "0 0 0", "1 0 0", ((-0.8, 3.9) if side == "l" else (-3.9, 0.8)),
             ((-0.65, 3.75) if side == "l" else (-3.75, 0.65)), None),
            (p + "upper_arm_joint", "fixed", p + "upper_arm_roll_link", p + "upper_arm_link", "0 0 0", "0 0 0",
             None, None, None, None),
            (p + "elbow_flex_joint", "revolute", p + "upper_arm_link", p + "elbow_flex_link", "0.4 0 0", "0 0 0",
             "0 1 0", (-2.3213, 0.0), (-2.1213, -0.15), None),
            (p + "forearm_roll_joint", "continuous", p + "elbow_flex_link", p + "forearm_roll_link", "0 0 0",
             "0 0 0", "1 0 0", None, None, None),
            (p + "forearm_joint", "fixed", p + "forearm_roll_link", p + "forearm_link", "0 0 0", "0 0 0", None,
             None, None, None),
            (p + "wrist_flex_joint", "revolute", p + "forearm_link", p + "wrist_flex_link", "0.321 0 0", "0 0 0",
             "0 1 0", (-2.18, 0.0), (-2.0, -0.1), None),
            (p + "wrist_roll_joint", "continuous", p + "wrist_flex_link", p + "wrist_roll_link", "0 0 0", "0 0 0",
             "1 0 0", None, None, None),
            (p + "gripper_palm_joint", "fixed", p + "wrist_roll_link", p + "gripper_palm_link", "0 0 0", "0 0 0",
             None, None, None, None),
            (p + "gripper_l_finger_joint", "revolute", p + "gripper_palm_link", p + "gripper_l_finger_link",
             "0.07691 0.01 0", "0 0 0", "0 0 1", (0.0, 0.548), None, None),
            (p + "gripper_r_finger_joint", "revolute", p + "gripper_palm_link", p + "gripper_r_finger_link",
             "0.07691 -0.01 0", "0 0 0", "0 0 -1", (0.0, 0.548), None, (p + "gripper_l_finger_joint", 1.0, 0.0)),
            (p + "gripper_l_finger_tip_joint", "revolute", p + "gripper_l_finger_link",
             p + "gripper_l_finger_tip_link", "0.09137 0.00495 0", "0 0 0", "0 0 -1", (0.0, 0.548), None,
             (p + "gripper_l_finger_joint", 1.0, 0.0)),
            (p + "gripper_r_finger_tip_joint", "revolute", p + "gripper_r_finger_link",
             p + "gripper_r_finger_tip_link", "0.09137 -0.00495 0", "0 0 0", "0 0 1", (0.0, 0.548), None,
             (p + "gripper_l_finger_joint", 1.0, 0.0)),
        ]
        joints += arm
        names = [j[3] for j in arm]
        links += names
        spheres[p + "shoulder_pan_link"] = seg_spheres((0.0, 0.0, -0.25), (0.0, 0.0, 0.05), 0.1, 4) \
            + [(0.1, 0.0, 0.0, 0.1)]
        spheres[p + "shoulder_lift_link"] = [(0.0, 0.0, 0.0, 0.1)]
        spheres[p + "upper_arm_link"] = seg_spheres((0.08, 0.0, 0.0), (0.38, 0.0, 0.0), 0.08, 7)
        spheres[p + "elbow_flex_link"] = [(0.0, 0.0, 0.0, 0.075)]
        spheres[p + "forearm_link"] = seg_spheres((0.06, 0.0, 0.0), (0.3, 0.0, 0.0), 0.065, 6)
        spheres[p + "wrist_flex_link"] = [(0.0, 0.0, 0.0, 0.05)]
        spheres[p + "gripper_palm_link"] = seg_spheres((0.04, -0.03, 0.0), (0.04, 0.03, 0.0), 0.04, 3) \
            + [(0.085, 0.0, 0.0, 0.04)]
        spheres[p + "gripper_l_finger_link"] = seg_spheres((0.02, 0.01, 0.0), (0.08, 0.01, 0.0), 0.02, 3)
        spheres[p + "gripper_r_finger_link"] = seg_spheres((0.02, -0.01, 0.0), (0.08, -0.01, 0.0), 0.02, 3)
        spheres[p + "gripper_l_finger_tip_link"] = seg_spheres((0.01, 0.0, 0.0), (0.03, 0.0, 0.0), 0.015, 2)
        spheres[p + "gripper_r_finger_tip_link"] = seg_spheres((0.01, 0.0, 0.0), (0.03, 0.0, 0.0), 0.015, 2)
        chain = ["torso_lift_link"] + [n for n in names if n in spheres]
        # adjacent (and adjacent-but-one) links never collide meaningfully: disable like the setup assistant does
        for a in range(len(chain)):
            for b in range(a + 1, min(a + 3, len(chain))):
                disabled.append((chain[a], chain[b], "Adjacent"))
        grip = [p + "gripper_palm_link", p + "gripper_l_finger_link", p + "gripper_r_finger_link",
                p + "gripper_l_finger_tip_link", p + "gripper_r_finger_tip_link", p + "wrist_flex_link",
                p + "forearm_link"]
        for a in range(len(grip)):
            for b in range(a + 1, len(grip)):
                if (grip[a], grip[b], "Adjacent") not in disabled:
                    disabled.append((grip[a], grip[b], "Never"))
        disabled.append((p + "shoulder_pan_link", "base_link", "Never"))
        disabled.append((p + "shoulder_pan_link", "head_pan_link", "Never"))
    disabled.append(("l_shoulder_pan_link", "r_shoulder_pan_link", "Never"))
    write_robot("dualarm_spheres.urdf", "dualarm", links, joints, spheres)
    s = ['<?xml version="1.0"?>', '<robot name="dualarm">',
         '  <virtual_joint child_link="base_link" name="world_joint" parent_frame="odom_combined" type="planar"/>',
         '  <group name="left_arm"><chain base_link="torso_lift_link" tip_link="l_wrist_roll_link"/></group>',
         '  <group name="right_arm"><chain base_link="torso_lift_link" tip_link="r_wrist_roll_link"/></group>',
         '  <group name="arms"><group name="left_arm"/><group name="right_arm"/></group>',
         '  <group name="torso"><joint name="torso_lift_joint"/></group>',
         '  <group name="whole_body"><group name="arms"/><group name="torso"/><joint name="world_joint"/>'
         '</group>']
    seen = set()
    for a, b, why in disabled:
        key = tuple(sorted((a, b)))
        if key in seen:
            continue
        seen.add(key)
        s.append('  <disable_collisions link1="%s" link2="%s" reason="%s"/>' % (a, b, why))
    s.append("</robot>")
    with open(os.path.join(OUT, "dualarm.srdf"), "w") as f:
        f.write("\n".join(s) + "\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_panda()
    gen_dualarm()
    gen_dualarm_dense()
    gen_panda_primitives()
    n = sum(len(v) for v in PANDA_SPHERES.values())
    print("panda: %d spheres (%d in group panda_arm)" % (n, n - len(PANDA_SPHERES["panda_link0"])))
