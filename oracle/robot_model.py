"""ORACLE (test infrastructure, not product code): Python restatement of how MoveIt 2 turns a URDF + SRDF into
the joint tree, variable layout and joint-model groups that forward kinematics runs over.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Follows (paths relative to /root/reference/moveit_core unless noted):
  robot_model/src/robot_model.cpp:848-890      buildRecursive: DFS pre-order link/joint indices
  robot_model/src/robot_model.cpp:895-1044     jointBoundsFromURDF, constructJointModel (virtual joints from SRDF)
  robot_model/src/robot_model.cpp:1165-1254    urdfPose2Isometry3d, constructLinkModel
  robot_model/src/robot_model.cpp:404-471      buildMimic (chains of mimics are flattened)
  robot_model/src/robot_model.cpp:136-169,240-258  computeDescendants (follows mimic requests too)
  robot_model/src/robot_model.cpp:538-598,728-845  buildGroups / addJointModelGroup (chains, joints, links, subgroups)
  robot_model/src/joint_model_group.cpp:60-92,114-300  includesParent, JointModelGroup ctor (variable order, roots,
                                                       updated links, group mimic updates)
  robot_model/src/link_model.cpp:60-67         jointOriginTransformIsIdentity
Third-party behaviour restated from published sources (un-vendored, no version pin in the reference):
  urdfdom   ModelInterface::initTree iterates its std::map<string,Joint> => a link's children are ordered by
            JOINT NAME; Rotation::setFromRPY (half-angle products, then normalize); defaults axis=(1,0,0),
            mimic multiplier 1 / offset 0.
  Eigen     Quaternion::toRotationMatrix.
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

FIXED, REVOLUTE, PRISMATIC, PLANAR, FLOATING = 0, 1, 2, 3, 4
JOINT_NVARS = {FIXED: 0, REVOLUTE: 1, PRISMATIC: 1, PLANAR: 3, FLOATING: 7}
EPS = np.finfo(np.float64).eps


def _floats(s, n, default):
    if s is None:
        return list(default)
    v = [float(t) for t in s.split()]
    assert len(v) == n, s
    return v


def quat_from_rpy(roll, pitch, yaw):
    """urdfdom Rotation::setFromRPY followed by normalize()."""
    phi, the, psi = roll / 2.0, pitch / 2.0, yaw / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    y = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    s = math.sqrt(x * x + y * y + z * z + w * w)
    if s == 0.0:
        return 0.0, 0.0, 0.0, 1.0
    return x / s, y / s, z / s, w / s


def quat_to_matrix(x, y, z, w):
    """Eigen::QuaternionBase::toRotationMatrix."""
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1.0 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1.0 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1.0 - (txx + tyy)]])


def pose_to_matrix(xyz, rpy):
    """urdfPose2Isometry3d (robot_model.cpp:1165-1170): Translation * Quaternion -> 4x4."""
    m = np.eye(4)
    m[:3, :3] = quat_to_matrix(*quat_from_rpy(*rpy))
    m[:3, 3] = xyz
    return m


def is_identity_transform(m):
    """link_model.cpp:64-66: linear().isIdentity() (Eigen default prec 1e-12) && translation norm < eps."""
    return bool(np.allclose(m[:3, :3], np.eye(3), rtol=0.0, atol=1e-12) and np.linalg.norm(m[:3, 3]) < EPS)


class Shape:
    def __init__(self, kind, dims, origin):
        self.kind, self.dims, self.origin = kind, dims, origin  # origin: 4x4 collision origin transform


def parse_urdf(text):
    root = ET.fromstring(text.strip())
    links, joints = {}, {}
    for le in root.findall("link"):
        shapes = []
        for ce in le.findall("collision"):
            ge = ce.find("geometry")
            if ge is None or len(ge) == 0:
                continue
            oe = ce.find("origin")
            xyz = _floats(oe.get("xyz") if oe is not None else None, 3, (0, 0, 0))
            rpy = _floats(oe.get("rpy") if oe is not None else None, 3, (0, 0, 0))
            g = ge[0]
            if g.tag == "sphere":
                shapes.append(Shape("sphere", [float(g.get("radius"))], pose_to_matrix(xyz, rpy)))
            elif g.tag == "box":
                shapes.append(Shape("box", _floats(g.get("size"), 3, None), pose_to_matrix(xyz, rpy)))
            elif g.tag == "cylinder":
                shapes.append(Shape("cylinder", [float(g.get("radius")), float(g.get("length"))],
                                    pose_to_matrix(xyz, rpy)))
            elif g.tag == "mesh":
                shapes.append(Shape("mesh", [], pose_to_matrix(xyz, rpy)))
        links[le.get("name")] = shapes
    for je in root.findall("joint"):
        oe = je.find("origin")
        ae, lim, saf, mim = je.find("axis"), je.find("limit"), je.find("safety_controller"), je.find("mimic")
        joints[je.get("name")] = dict(
            name=je.get("name"), type=je.get("type"), parent=je.find("parent").get("link"),
            child=je.find("child").get("link"),
            xyz=_floats(oe.get("xyz") if oe is not None else None, 3, (0, 0, 0)),
            rpy=_floats(oe.get("rpy") if oe is not None else None, 3, (0, 0, 0)),
            axis=_floats(ae.get("xyz") if ae is not None else None, 3, (1, 0, 0)),
            limit=None if lim is None else (float(lim.get("lower", 0.0)), float(lim.get("upper", 0.0))),
            safety=None if saf is None else (float(saf.get("soft_lower_limit", 0.0)),
                                             float(saf.get("soft_upper_limit", 0.0))),
            mimic=None if mim is None else (mim.get("joint"), float(mim.get("multiplier", 1.0)),
                                            float(mim.get("offset", 0.0))))
    return root.get("name"), links, joints


def parse_srdf(text):
    root = ET.fromstring(text.strip())
    groups = []
    for ge in root.findall("group"):
        groups.append(dict(name=ge.get("name"),
                           chains=[(c.get("base_link"), c.get("tip_link")) for c in ge.findall("chain")],
                           joints=[j.get("name") for j in ge.findall("joint")],
                           links=[l.get("name") for l in ge.findall("link")],
                           subgroups=[g.get("name") for g in ge.findall("group")]))
    vjs = [dict(name=v.get("name"), child_link=v.get("child_link"), parent_frame=v.get("parent_frame"),
                type=v.get("type")) for v in root.findall("virtual_joint")]
    disabled = [(d.get("link1"), d.get("link2")) for d in root.findall("disable_collisions")]
    enabled = [(d.get("link1"), d.get("link2")) for d in root.findall("enable_collisions")]
    no_default = [d.get("link") for d in root.findall("disable_default_collisions")]
    return dict(groups=groups, virtual_joints=vjs, disabled=disabled, enabled=enabled, no_default=no_default)


class Joint:
    pass


class Link:
    pass


class RobotModel:
    """Joint tree in the reference's index order.  links[i]/joints[i]: joint i is the parent joint of link i
    (buildRecursive pushes one of each per URDF link, so both share the DFS pre-order index)."""

    def __init__(self, urdf_text, srdf_text):
        self.name, ulinks, ujoints = parse_urdf(urdf_text)
        self.srdf = parse_srdf(srdf_text) if srdf_text else dict(groups=[], virtual_joints=[], disabled=[],
                                                                 enabled=[], no_default=[])
        children = {n: [] for n in ulinks}
        is_child = set()
        for jn in sorted(ujoints):  # urdfdom: std::map iteration => children ordered by joint name
            j = ujoints[jn]
            children[j["parent"]].append(j)
            is_child.add(j["child"])
        roots = [n for n in ulinks if n not in is_child]
        assert len(roots) == 1, roots
        self.links, self.joints = [], []
        self.link_index, self.joint_index = {}, {}
        self._build(None, roots[0], None, ulinks, children)
        self.n_vars = sum(j.n_vars for j in self.joints)
        self._build_mimic(ujoints)
        self._compute_descendants()
        self.groups = {}
        self._build_groups()

    # --- constructJointModel / constructLinkModel / buildRecursive -----------------------------------
    def _build(self, parent_link, link_name, ujoint, ulinks, children):
        j = Joint()
        j.index = len(self.joints)
        j.first_var = 0 if not self.joints else self.joints[-1].first_var + self.joints[-1].n_vars
        j.axis = np.array([1.0, 0.0, 0.0])
        j.bounds = None
        j.mimic, j.mimic_factor, j.mimic_offset, j.mimic_requests = None, 1.0, 0.0, []
        if ujoint is not None:
            j.name = ujoint["name"]
            t = ujoint["type"]
            j.type = {"revolute": REVOLUTE, "continuous": REVOLUTE, "prismatic": PRISMATIC, "floating": FLOATING,
                      "planar": PLANAR, "fixed": FIXED}[t]
            if j.type in (REVOLUTE, PRISMATIC):
                a = np.array(ujoint["axis"], dtype=np.float64)
                # RevoluteJointModel::setAxis normalizes (revolute_joint_model.cpp:73-82); PrismaticJointModel::setAxis
                # stores the raw URDF axis (prismatic_joint_model.hpp:77-80).
                j.axis = a / math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) if j.type == REVOLUTE else a
                j.bounds = self._bounds_from_urdf(ujoint, t)
            origin = pose_to_matrix(ujoint["xyz"], ujoint["rpy"])
        else:
            j.name, j.type = "ASSUMED_FIXED_ROOT_JOINT", FIXED
            for vj in self.srdf["virtual_joints"]:
                if vj["child_link"] != link_name or not vj["parent_frame"]:
                    continue
                if vj["type"] in ("fixed", "planar", "floating"):
                    j.name = vj["name"]
                    j.type = {"fixed": FIXED, "planar": PLANAR, "floating": FLOATING}[vj["type"]]
                    break
            origin = np.eye(4)
        j.n_vars = JOINT_NVARS[j.type]
        j.parent_link = parent_link
        l = Link()
        l.index = len(self.links)
        l.name = link_name
        l.shapes = ulinks[link_name]
        l.parent = parent_link
        l.parent_joint = j
        l.origin = origin
        l.origin_is_identity = is_identity_transform(origin)
        l.child_joints = []
        j.child_link = l
        self.joints.append(j)
        self.links.append(l)
        self.joint_index[j.name] = j.index
        self.link_index[l.name] = l.index
        for cj in children[link_name]:
            l.child_joints.append(self._build(l, cj["child"], cj, ulinks, children))
        return j

    @staticmethod
    def _bounds_from_urdf(uj, t):
        """jointBoundsFromURDF (robot_model.cpp:895-926) + RevoluteJointModel defaults / setContinuous."""
        lo, hi = None, None
        if uj["safety"] is not None:
            lo, hi = uj["safety"]
            if uj["limit"] is not None:
                lo = max(lo, uj["limit"][0])
                hi = min(hi, uj["limit"][1])
        elif uj["limit"] is not None:
            lo, hi = uj["limit"]
        if t == "continuous" or (lo is None and t == "revolute"):
            return (-math.pi, math.pi)
        if lo is None:
            return (-np.finfo(np.float64).max, np.finfo(np.float64).max)
        return (lo, hi)

    def _build_mimic(self, ujoints):
        for j in self.joints:
            uj = ujoints.get(j.name)
            if uj and uj["mimic"] and uj["mimic"][0] in self.joint_index:
                src = self.joints[self.joint_index[uj["mimic"][0]]]
                if src.n_vars == j.n_vars:
                    j.mimic, j.mimic_factor, j.mimic_offset = src, uj["mimic"][1], uj["mimic"][2]
        change = True
        while change:
            change = False
            for j in self.joints:
                if j.mimic is not None and j.mimic.mimic is not None:
                    m = j.mimic
                    j.mimic, j.mimic_factor, j.mimic_offset = (m.mimic, j.mimic_factor * m.mimic_factor,
                                                               j.mimic_offset + j.mimic_factor * m.mimic_offset)
                    change = True
                if j.mimic is j:
                    for k in self.joints:
                        k.mimic = None
                    change = False
                    break
        self.mimic_joints = [j for j in self.joints if j.mimic is not None]
        for j in self.mimic_joints:
            j.mimic.mimic_requests.append(j)

    def _compute_descendants(self):
        desc = {j.index: set() for j in self.joints}  # descendant LINK indices per joint
        seen = set()

        def helper(j, parents):
            if j.index in seen:
                return
            seen.add(j.index)
            lm = j.child_link
            for p in parents:
                desc[p.index].add(lm.index)
            desc[j.index].add(lm.index)
            parents.append(j)
            for c in lm.child_joints:
                helper(c, parents)
            for m in j.mimic_requests:
                helper(m, parents)
            parents.pop()

        helper(self.joints[0], [])
        for j in self.joints:
            j.descendant_links = sorted(desc[j.index])

    # --- variables ------------------------------------------------------------------------------------
    def default_positions(self):
        """RobotModel::getVariableDefaultPositions (robot_model.cpp:1428-1433) incl. updateMimicJoints."""
        v = np.zeros(self.n_vars)
        for j in self.joints:
            if j.mimic is not None or j.n_vars == 0:
                continue
            if j.type in (REVOLUTE, PRISMATIC):
                lo, hi = j.bounds
                v[j.first_var] = 0.0 if (lo <= 0.0 <= hi) else (lo + hi) / 2.0
            elif j.type == FLOATING:
                v[j.first_var + 6] = 1.0
        for j in self.mimic_joints:
            v[j.first_var] = v[j.mimic.first_var] * j.mimic_factor + j.mimic_offset
        return v

    # --- groups ---------------------------------------------------------------------------------------
    def _build_groups(self):
        cfgs = self.srdf["groups"]
        processed = [False] * len(cfgs)
        added = True
        while added:
            added = False
            for i, gc in enumerate(cfgs):
                if processed[i] or any(s not in self.groups for s in gc["subgroups"]):
                    continue
                added = True
                processed[i] = True
                self._add_group(gc)

    def _add_group(self, gc):
        if gc["name"] in self.groups:
            return
        jset = set()
        for base, tip in gc["chains"]:
            if base not in self.link_index or tip not in self.link_index:
                continue
            base_l, lm = self.links[self.link_index[base]], self.links[self.link_index[tip]]
            cj = []
            while lm is not None:
                if lm is base_l:
                    break
                cj.append(lm.parent_joint)
                lm = lm.parent_joint.parent_link
            if lm is not base_l:
                lm, index, cj2 = base_l, 0, []
                while lm is not None:
                    for k, c in enumerate(cj):
                        if c is lm.parent_joint:
                            index = k + 1
                            break
                    if index > 0:
                        break
                    cj2.append(lm.parent_joint)
                    lm = lm.parent_joint.parent_link
                if index > 0:
                    jset.update(c.index for c in cj[:index])
                    jset.update(c.index for c in cj2)
            else:
                jset.update(c.index for c in cj)
        for jn in gc["joints"]:
            if jn in self.joint_index:
                jset.add(self.joint_index[jn])
        for ln in gc["links"]:
            if ln in self.link_index:
                jset.add(self.links[self.link_index[ln]].parent_joint.index)
        for sg in gc["subgroups"]:
            if sg in self.groups:
                jset.update(j.index for j in self.groups[sg].joints)  # active + fixed + mimic
        if not jset:
            return
        self.groups[gc["name"]] = JointModelGroup(gc["name"], sorted(jset), self)


class JointModelGroup:
    def __init__(self, name, joint_indices, model):
        self.name = name
        self.joints = [model.joints[i] for i in joint_indices]  # sorted by joint index (DFS order)
        names = {j.name for j in self.joints}
        self.joint_names = [j.name for j in self.joints]
        self.variable_index_list = []
        self.active_joints, self.mimic_joints, self.fixed_joints = [], [], []
        local = {}
        count = 0
        for j in self.joints:
            if j.n_vars > 0:
                (self.active_joints if j.mimic is None else self.mimic_joints).append(j)
                self.variable_index_list += [j.first_var + k for k in range(j.n_vars)]
                local[j.name] = count
                count += j.n_vars
            else:
                self.fixed_joints.append(j)
        self.variable_count = count

        def includes_parent(j):
            while j.parent_link is not None:
                j = j.parent_link.parent_joint
                if j.name in names and j.n_vars > 0 and j.mimic is None:
                    return True
                if j.mimic is not None:
                    m = j.mimic
                    if (m.name in names and m.n_vars > 0 and m.mimic is None) or includes_parent(m):
                        return True
            return False

        self.joint_roots = [j for j in self.active_joints if not includes_parent(j)]
        # group mimic updates as applied by RobotState::updateMimicJoints(group) (robot_state.cpp:210-220):
        # every mimic joint OF THE GROUP is refreshed from its source variable in the FULL state.
        self.mimic_updates = [(j.first_var, j.mimic.first_var, j.mimic_factor, j.mimic_offset)
                              for j in self.mimic_joints]
        self.link_names = [j.child_link.name for j in sorted(self.joints, key=lambda q: q.child_link.index)]
        upd = set()
        for r in self.joint_roots:
            upd.update(r.descendant_links)
        self.updated_links = [model.links[i] for i in sorted(upd)]
        self.updated_link_names = [l.name for l in self.updated_links]
        self.updated_links_with_geometry = [l for l in self.updated_links if l.shapes]


# ---------------------------------------------------------------------------------------------------------
# numpy FK (small cases / cross-check of the C oracle): RobotState::updateLinkTransformsInternal
# (robot_state/src/robot_state.cpp:793-848) with *JointModel::computeTransform.
def joint_transform(j, vals):
    m = np.eye(4)
    if j.type == REVOLUTE:  # revolute_joint_model.cpp:250-287
        c, s = math.cos(vals[0]), math.sin(vals[0])
        t = 1.0 - c
        x, y, z = j.axis
        txy, txz, tyz = t * (x * y), t * (x * z), t * (y * z)
        zs, ys, xs = z * s, y * s, x * s
        m[:3, :3] = [[t * (x * x) + c, txy - zs, txz + ys],
                     [txy + zs, t * (y * y) + c, tyz - xs],
                     [txz - ys, tyz + xs, t * (z * z) + c]]
    elif j.type == PRISMATIC:  # prismatic_joint_model.cpp:124-146
        m[:3, 3] = j.axis * vals[0]
    elif j.type == PLANAR:  # planar_joint_model.cpp:345-349 (Eigen AngleAxis about UnitZ)
        c, s = math.cos(vals[2]), math.sin(vals[2])
        m[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, (1.0 - c) + c]]
        m[0, 3], m[1, 3] = vals[0], vals[1]
    elif j.type == FLOATING:  # floating_joint_model.cpp:231-236
        x, y, z, w = vals[3], vals[4], vals[5], vals[6]
        n = math.sqrt(x * x + y * y + z * z + w * w)
        m[:3, :3] = quat_to_matrix(x / n, y / n, z / n, w / n)
        m[:3, 3] = vals[0:3]
    return m


def set_group_positions(model, group, positions, gstate):
    """RobotState::setJointGroupPositions (robot_state.cpp:571-584) + updateMimicJoints(group)."""
    for i, vi in enumerate(group.variable_index_list):
        positions[vi] = gstate[i]
    for dst, src, f, o in group.mimic_updates:
        positions[dst] = f * positions[src] + o
    return positions


def fk_all(model, positions):
    """Global transform of every link, link-index order (parents first)."""
    out = np.zeros((len(model.links), 4, 4))
    for l in model.links:
        j = l.parent_joint
        jt = joint_transform(j, positions[j.first_var:j.first_var + j.n_vars]) if j.n_vars else np.eye(4)
        if l.parent is not None:
            p = out[l.parent.index]
            if j.type == FIXED:
                out[l.index] = p @ l.origin
            elif l.origin_is_identity:
                out[l.index] = p @ jt
            else:
                out[l.index] = p @ l.origin @ jt
        else:
            if j.n_vars == 0:
                out[l.index] = np.eye(4)
            elif l.origin_is_identity:
                out[l.index] = jt
            else:
                out[l.index] = l.origin @ jt
    return out
