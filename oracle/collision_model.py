"""ORACLE (test infrastructure): builds the flat robot / collision-model structs the C restatement consumes,
following how CollisionEnvDistanceField derives them (collision_distance_field/src/...):

  collision_env_distance_field.cpp:1049-1078   addLinkBodyDecompositions (BodyDecomposition per link with geometry)
  collision_distance_field_types.cpp:59-85     determineCollisionSpheres (SPHERE shape -> one sphere at the body pose)
  collision_distance_field_types.cpp:292-335   BodyDecomposition::init (spheres, interior points, merged bounding sphere)
  distance_field/src/find_internal_points.cpp:39-64   interior lattice points of a convex body
  collision_env_distance_field.cpp:710-958     generateDistanceFieldCacheEntry (enable masks, self-field points)
  collision_detection/src/collision_matrix.cpp:67-78, 127-138   ACM from SRDF, explicit-entry lookup

Third-party arithmetic restated from published geometric_shapes sources (un-vendored, NO version pin in the
reference => parity unpinned for these two): bodies::Sphere::containsPoint (|c-p|^2 <= r^2; its semantics ARE
pinned by test_big.df, see tests/test_oracle_df.py) and bodies::mergeBoundingSpheres.
Only SPHERE collision geometry is supported: box/cylinder/mesh decompositions go through
bodies::Body::computeBoundingCylinder which is not available here.
"""
import ctypes
import math

import numpy as np

from . import lib
from .robot_model import RobotModel

C = ctypes


class OrcRobot(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("n_vars", C.c_int32),
                ("parent", C.POINTER(C.c_int32)), ("joint_type", C.POINTER(C.c_int32)),
                ("axis", C.POINTER(C.c_double)), ("origin", C.POINTER(C.c_double)),
                ("origin_is_identity", C.POINTER(C.c_int32)), ("first_var", C.POINTER(C.c_int32)),
                ("n_group_vars", C.c_int32), ("group_var_index", C.POINTER(C.c_int32)),
                ("n_mimic", C.c_int32), ("mimic_dst", C.POINTER(C.c_int32)), ("mimic_src", C.POINTER(C.c_int32)),
                ("mimic_factor", C.POINTER(C.c_double)), ("mimic_offset", C.POINTER(C.c_double))]


class OrcCollisionModel(C.Structure):
    _fields_ = [("n_glinks", C.c_int32), ("glink_index", C.POINTER(C.c_int32)),
                ("has_geometry", C.POINTER(C.c_int32)), ("self_enabled", C.POINTER(C.c_int32)),
                ("intra_enabled", C.POINTER(C.c_uint8)), ("sphere_start", C.POINTER(C.c_int32)),
                ("sphere_rel", C.POINTER(C.c_double)), ("sphere_radius", C.POINTER(C.c_double)),
                ("bs_center", C.POINTER(C.c_double)), ("bs_radius", C.POINTER(C.c_double)),
                ("point_start", C.POINTER(C.c_int32)), ("point_rel", C.POINTER(C.c_double)),
                ("max_propagation_distance", C.c_double), ("collision_tolerance", C.c_double),
                ("body_id", C.POINTER(C.c_int32))]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def find_internal_points_sphere(center, radius, resolution):
    """findInternalPointsConvex (find_internal_points.cpp:39-64) for a bodies::Sphere: bounding sphere = the sphere,
    accumulating `pt += resolution` loops, containsPoint = |c - p|^2 <= r^2."""
    cx, cy, cz = (float(v) for v in center)
    r, res = float(radius), float(resolution)

    def axis(c):
        v = math.floor((c - r - res) / res) * res
        e = c + r + res
        out = []
        while v <= e:
            out.append(v)
            v += res
        return np.array(out)

    xs, ys, zs = axis(cx), axis(cy), axis(cz)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    dx, dy, dz = cx - X, cy - Y, cz - Z
    inside = (dx * dx + dy * dy + dz * dz) <= r * r
    return np.stack([X[inside], Y[inside], Z[inside]], axis=1)


def find_internal_points_body(shape_type, dims, pose, resolution, lattice_frame="world"):
    """findInternalPointsConvex (find_internal_points.cpp:39-64) for a posed bodies::Sphere / Cylinder / Box, restated from
    geometric_shapes' published bodies.cpp (scale 1, padding 0; un-vendored in the reference: parity unpinned for cylinders
    and boxes, spheres are pinned by test_big.df).  shape_type = shapes::ShapeType (1 sphere, 2 cylinder, 4 box).
    lattice_frame "world": DistanceField::addShapeToField (distance_field.cpp:206-238) -- body posed, lattice in the world.
    lattice_frame "body": BodyDecomposition::init at the identity pose (types.cpp:292-335), then
    PosedBodyPointDecomposition::updatePose (:388-406) moves the kept points by `pose`."""
    pose = np.asarray(pose, np.float64)
    R, t = pose[:3, :3], pose[:3, 3]
    res = float(resolution)
    d = [float(v) for v in np.asarray(dims, np.float64).reshape(-1)]
    if shape_type == 1:
        radius = d[0]
    elif shape_type == 2:
        radius2, length2 = d[0] * d[0], 1.0 * d[1] / 2.0
        radius = math.sqrt(length2 * length2 + radius2)
    elif shape_type == 4:
        h = [d[0] * (1.0 / 2.0), d[1] * (1.0 / 2.0), d[2] * (1.0 / 2.0)]
        radius = math.sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2])
    else:
        raise ValueError("sphere, cylinder or box")
    body = lattice_frame == "body"
    c = np.zeros(3) if body else t
    Rb = np.eye(3) if body else R

    def axis(ck):
        v = math.floor((ck - radius - res) / res) * res
        e = ck + radius + res
        out = []
        while v <= e:
            out.append(v)
            v += res
        return np.array(out)

    X, Y, Z = np.meshgrid(axis(c[0]), axis(c[1]), axis(c[2]), indexing="ij")
    if shape_type == 1:
        dx, dy, dz = c[0] - X, c[1] - Y, c[2] - Z
        inside = (dx * dx + dy * dy + dz * dz) <= d[0] * d[0]
    else:
        vx, vy, vz = X - c[0], Y - c[1], Z - c[2]
        a = [vx * Rb[0, k] + vy * Rb[1, k] + vz * Rb[2, k] for k in range(3)]  # v . column k
        if shape_type == 4:
            inside = (np.abs(a[0]) <= h[0]) & (np.abs(a[1]) <= h[1]) & (np.abs(a[2]) <= h[2])
        else:
            remaining = radius2 - a[0] * a[0]
            inside = ~(np.abs(a[2]) > length2) & ~(remaining < 0.0) & (a[1] * a[1] <= remaining)
    P = np.stack([X[inside], Y[inside], Z[inside]], axis=1)
    if body:
        P = np.stack([(R[r, 0] * P[:, 0] + R[r, 1] * P[:, 1] + R[r, 2] * P[:, 2]) + t[r] for r in range(3)], axis=1)
    return P


def find_internal_points_convex_planes(planes, bounding_sphere, pose, resolution, lattice_frame="world"):
    """findInternalPointsConvex (find_internal_points.cpp:39-64) for a bodies::ConvexMesh given by the planes of its hull,
    restated from geometric_shapes' published bodies.cpp (un-vendored in the reference: parity unpinned):
    containsPoint(p) = isPointInsidePlanes(i_pose_ * p): for every plane, normal . ip + w - 1e-9 <= 0 (the bounding-box
    pre-test is implied by a closed hull).  planes [n, 4] = unit normal, w; bounding_sphere = centre (body frame), radius
    (computeBoundingSphere).  lattice_frame as in find_internal_points_body."""
    pose = np.asarray(pose, np.float64)
    R, t = pose[:3, :3], pose[:3, 3]
    res = float(resolution)
    planes = np.asarray(planes, np.float64).reshape(-1, 4)
    bc, radius = np.asarray(bounding_sphere[:3], np.float64), float(bounding_sphere[3])
    body = lattice_frame == "body"
    if body:
        c = bc
        iR, it = np.eye(3), np.zeros(3)
    else:
        c = np.array([(R[r, 0] * bc[0] + R[r, 1] * bc[1]) + R[r, 2] * bc[2] + t[r] for r in range(3)])
        iR = R.T.copy()
        it = np.array([-((R[0, r] * t[0] + R[1, r] * t[1]) + R[2, r] * t[2]) for r in range(3)])

    def axis(ck):
        v = math.floor((ck - radius - res) / res) * res
        e = ck + radius + res
        out = []
        while v <= e:
            out.append(v)
            v += res
        return np.array(out)

    X, Y, Z = np.meshgrid(axis(c[0]), axis(c[1]), axis(c[2]), indexing="ij")
    B = [((iR[r, 0] * X + iR[r, 1] * Y) + iR[r, 2] * Z) + it[r] for r in range(3)]
    inside = np.ones(X.shape, bool)
    for n in planes:
        dist = (((n[0] * B[0] + n[1] * B[1]) + n[2] * B[2]) + n[3]) - 1e-9
        inside &= ~(dist > 0.0)
    P = np.stack([X[inside], Y[inside], Z[inside]], axis=1)
    if body:
        P = np.stack([(R[r, 0] * P[:, 0] + R[r, 1] * P[:, 1] + R[r, 2] * P[:, 2]) + t[r] for r in range(3)], axis=1)
    return P


def octree_leaf_points(leaves, resolution):
    """DistanceField::getOcTreePoints (distance_field/src/distance_field.cpp:240-284) behind octomap's leaf iteration:
    leaves [n, 4] = centre, size of every occupied leaf inside the field's bounding box.  A leaf no larger than a cell gives
    its centre; a larger one the lattice of the reference's three `x += resolution_` loops (accumulated)."""
    res = float(resolution)
    out = []
    for cx, cy, cz, size in np.asarray(leaves, np.float64).reshape(-1, 4):
        if size <= res:
            out.append((cx, cy, cz))
            continue
        ceil_val = math.ceil(size / res) * res / 2.0

        def axis(c):
            v, vals = c - ceil_val, []
            while v <= c + ceil_val:
                vals.append(v)
                v += res
            return vals

        xs, ys, zs = axis(cx), axis(cy), axis(cz)
        out.extend((x, y, z) for x in xs for y in ys for z in zs)
    return np.array(out, np.float64).reshape(-1, 3)


def merge_bounding_spheres(spheres):
    """geometric_shapes bodies::mergeBoundingSpheres, restated from the published source (parity unpinned)."""
    if not spheres:
        return np.zeros(3), 0.0
    c, r = np.array(spheres[0][0], dtype=np.float64), float(spheres[0][1])
    for ci, ri in spheres[1:]:
        ci = np.array(ci, dtype=np.float64)
        if ri <= 0.0:
            continue
        d = float(np.linalg.norm(ci - c))
        if d + r <= ri:
            c, r = ci, ri
        elif d + ri > r:
            delta = c - ci
            new_r = (float(np.linalg.norm(delta)) + ri + r) / 2.0
            c = delta / np.linalg.norm(delta) * (new_r - ri) + ci
            r = new_r
    return c, r


def bounding_cylinder(kind, dims, pose):
    """bodies::Cylinder / bodies::Box::computeBoundingCylinder of geometric_shapes (published bodies.cpp, scale 1, padding 0;
    un-vendored in the reference: parity unpinned).  Returns (pose 4x4, radius, length), axis = local z."""
    if kind == "cylinder":
        return pose.copy(), float(dims[0]), 2.0 * (1.0 * float(dims[1]) / 2.0)
    l2, w2, h2 = (float(d) * (1.0 / 2.0) for d in dims)
    ang = 90.0 * (math.pi / 180.0)
    c, sn = math.cos(ang), math.sin(ang)
    rot = np.eye(4)
    if l2 > w2 and l2 > h2:
        length, a, b = l2 * 2.0, w2, h2
        rot[:3, :3] = [[c, 0.0, sn], [0.0, 1.0, 0.0], [-sn, 0.0, c]]  # AngleAxis(90 deg, UnitY)
    elif w2 > h2:
        length, a, b = w2 * 2.0, h2, l2
        rot[:3, :3] = [[1.0, 0.0, 0.0], [0.0, c, -sn], [0.0, sn, c]]  # AngleAxis(90 deg, UnitX)
    else:
        length, a, b = h2 * 2.0, w2, l2
    return affine_mul(pose, rot), math.sqrt(a * a + b * b), length


def affine_mul(a, b):
    """Eigen Isometry3d product, plain left-to-right sums (the order csrc/host_model.cpp uses too)."""
    out = np.eye(4)
    for r in range(3):
        for c in range(4):
            v = a[r, 0] * b[0, c]
            v += a[r, 1] * b[1, c]
            v += a[r, 2] * b[2, c]
            v += a[r, 3] * b[3, c]
            out[r, c] = v
    return out


class BodyDecomposition:
    """collision_distance_field_types.cpp:292-335 (padding 0, scale 1): determineCollisionSpheres (:59-85) -- a SPHERE shape is
    one sphere, anything else gets spheres of the bounding cylinder's radius along its axis (num = ceil(len / (r / 2)),
    spacing = len / (num - 1), centres i = 1 .. num - 2: short bodies get NONE, the reference's behaviour) -- plus the interior
    lattice points and the merged bounding sphere.  Spheres, cylinders and boxes (bodies restated from geometric_shapes)."""

    def __init__(self, link, resolution):
        self.spheres = []  # (rel xyz, radius)
        pts = []
        bss = []
        for s in link.shapes:
            if s.kind == "sphere":
                c, r = s.origin[:3, 3].copy(), float(s.dims[0])
                self.spheres.append((c, r))
                pts.append(find_internal_points_sphere(c, r, resolution))
                bss.append((c, r))
                continue
            if s.kind not in ("cylinder", "box"):
                raise NotImplementedError("oracle supports sphere, cylinder and box collision geometry (link %s has %s)"
                                          % (link.name, s.kind))
            cp, cr, cl = bounding_cylinder(s.kind, s.dims, s.origin)
            num_points = int(math.ceil(cl / (cr / 2.0)))
            spacing = cl / ((num_points * 1.0) - 1.0) if num_points != 1 else float("inf")
            for i in range(1, num_points - 1):
                z = (-cl / 2.0) + i * spacing
                c = np.array([cp[r, 2] * z + cp[r, 3] for r in range(3)])  # pose * (0, 0, z)
                self.spheres.append((c, cr))
            ty = 2 if s.kind == "cylinder" else 4
            pts.append(find_internal_points_body(ty, s.dims, s.origin, resolution, "world"))
            if s.kind == "cylinder":
                length2 = 1.0 * float(s.dims[1]) / 2.0
                rb = math.sqrt(length2 * length2 + float(s.dims[0]) * float(s.dims[0]))
            else:
                h = [float(d) * (1.0 / 2.0) for d in s.dims]
                rb = math.sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2])
            bss.append((s.origin[:3, 3].copy(), rb))
        self.points = np.concatenate(pts, axis=0) if pts else np.zeros((0, 3))
        self.bs_center, self.bs_radius = merge_bounding_spheres(bss)


class AllowedCollisionMatrix:
    """Explicit entries only, as CollisionEnvDistanceField consults them (collision_matrix.cpp:127-138)."""

    def __init__(self, srdf):
        self.entries = {}
        for a, b in srdf["enabled"]:
            self.set_entry(a, b, False)
        for a, b in srdf["disabled"]:
            self.set_entry(a, b, True)

    def set_entry(self, a, b, allowed):
        self.entries.setdefault(a, {})[b] = allowed
        self.entries.setdefault(b, {})[a] = allowed

    def always(self, a, b):
        return self.entries.get(a, {}).get(b, False)


class OracleEnv:
    """Restated CollisionEnvDistanceField for one (robot, group, ACM): owns world and self fields."""

    def __init__(self, model: RobotModel, group_name, size, origin, resolution, max_propagation_distance,
                 collision_tolerance=0.0, use_acm=True, base_positions=None, use_signed_distance_field=False):
        self.L = lib()
        self.model, self.group = model, model.groups[group_name]
        g = self.group
        self.resolution = float(resolution)
        self.max_prop, self.tol = float(max_propagation_distance), float(collision_tolerance)
        self.size, self.origin = tuple(float(v) for v in size), tuple(float(v) for v in origin)
        self.base_positions = np.array(model.default_positions() if base_positions is None else base_positions,
                                       dtype=np.float64)
        # ---- flat robot -------------------------------------------------------------------------------
        L = model.links
        a = self.arr = {}
        a["parent"] = np.array([-1 if l.parent is None else l.parent.index for l in L], np.int32)
        a["joint_type"] = np.array([l.parent_joint.type for l in L], np.int32)
        a["axis"] = np.array([l.parent_joint.axis for l in L], np.float64).reshape(-1)
        a["origin"] = np.array([l.origin.T.reshape(-1) for l in L], np.float64).reshape(-1)  # column-major
        a["origin_is_identity"] = np.array([int(l.origin_is_identity) for l in L], np.int32)
        a["first_var"] = np.array([l.parent_joint.first_var for l in L], np.int32)
        a["group_var_index"] = np.array(g.variable_index_list, np.int32)
        a["mimic_dst"] = np.array([m[0] for m in g.mimic_updates] + [0], np.int32)
        a["mimic_src"] = np.array([m[1] for m in g.mimic_updates] + [0], np.int32)
        a["mimic_factor"] = np.array([m[2] for m in g.mimic_updates] + [0], np.float64)
        a["mimic_offset"] = np.array([m[3] for m in g.mimic_updates] + [0], np.float64)
        self.robot = OrcRobot(len(L), model.n_vars, _p(a["parent"], C.c_int32), _p(a["joint_type"], C.c_int32),
                              _p(a["axis"], C.c_double), _p(a["origin"], C.c_double),
                              _p(a["origin_is_identity"], C.c_int32), _p(a["first_var"], C.c_int32),
                              len(g.variable_index_list), _p(a["group_var_index"], C.c_int32),
                              len(g.mimic_updates), _p(a["mimic_dst"], C.c_int32), _p(a["mimic_src"], C.c_int32),
                              _p(a["mimic_factor"], C.c_double), _p(a["mimic_offset"], C.c_double))
        # ---- body decompositions for every link with geometry (addLinkBodyDecompositions) --------------
        self.decomp = {l.name: BodyDecomposition(l, self.resolution) for l in L if l.shapes}
        # ---- DistanceFieldCacheEntry masks (generateDistanceFieldCacheEntry :735-838) -------------------
        acm = AllowedCollisionMatrix(model.srdf) if use_acm else None
        gl = g.updated_links
        n = len(gl)
        has_geom = np.array([int(bool(l.shapes)) for l in gl], np.int32)
        self_en = np.ones(n, np.int32)
        intra = np.zeros((n, n), np.uint8)
        for i, li in enumerate(gl):
            if not li.shapes:
                self_en[i] = 0
                continue
            intra[i, :] = 1
            if acm is not None:
                if acm.always(li.name, li.name):
                    self_en[i] = 0
                for j in range(i + 1, n):
                    if gl[j].name == li.name or acm.always(li.name, gl[j].name):
                        intra[i, j] = 0
        sph_start, sph_rel, sph_rad, bs_c, bs_r, pt_start, pt_rel = [0], [], [], [], [], [0], []
        for l in gl:
            d = self.decomp.get(l.name)
            if d is not None:
                for c, r in d.spheres:
                    sph_rel.append(c)
                    sph_rad.append(r)
                pt_rel.append(d.points)
                bs_c.append(d.bs_center)
                bs_r.append(d.bs_radius)
            else:
                bs_c.append(np.zeros(3))
                bs_r.append(0.0)
            sph_start.append(len(sph_rad))
            pt_start.append(pt_start[-1] + (len(d.points) if d is not None else 0))
        a["glink_index"] = np.array([l.index for l in gl], np.int32)
        a["has_geometry"], a["self_enabled"], a["intra_enabled"] = has_geom, self_en, intra.reshape(-1).copy()
        a["sphere_start"] = np.array(sph_start, np.int32)
        a["sphere_rel"] = np.array(sph_rel + [np.zeros(3)], np.float64).reshape(-1)
        a["sphere_radius"] = np.array(sph_rad + [0.0], np.float64)
        a["bs_center"] = np.array(bs_c + [np.zeros(3)], np.float64).reshape(-1)
        a["bs_radius"] = np.array(bs_r + [0.0], np.float64)
        a["point_start"] = np.array(pt_start, np.int32)
        a["point_rel"] = (np.concatenate(pt_rel + [np.zeros((1, 3))], axis=0)).astype(np.float64).reshape(-1)
        self.n_spheres = len(sph_rad)
        self.cm = OrcCollisionModel(n, _p(a["glink_index"], C.c_int32), _p(a["has_geometry"], C.c_int32),
                                    _p(a["self_enabled"], C.c_int32), _p(a["intra_enabled"], C.c_uint8),
                                    _p(a["sphere_start"], C.c_int32), _p(a["sphere_rel"], C.c_double),
                                    _p(a["sphere_radius"], C.c_double), _p(a["bs_center"], C.c_double),
                                    _p(a["bs_radius"], C.c_double), _p(a["point_start"], C.c_int32),
                                    _p(a["point_rel"], C.c_double), self.max_prop, self.tol)
        # ---- fields: same geometry for world and self (:931-933, :1785-1787): corner = origin - size/2 ---
        from .df import PropagationDistanceField
        sx, sy, sz = self.size
        ox, oy, oz = self.origin
        geo = (sx, sy, sz, self.resolution, ox - 0.5 * sx, oy - 0.5 * sy, oz - 0.5 * sz, self.max_prop)
        # use_signed_distance_field_ (:1785-1797; default false, collision_env_distance_field.hpp:52)
        self.world = PropagationDistanceField(*geo, use_signed_distance_field)
        self.self_field = PropagationDistanceField(*geo, use_signed_distance_field)
        self.self_points = self.compute_self_points()
        self.self_field.add_points(self.self_points)

    def attach_bodies(self, bodies):
        """Append attached bodies to the DistanceFieldCacheEntry view (dfce->attached_body_names_ follow the links,
        collision_env_distance_field.cpp:798-822, 840-866).  bodies: list of dicts with
          link       name of the group link the body is attached to
          name       body name (ACM key)
          spheres    [(centre xyz in the LINK frame, radius), ...]   (attach pose and shape pose folded in)
          bs_center, bs_radius   bounding sphere in the link frame
          touch_links  names; the reference consults them -- and the ACM entry (link, body) -- ONLY for the link the
                       body is attached to (:801-822); every other link keeps all_true for the body.
        Masks: self_collision_enabled_ false iff ACM (body, body) is ALWAYS; body-body pairs disabled iff ALWAYS."""
        a = self.arr
        gl = self.group.updated_links
        n0 = len(gl)
        names = [l.name for l in gl]
        acm = AllowedCollisionMatrix(self.model.srdf)
        n = n0 + len(bodies)
        intra0 = a["intra_enabled"].reshape(n0, n0)
        intra = np.ones((n, n), np.uint8)
        intra[:n0, :n0] = intra0
        for i in range(n0):
            if not a["has_geometry"][i]:
                intra[i, :] = 0
        glink = list(a["glink_index"])
        has_geom, self_en = list(a["has_geometry"]), list(a["self_enabled"])
        S0 = self.n_spheres
        sph_start = list(a["sphere_start"])
        sph_rel = list(a["sphere_rel"].reshape(-1, 3)[:S0])
        sph_rad = list(a["sphere_radius"][:S0])
        bs_c = list(a["bs_center"].reshape(-1, 3)[:n0])
        bs_r = list(a["bs_radius"][:n0])
        pt_start = list(a["point_start"])
        # entries named alike are the shapes of ONE body: the reference keeps it as one entry whose bounding spheres are
        # tested shape by shape and whose spheres pass as a whole (collision_env_distance_field.cpp:455-507)
        body_names = [body["name"] for body in bodies]
        a["body_id"] = np.array([-1] * n0 + [body_names.index(nm) for nm in body_names], np.int32)
        for b, body in enumerate(bodies):
            ip = names.index(body["link"])
            glink.append(gl[ip].index)
            has_geom.append(1)
            self_en.append(0 if acm.always(body["name"], body["name"]) else 1)
            if acm.always(body["link"], body["name"]) or body["link"] in body.get("touch_links", ()):
                intra[ip, n0 + b] = 0
            for b2 in range(b + 1, len(bodies)):
                # shapes of one body are one entry in the reference: never tested against each other (i == j is skipped)
                if bodies[b2]["name"] == body["name"] or acm.always(body["name"], bodies[b2]["name"]):
                    intra[n0 + b, n0 + b2] = 0
            for cxyz, rad in body["spheres"]:
                sph_rel.append(np.asarray(cxyz, np.float64))
                sph_rad.append(float(rad))
            sph_start.append(len(sph_rad))
            bs_c.append(np.asarray(body["bs_center"], np.float64))
            bs_r.append(float(body["bs_radius"]))
            pt_start.append(pt_start[-1])
        a["glink_index"] = np.array(glink, np.int32)
        a["has_geometry"], a["self_enabled"] = np.array(has_geom, np.int32), np.array(self_en, np.int32)
        a["intra_enabled"] = intra.reshape(-1).copy()
        a["sphere_start"] = np.array(sph_start, np.int32)
        a["sphere_rel"] = np.array(sph_rel + [np.zeros(3)], np.float64).reshape(-1)
        a["sphere_radius"] = np.array(sph_rad + [0.0], np.float64)
        a["bs_center"] = np.array(bs_c + [np.zeros(3)], np.float64).reshape(-1)
        a["bs_radius"] = np.array(bs_r + [0.0], np.float64)
        a["point_start"] = np.array(pt_start, np.int32)
        self.n_spheres = len(sph_rad)
        self.cm = OrcCollisionModel(n, _p(a["glink_index"], C.c_int32), _p(a["has_geometry"], C.c_int32),
                                    _p(a["self_enabled"], C.c_int32), _p(a["intra_enabled"], C.c_uint8),
                                    _p(a["sphere_start"], C.c_int32), _p(a["sphere_rel"], C.c_double),
                                    _p(a["sphere_radius"], C.c_double), _p(a["bs_center"], C.c_double),
                                    _p(a["bs_radius"], C.c_double), _p(a["point_start"], C.c_int32),
                                    _p(a["point_rel"], C.c_double), self.max_prop, self.tol, _p(a["body_id"], C.c_int32))

    def flat_arrays(self):
        """(robot, model) dicts in the layout of b200sv_robot / b200sv_collision_model (include/b200sv.h)."""
        a = self.arr
        n = int(self.cm.n_glinks)
        S = self.n_spheres
        robot = dict(n_links=len(self.model.links), n_vars=self.model.n_vars, parent=a["parent"], joint_type=a["joint_type"],
                     joint_axis=a["axis"], joint_origin=a["origin"], first_var=a["first_var"],
                     base_positions=self.base_positions, group_var_index=a["group_var_index"],
                     mimic_dst=a["mimic_dst"][:len(self.group.mimic_updates)], mimic_src=a["mimic_src"][:len(self.group.mimic_updates)],
                     mimic_factor=a["mimic_factor"][:len(self.group.mimic_updates)],
                     mimic_offset=a["mimic_offset"][:len(self.group.mimic_updates)])
        xyzr = np.concatenate([a["sphere_rel"].reshape(-1, 3)[:S], a["sphere_radius"][:S, None]], axis=1)
        bsph = np.concatenate([a["bs_center"].reshape(-1, 3)[:n], a["bs_radius"][:n, None]], axis=1)
        model = dict(n_glinks=n, glink_index=a["glink_index"], self_enabled=a["self_enabled"].astype(np.uint8),
                     intra_enabled=a["intra_enabled"], sphere_start=a["sphere_start"], sphere_xyzr=xyzr,
                     bounding_sphere=bsph)
        if "body_id" in a:
            model["body_id"] = a["body_id"]
        return robot, model

    def compute_self_points(self):
        """:907-953 -- interior points of links with geometry that are NOT updated by the group, posed at the
        cached state (here: the base state), in getLinkModelsWithCollisionGeometry (= link index) order."""
        T = self.fk(self.group_values_of(self.base_positions)[None, :])[0]
        updated = {l.name for l in self.group.updated_links_with_geometry}
        pts = []
        for l in self.model.links:
            if not l.shapes or l.name in updated:
                continue
            P = self.decomp[l.name].points
            M = T[l.index]
            # Isometry3d * Vector3d, k in order
            pts.append((M[:3, 0] * P[:, 0:1] + M[:3, 1] * P[:, 1:2]) + M[:3, 2] * P[:, 2:3] + M[:3, 3])
        return np.concatenate(pts, axis=0) if pts else np.zeros((0, 3))

    def group_values_of(self, positions):
        return np.array([positions[i] for i in self.group.variable_index_list], np.float64)

    def fk(self, q):
        """Global link transforms [N, n_links, 4, 4] (row, col) for N group-state vectors."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, max(self.group.variable_count, 1))
        n = q.shape[0]
        out = np.zeros((n, len(self.model.links), 16))
        self.L.orc_fk_batch(C.byref(self.robot), _p(self.base_positions, C.c_double), _p(q, C.c_double), n,
                            _p(out, C.c_double))
        return out.reshape(n, -1, 4, 4).transpose(0, 1, 3, 2)  # column-major -> [row, col]

    def check(self, q, flags=7, faithful=False, threads=0):
        """collision byte per state (1 = in collision); returns (array, threads used)."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, max(self.group.variable_count, 1))
        n = q.shape[0]
        out = np.zeros(n, np.uint8)
        used = self.L.orc_check_batch(C.byref(self.robot), C.byref(self.cm), self.world.h, self.self_field.h,
                                      _p(self.base_positions, C.c_double), _p(q, C.c_double), n, int(flags),
                                      int(faithful), int(threads), _p(out, C.c_uint8))
        return out, used

    def margin(self, q1, flags=7):
        q1 = np.ascontiguousarray(q1, np.float64)
        return self.L.orc_state_margin(C.byref(self.robot), C.byref(self.cm), self.world.h, self.self_field.h,
                                       _p(self.base_positions, C.c_double), _p(q1, C.c_double), int(flags))

    def build_link_fields(self):
        """The per-link PosedDistanceFields of the GroupStateRepresentation (collision_env_distance_field.cpp:1177-1197):
        for every updated link with geometry a PropagationDistanceField of size diameter^3 of its bounding sphere, corner =
        bounding-sphere centre - size / 2 IN THE LINK FRAME (the posed decomposition is still at the identity when the
        field is built), filled with the link's interior points.  Also the matrix of (i, j) pairs getSelfProximityGradients
        consults them for (:386-411): names differ and the ACM has no entry for the pair other than NEVER."""
        from .df import PropagationDistanceField
        gl = self.group.updated_links
        n = len(gl)
        acm = AllowedCollisionMatrix(self.model.srdf)
        self.link_fields = []
        for l in gl:
            d = self.decomp.get(l.name)
            if d is None:
                self.link_fields.append(None)
                continue
            diameter = 2 * d.bs_radius
            size = np.array([diameter, diameter, diameter])
            origin = np.asarray(d.bs_center, np.float64) - 0.5 * size
            f = PropagationDistanceField(size[0], size[1], size[2], self.resolution, origin[0], origin[1], origin[2],
                                         self.max_prop)
            f.add_points(d.points)
            f.size, f.corner = size, origin
            self.link_fields.append(f)
        en = np.zeros((int(self.cm.n_glinks), int(self.cm.n_glinks)), np.uint8)
        for i in range(n):
            for j in range(n):
                if gl[i].name != gl[j].name and not acm.always(gl[i].name, gl[j].name):
                    en[i, j] = 1
        self.lf_enabled = en
        return self.link_fields, en

    def collision_gradients(self, q, with_link_fields=False):
        """getCollisionGradients for every state: (distances [n, S], gradients [n, S, 3], types [n, S], closest [n, L]).
        with_link_fields: the cache entry carries a non-empty ACM, so getSelfProximityGradients also consults the other
        links' PosedDistanceFields (build_link_fields)."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, max(self.group.variable_count, 1))
        n, S, L = q.shape[0], self.n_spheres, int(self.cm.n_glinks)
        d, g = np.zeros((n, S)), np.zeros((n, S, 3))
        t, c = np.zeros((n, S), np.uint8), np.zeros((n, L))
        if with_link_fields:
            if not hasattr(self, "link_fields"):
                self.build_link_fields()
            arr = (C.c_void_p * L)(*[(f.h if f is not None else None) for f in self.link_fields] +
                                   [None] * (L - len(self.link_fields)))
            en = np.ascontiguousarray(self.lf_enabled, np.uint8)
            self.L.orc_collision_gradients_lf(C.byref(self.robot), C.byref(self.cm), self.world.h, self.self_field.h,
                                              _p(self.base_positions, C.c_double), _p(q, C.c_double), C.c_size_t(n),
                                              _p(d, C.c_double), _p(g, C.c_double), _p(t, C.c_uint8), _p(c, C.c_double),
                                              arr, _p(en, C.c_uint8))
            return d, g, t, c
        self.L.orc_collision_gradients(C.byref(self.robot), C.byref(self.cm), self.world.h, self.self_field.h,
                                       _p(self.base_positions, C.c_double), _p(q, C.c_double), C.c_size_t(n),
                                       _p(d, C.c_double), _p(g, C.c_double), _p(t, C.c_uint8), _p(c, C.c_double))
        return d, g, t, c

    CONTACT_DTYPE = np.dtype([("body_1", np.int32), ("body_2", np.int32), ("sphere_1", np.int32), ("sphere_2", np.int32),
                              ("pos", np.float64, 3)])

    def check_contacts(self, q, flags=7, max_contacts=1, max_contacts_per_pair=1):
        """checkCollision with req.contacts = true: (collision [n], contact_count [n], contacts [n, max_contacts])."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, max(self.group.variable_count, 1))
        n = q.shape[0]
        col, cnt = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        con = np.zeros((n, max(max_contacts, 1)), self.CONTACT_DTYPE)
        con["body_1"] = con["body_2"] = con["sphere_1"] = con["sphere_2"] = -9
        self.L.orc_check_contacts(C.byref(self.robot), C.byref(self.cm), self.world.h, self.self_field.h,
                                  _p(self.base_positions, C.c_double), _p(q, C.c_double), C.c_size_t(n), int(flags),
                                  int(max_contacts), int(max_contacts_per_pair), _p(col, C.c_uint8), _p(cnt, C.c_int32),
                                  con.ctypes.data_as(C.c_void_p))
        return col, cnt, con

    def sphere_centers(self, q1):
        q1 = np.ascontiguousarray(q1, np.float64)
        out = np.zeros((self.n_spheres, 3))
        self.L.orc_state_spheres(C.byref(self.robot), C.byref(self.cm), _p(self.base_positions, C.c_double),
                                 _p(q1, C.c_double), _p(out, C.c_double))
        return out

    def joint_bounds(self):
        """Bounds of the group's variables, as setToRandomPositions samples them (soft & hard limits)."""
        lo, hi = [], []
        for j in self.group.joints:
            if j.n_vars == 1:
                lo.append(j.bounds[0])
                hi.append(j.bounds[1])
            elif j.n_vars == 3:  # planar: benchmark samples a 1 m box and full yaw
                lo += [-0.5, -0.5, -math.pi]
                hi += [0.5, 0.5, math.pi]
            elif j.n_vars == 7:
                raise NotImplementedError("floating joints in a planning group are not sampled")
        return np.array(lo), np.array(hi)
