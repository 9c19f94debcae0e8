"""Python host side above the C-ABI: a thin, numpy-facing wrapper of one b200sv context plus a mirror of the
reference's CollisionEnv-style entry points so parity tests read like the reference's own tests.

Nothing here computes: every method forwards to libb200sv.so and raises B200svError on a non-zero status.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import (CHECK_ALL, CHECK_FP64, CHECK_INTRA, CHECK_SELF, CHECK_WORLD, DEVICE_NONE, FIELD_SELF,  # noqa: F401
                   FIELD_WORLD)

RESOURCES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources")


class B200svError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200sv error %d: %s" % (code, msg))
        self.code = code


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def field_cfg(size, origin, resolution, max_distance, tolerance=0.0, signed=False):
    f = _lib.FieldCfg()
    f.size[:] = [float(v) for v in size]
    f.origin[:] = [float(v) for v in origin]
    f.resolution = float(resolution)
    f.max_propagation_distance = float(max_distance)
    f.collision_tolerance = float(tolerance)
    f.use_signed_distance_field = int(bool(signed))
    return f


class StateValidityEngine:
    """One b200sv context: a robot + planning group + world/self distance fields on one CUDA device."""

    def __init__(self, handle):
        self.L = _lib.load()
        self.h = handle
        inf = _lib.Info()
        self._check(self.L.b200sv_get_info(self.h, C.byref(inf)))
        self.info = inf
        self.n_links, self.n_vars, self.n_group_vars = inf.n_links, inf.n_vars, inf.n_group_vars
        self.n_spheres, self.n_glinks = inf.n_spheres, inf.n_glinks
        self.num_cells = tuple(inf.num_cells)
        self.max_distance_sq = inf.max_distance_sq

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_urdf(cls, urdf_xml, srdf_xml, group, size, origin, resolution, max_distance, tolerance=0.0, device=0,
                  signed=False):
        L = _lib.load()
        h = C.c_void_p()
        cfg = field_cfg(size, origin, resolution, max_distance, tolerance, signed)
        rc = L.b200sv_create_from_urdf(int(device), urdf_xml.encode(), (srdf_xml or "").encode(), group.encode(),
                                       C.byref(cfg), C.byref(h))
        if rc != 0:
            raise B200svError(rc, (L.b200sv_last_error(None) or b"").decode())
        return cls(h)

    @classmethod
    def from_arrays(cls, robot: dict, model: dict, size, origin, resolution, max_distance, tolerance=0.0, device=0):
        """robot / model: dicts of numpy arrays named after the fields of b200sv_robot / b200sv_collision_model."""
        L = _lib.load()
        keep = {}

        def arr(d, k, dt):
            a = np.ascontiguousarray(d[k], dtype=dt).reshape(-1)
            if a.size == 0:
                a = np.zeros(1, dt)
            keep[k] = a
            return a

        r = _lib.Robot()
        r.n_links, r.n_vars = int(robot["n_links"]), int(robot["n_vars"])
        r.parent = _ip(arr(robot, "parent", np.int32))
        r.joint_type = _ip(arr(robot, "joint_type", np.int32))
        r.joint_axis = _dp(arr(robot, "joint_axis", np.float64))
        r.joint_origin = _dp(arr(robot, "joint_origin", np.float64))
        r.first_var = _ip(arr(robot, "first_var", np.int32))
        r.base_positions = _dp(arr(robot, "base_positions", np.float64))
        r.n_group_vars = len(np.asarray(robot["group_var_index"]).reshape(-1))
        r.group_var_index = _ip(arr(robot, "group_var_index", np.int32))
        r.n_mimic = len(np.asarray(robot.get("mimic_dst", [])).reshape(-1))
        r.mimic_dst = _ip(arr(robot, "mimic_dst", np.int32)) if r.n_mimic else None
        r.mimic_src = _ip(arr(robot, "mimic_src", np.int32)) if r.n_mimic else None
        r.mimic_factor = _dp(arr(robot, "mimic_factor", np.float64)) if r.n_mimic else None
        r.mimic_offset = _dp(arr(robot, "mimic_offset", np.float64)) if r.n_mimic else None
        m = _lib.CollisionModel()
        m.n_glinks = int(model["n_glinks"])
        m.glink_index = _ip(arr(model, "glink_index", np.int32))
        m.self_enabled = _u8p(arr(model, "self_enabled", np.uint8))
        m.intra_enabled = _u8p(arr(model, "intra_enabled", np.uint8))
        m.sphere_start = _ip(arr(model, "sphere_start", np.int32))
        m.sphere_xyzr = _dp(arr(model, "sphere_xyzr", np.float64))
        m.bounding_sphere = _dp(arr(model, "bounding_sphere", np.float64))
        m.body_id = _ip(arr(model, "body_id", np.int32)) if model.get("body_id") is not None else None
        cfg = field_cfg(size, origin, resolution, max_distance, tolerance)
        h = C.c_void_p()
        rc = L.b200sv_create(int(device), C.byref(r), C.byref(m), C.byref(cfg), C.byref(h))
        if rc != 0:
            raise B200svError(rc, (L.b200sv_last_error(None) or b"").decode())
        return cls(h)

    @classmethod
    def for_robot(cls, robot_name, group, **kw):
        """Bundled benchmark robots: 'panda' (panda_arm, hand, panda_arm_hand) and 'dualarm' (PR2-like)."""
        files = {"panda": ("panda_spheres.urdf", "panda.srdf"), "dualarm": ("dualarm_spheres.urdf", "dualarm.srdf"),
             "dualarm_dense": ("dualarm_dense_spheres.urdf", "dualarm.srdf"),
             "panda_primitives": ("panda_primitives.urdf", "panda.srdf")}
        u, s = files[robot_name]
        with open(os.path.join(RESOURCES, u)) as f:
            urdf = f.read()
        with open(os.path.join(RESOURCES, s)) as f:
            srdf = f.read()
        return cls.from_urdf(urdf, srdf, group, **kw)

    def close(self):
        if self.h:
            self.L.b200sv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise B200svError(rc, (self.L.b200sv_last_error(self.h) or b"").decode())

    # ---- introspection --------------------------------------------------------------------------------
    def robot_arrays(self):
        n, v, g = self.n_links, self.n_vars, self.n_group_vars
        out = dict(parent=np.zeros(n, np.int32), joint_type=np.zeros(n, np.int32), joint_axis=np.zeros(3 * n),
                   joint_origin=np.zeros(16 * n), first_var=np.zeros(n, np.int32), base_positions=np.zeros(max(v, 1)),
                   group_var_index=np.zeros(g, np.int32))
        self._check(self.L.b200sv_get_robot_arrays(self.h, _ip(out["parent"]), _ip(out["joint_type"]),
                                                   _dp(out["joint_axis"]), _dp(out["joint_origin"]),
                                                   _ip(out["first_var"]), _dp(out["base_positions"]),
                                                   _ip(out["group_var_index"])))
        out["base_positions"] = out["base_positions"][:v]
        return out

    def model_arrays(self):
        n = self.n_glinks
        start = np.zeros(n + 1, np.int32)
        self._check(self.L.b200sv_get_model_arrays(self.h, None, None, None, _ip(start), None, None))
        ns = int(start[-1])
        out = dict(glink_index=np.zeros(n, np.int32), self_enabled=np.zeros(n, np.uint8),
                   intra_enabled=np.zeros(n * n, np.uint8), sphere_start=start, sphere_xyzr=np.zeros(max(4 * ns, 1)),
                   bounding_sphere=np.zeros(max(4 * n, 1)))
        self._check(self.L.b200sv_get_model_arrays(self.h, _ip(out["glink_index"]), _u8p(out["self_enabled"]),
                                                   _u8p(out["intra_enabled"]), _ip(start), _dp(out["sphere_xyzr"]),
                                                   _dp(out["bounding_sphere"])))
        out["sphere_xyzr"] = out["sphere_xyzr"][:4 * ns].reshape(-1, 4)
        out["bounding_sphere"] = out["bounding_sphere"][:4 * n].reshape(-1, 4)
        out["intra_enabled"] = out["intra_enabled"].reshape(n, n)
        return out

    def link_names(self):
        buf = C.create_string_buffer(1 << 16)
        self._check(self.L.b200sv_get_link_names(self.h, buf, len(buf)))
        return buf.value.decode().split("\n")

    def group_bounds(self):
        lo, hi = np.zeros(self.n_group_vars), np.zeros(self.n_group_vars)
        self._check(self.L.b200sv_get_group_bounds(self.h, _dp(lo), _dp(hi)))
        return lo, hi

    def sphere_thresholds(self):
        t = np.zeros(max(self.n_spheres, 1), np.int32)
        self._check(self.L.b200sv_get_sphere_thresholds(self.h, _ip(t)))
        return t[: self.n_spheres]

    def self_points(self):
        n = C.c_size_t(0)
        self._check(self.L.b200sv_get_self_points(self.h, None, 0, C.byref(n)))
        pts = np.zeros((max(n.value, 1), 3))
        self._check(self.L.b200sv_get_self_points(self.h, _dp(pts), n.value, C.byref(n)))
        return pts[: n.value]

    def stats(self):
        s = _lib.Stats()
        self._check(self.L.b200sv_get_stats(self.h, C.byref(s)))
        return dict(n_states=s.n_states, n_valid=s.n_valid, n_rechecked=s.n_rechecked, check_ms=s.check_ms,
                    edt_ms=s.edt_ms)

    # ---- distance fields ------------------------------------------------------------------------------
    def field_reset(self, which=FIELD_WORLD):
        self._check(self.L.b200sv_field_reset(self.h, which))

    def field_set_propagation(self, mode, which=FIELD_WORLD):
        """_lib.PROPAGATION_EXACT (default: exact truncated EDT) or _lib.PROPAGATION_REFERENCE (the reference's bucket-queue
        propagation reproduced on the device: distances equal PropagationDistanceField's after the same calls).  Resets the
        field."""
        self._check(self.L.b200sv_field_set_propagation(self.h, which, int(mode)))

    def field_propagation_stats(self, which=FIELD_WORLD):
        st = _lib.PropagationStats()
        self._check(self.L.b200sv_field_propagation_stats(self.h, which, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in st._fields_}

    def field_begin_update(self, which=FIELD_WORLD):
        """Bracket several occupancy edits; field_end_update runs the EDT once."""
        self._check(self.L.b200sv_field_begin_update(self.h, which))

    def field_end_update(self, which=FIELD_WORLD):
        self._check(self.L.b200sv_field_end_update(self.h, which))

    def field_add_points(self, points, which=FIELD_WORLD):
        p = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        self._check(self.L.b200sv_field_add_points(self.h, which, _dp(p), p.shape[0]))

    def field_remove_points(self, points, which=FIELD_WORLD):
        p = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        self._check(self.L.b200sv_field_remove_points(self.h, which, _dp(p), p.shape[0]))

    def field_add_sphere(self, center, radius, which=FIELD_WORLD):
        """DistanceField::addShapeToField for a sphere: interior lattice points marked on the device."""
        c = np.ascontiguousarray(center, np.float64).reshape(3)
        self._check(self.L.b200sv_field_add_sphere(self.h, which, _dp(c), float(radius)))

    def field_add_shape(self, shape_type, dims, pose, which=FIELD_WORLD, lattice_frame=_lib.LATTICE_WORLD, remove=False):
        """Sphere / cylinder / box decomposed on the device.  pose: 4x4 (or 3x4) homogeneous matrix.  lattice_frame:
        LATTICE_WORLD = DistanceField::addShapeToField, LATTICE_BODY = BodyDecomposition + PosedBodyPointDecomposition."""
        sh = _lib.Shape()
        sh.type, sh.lattice_frame = int(shape_type), int(lattice_frame)
        d = list(np.asarray(dims, np.float64).reshape(-1)) + [0.0, 0.0, 0.0]
        for k in range(3):
            sh.dims[k] = float(d[k])
        m = np.asarray(pose, np.float64)
        for r in range(3):
            for k in range(4):
                sh.pose[4 * r + k] = float(m[r, k])
        self._check(self.L.b200sv_field_add_shape(self.h, which, C.byref(sh), int(bool(remove))))

    def field_add_convex(self, planes, bounding_sphere, pose, which=FIELD_WORLD, lattice_frame=_lib.LATTICE_WORLD,
                         remove=False):
        """A convex mesh given by its hull planes [n, 4] (unit normal, w), decomposed on the device."""
        pl = np.ascontiguousarray(planes, np.float64).reshape(-1, 4)
        cv = _lib.Convex()
        cv.n_planes, cv.lattice_frame, cv.planes = pl.shape[0], int(lattice_frame), _dp(pl)
        for k in range(4):
            cv.bounding_sphere[k] = float(bounding_sphere[k])
        m = np.asarray(pose, np.float64)
        for r in range(3):
            for k in range(4):
                cv.pose[4 * r + k] = float(m[r, k])
        self._check(self.L.b200sv_field_add_convex(self.h, which, C.byref(cv), int(bool(remove))))

    def field_add_octree_leaves(self, leaves, which=FIELD_WORLD):
        """Occupied octree leaves [n, 4] = centre xyz, size (DistanceField::addOcTreeToField behind octomap's iteration)."""
        a = np.ascontiguousarray(leaves, np.float64).reshape(-1, 4)
        self._check(self.L.b200sv_field_add_octree_leaves(self.h, which, _dp(a) if len(a) else None, a.shape[0]))

    def field_update_points(self, old_points, new_points, which=FIELD_WORLD):
        """PropagationDistanceField::updatePointsInField."""
        a = np.ascontiguousarray(old_points, np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(new_points, np.float64).reshape(-1, 3)
        self._check(self.L.b200sv_field_update_points(self.h, which, _dp(a) if len(a) else None, a.shape[0],
                                                      _dp(b) if len(b) else None, b.shape[0]))

    def field_add_points_device(self, d_ptr, n_points, which=FIELD_WORLD, stream=0):
        self._check(self.L.b200sv_field_add_points_device(self.h, which, C.c_void_p(d_ptr), n_points,
                                                          C.c_void_p(stream)))

    def field_get_d2(self, which=FIELD_WORLD):
        out = np.zeros(self.num_cells, np.int32)
        self._check(self.L.b200sv_field_get_d2(self.h, which, _ip(out)))
        return out

    def field_set_d2(self, d2, which=FIELD_WORLD):
        a = np.ascontiguousarray(d2, np.int32).reshape(-1)
        assert a.size == self.num_cells[0] * self.num_cells[1] * self.num_cells[2]
        self._check(self.L.b200sv_field_set_d2(self.h, which, _ip(a)))

    def field_get_neg_d2(self, which=FIELD_WORLD):
        """negative_distance_square_ per voxel (signed fields; all 0 for an unsigned field)."""
        out = np.zeros(self.num_cells, np.int32)
        self._check(self.L.b200sv_field_get_neg_d2(self.h, which, _ip(out)))
        return out

    def field_set_neg_d2(self, neg_d2, which=FIELD_WORLD):
        a = np.ascontiguousarray(neg_d2, np.int32).reshape(-1)
        assert a.size == self.num_cells[0] * self.num_cells[1] * self.num_cells[2]
        self._check(self.L.b200sv_field_set_neg_d2(self.h, which, _ip(a)))

    # ---- hot path -------------------------------------------------------------------------------------
    def fk_batch(self, q, precision=64):
        """[n, n_links, 4, 4] (row, col) global link transforms."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.n_group_vars)
        n = q.shape[0]
        out = np.zeros((n, self.n_links, 16))
        self._check(self.L.b200sv_fk_batch(self.h, _dp(q), n, int(precision), _dp(out)))
        return out.reshape(n, self.n_links, 4, 4).transpose(0, 1, 3, 2)

    def sphere_centers(self, q_one):
        q = np.ascontiguousarray(q_one, np.float64).reshape(self.n_group_vars)
        out = np.zeros((max(self.n_spheres, 1), 3))
        self._check(self.L.b200sv_sphere_centers(self.h, _dp(q), _dp(out)))
        return out[: self.n_spheres]

    def check_batch_bitmap(self, q, flags=CHECK_ALL, bitmap=None):
        """Host-pointer entry point: q [n, V] float64 (any host memory; pinned memory makes the copies DMA)."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.n_group_vars)
        n = q.shape[0]
        if bitmap is None:
            bitmap = np.zeros((n + 7) // 8, np.uint8)
        self._check(self.L.b200sv_check_batch(self.h, C.c_void_p(q.ctypes.data), n, int(flags),
                                              C.c_void_p(bitmap.ctypes.data)))
        return bitmap

    def check_batch_ptr(self, q_ptr, n, flags, bitmap_ptr):
        """Same with raw host addresses (e.g. of pinned torch tensors)."""
        self._check(self.L.b200sv_check_batch(self.h, C.c_void_p(q_ptr), n, int(flags), C.c_void_p(bitmap_ptr)))

    def check_batch(self, q, flags=CHECK_ALL):
        """bool [n]: True where the state is VALID (collision-free under the selected checks)."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.n_group_vars)
        bm = self.check_batch_bitmap(q, flags)
        return np.unpackbits(bm, bitorder="little")[: q.shape[0]].astype(bool)

    def check_batch_device(self, d_q_ptr, n, flags, d_bitmap_ptr, stream=0):
        """Device-resident q (float64 [n, V]) and bitmap (4 * ceil(n / 32) bytes); asynchronous on `stream`."""
        self._check(self.L.b200sv_check_batch_device(self.h, C.c_void_p(d_q_ptr), n, int(flags),
                                                     C.c_void_p(d_bitmap_ptr), C.c_void_p(stream)))

    def check_batch_f32(self, q, flags=CHECK_ALL):
        """float32 states (b200sv_check_batch_f32; not the reference's layout): bool [n], True = valid."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.n_group_vars)
        n = q.shape[0]
        bm = np.zeros((n + 7) // 8, np.uint8)
        self._check(self.L.b200sv_check_batch_f32(self.h, C.c_void_p(q.ctypes.data), n, int(flags),
                                                  C.c_void_p(bm.ctypes.data)))
        return np.unpackbits(bm, bitorder="little")[:n].astype(bool)

    def check_batch_f32_ptr(self, q_ptr, n, flags, bitmap_ptr):
        self._check(self.L.b200sv_check_batch_f32(self.h, C.c_void_p(q_ptr), n, int(flags), C.c_void_p(bitmap_ptr)))

    def fk_batch_device(self, d_q_ptr, n, precision, d_out_ptr, stream=0):
        """Device-resident q -> device-resident [n, n_links, 16] column-major doubles; asynchronous on `stream`."""
        self._check(self.L.b200sv_fk_batch_device(self.h, C.c_void_p(d_q_ptr), n, int(precision), C.c_void_p(d_out_ptr),
                                                  C.c_void_p(stream)))

    # ---- cache-entry updates on a live context --------------------------------------------------------------
    def set_acm_masks(self, self_enabled, intra_enabled):
        a = np.ascontiguousarray(self_enabled, np.uint8).reshape(self.n_glinks)
        b = np.ascontiguousarray(intra_enabled, np.uint8).reshape(self.n_glinks * self.n_glinks)
        self._check(self.L.b200sv_set_acm_masks(self.h, _u8p(a), _u8p(b)))

    def set_base_state(self, base_positions, rebuild_self_field=True):
        a = np.ascontiguousarray(base_positions, np.float64).reshape(self.n_vars)
        self._check(self.L.b200sv_set_base_state(self.h, _dp(a), int(bool(rebuild_self_field))))

    def set_self_link_points(self, link_index, xyz_link_frame):
        a = np.ascontiguousarray(xyz_link_frame, np.float64).reshape(-1, 3)
        self._check(self.L.b200sv_set_self_link_points(self.h, int(link_index), _dp(a) if len(a) else None, a.shape[0]))

    # ---- field replication ------------------------------------------------------------------------------------
    def field_export_size(self):
        n = C.c_size_t(0)
        self._check(self.L.b200sv_field_export_size(self.h, C.byref(n)))
        return int(n.value)

    def field_export_device(self, d_buf_ptr, which=FIELD_WORLD, stream=0):
        self._check(self.L.b200sv_field_export_device(self.h, int(which), C.c_void_p(d_buf_ptr), C.c_void_p(stream)))

    def field_import_device(self, d_buf_ptr, which=FIELD_WORLD, stream=0):
        self._check(self.L.b200sv_field_import_device(self.h, int(which), C.c_void_p(d_buf_ptr), C.c_void_p(stream)))

    def check_motions(self, q_from, q_to, max_segment_length, continuous=None, distance_factor=None,
                      max_states_per_edge=0, flags=CHECK_ALL):
        """Batched edge validation (b200sv_check_motions): (valid bool [M], first_invalid int32 [M], n_states)."""
        V = self.n_group_vars
        a = np.ascontiguousarray(q_from, np.float64).reshape(-1, V)
        b = np.ascontiguousarray(q_to, np.float64).reshape(-1, V)
        if a.shape != b.shape:
            raise ValueError("q_from and q_to differ in shape")
        m = a.shape[0]
        bm = np.zeros((m + 7) // 8 + 1, np.uint8)
        first = np.full(max(m, 1), -1, np.int32)
        cont = None if continuous is None else np.ascontiguousarray(continuous, np.uint8).reshape(V)
        fac = None if distance_factor is None else np.ascontiguousarray(distance_factor, np.float64).reshape(V)
        n = C.c_uint64(0)
        self._check(self.L.b200sv_check_motions(self.h, _dp(a), _dp(b), m, float(max_segment_length),
                                                _u8p(cont) if cont is not None else None,
                                                _dp(fac) if fac is not None else None, int(max_states_per_edge),
                                                int(flags), _u8p(bm), _ip(first), C.byref(n)))
        return np.unpackbits(bm, bitorder="little")[:m].astype(bool), first[:m], int(n.value)

    def collision_gradients(self, q):
        """Batched getCollisionGradients: (distances [n, S], gradients [n, S, 3], types [n, S], closest [n, n_glinks],
        sphere_locations [n, S, 3])."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.n_group_vars)
        n, S, L = q.shape[0], self.n_spheres, self.n_glinks
        d, g = np.zeros((n, S)), np.zeros((n, S, 3))
        t, c, loc = np.zeros((n, S), np.uint8), np.zeros((n, L)), np.zeros((n, S, 3))
        self._check(self.L.b200sv_collision_gradients(self.h, _dp(q), n, _dp(d), _dp(g), _u8p(t), _dp(c), _dp(loc)))
        return d, g, t, c, loc

    def field_write_df(self, which=FIELD_WORLD):
        """PropagationDistanceField::writeToStream: the field's occupancy as .df bytes."""
        n = C.c_size_t(0)
        self._check(self.L.b200sv_field_write_df(self.h, int(which), None, 0, C.byref(n)))
        buf = np.zeros(n.value, np.uint8)
        self._check(self.L.b200sv_field_write_df(self.h, int(which), _u8p(buf), buf.size, C.byref(n)))
        return buf[: n.value].tobytes()

    def field_read_df(self, raw, which=FIELD_WORLD):
        """PropagationDistanceField::readFromStream: reset, then rebuild from the stream's occupied voxels."""
        buf = np.frombuffer(bytes(raw), np.uint8).copy()
        self._check(self.L.b200sv_field_read_df(self.h, int(which), _u8p(buf), buf.size))

    CONTACT_DTYPE = np.dtype([("body_1", np.int32), ("body_2", np.int32), ("sphere_1", np.int32), ("sphere_2", np.int32),
                              ("pos", np.float64, 3)])

    def check_contacts(self, q, flags=CHECK_ALL, max_contacts=1, max_contacts_per_pair=1):
        """Batched checkCollision with req.contacts: (collision [n], contact_count [n], contacts [n, max_contacts])."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.n_group_vars)
        n = q.shape[0]
        col, cnt = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        con = np.zeros((n, max_contacts), self.CONTACT_DTYPE)
        self._check(self.L.b200sv_check_contacts(self.h, _dp(q), n, int(flags), int(max_contacts), int(max_contacts_per_pair),
                                                 _u8p(col), _ip(cnt), con.ctypes.data_as(C.c_void_p)))
        return col, cnt, con

    # ---- per-link PosedDistanceFields of the gradient path ---------------------------------------------------
    def set_link_field(self, glink, size, corner, d2):
        sz = np.ascontiguousarray(size, np.float64).reshape(3)
        co = np.ascontiguousarray(corner, np.float64).reshape(3)
        a = None if d2 is None else np.ascontiguousarray(d2, np.int32).reshape(-1)
        self._check(self.L.b200sv_set_link_field(self.h, int(glink), _dp(sz), _dp(co), _ip(a) if a is not None else None))

    def build_link_fields(self):
        self._check(self.L.b200sv_build_link_fields(self.h))

    def get_link_field(self, glink):
        """(d2 [nx, ny, nz] int32 or None, corner)"""
        dims, corner = np.zeros(3, np.int32), np.zeros(3)
        self._check(self.L.b200sv_get_link_field(self.h, int(glink), _ip(dims), _dp(corner), None))
        if dims[0] == 0:
            return None, corner
        d2 = np.zeros(tuple(int(v) for v in dims), np.int32)
        self._check(self.L.b200sv_get_link_field(self.h, int(glink), _ip(dims), _dp(corner), _ip(d2)))
        return d2, corner

    def set_gradient_link_mask(self, enabled):
        """enabled [n_glinks, n_glinks] (entry i consults link j's field) or None (the ACM is empty: none consulted)."""
        if enabled is None:
            self._check(self.L.b200sv_set_gradient_link_mask(self.h, None))
        else:
            a = np.ascontiguousarray(enabled, np.uint8).reshape(self.n_glinks * self.n_glinks)
            self._check(self.L.b200sv_set_gradient_link_mask(self.h, _u8p(a)))

    def fp32_budget(self):
        """(eps the tables were built with, eps derived from the model), metres."""
        a, b = C.c_double(0.0), C.c_double(0.0)
        self._check(self.L.b200sv_get_fp32_budget(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_fast_centres(self, q):
        """Sphere centres [n, n_spheres, 3] (metres) as the FAST kernel's FP32 arithmetic poses them."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.n_group_vars)
        out = np.zeros((q.shape[0], max(self.n_spheres, 1), 3))
        self._check(self.L.b200sv_debug_fast_centres(self.h, _dp(q), q.shape[0], _dp(out)))
        return out[:, : self.n_spheres]

    def gather_roofline(self, n_reads=1 << 28, iters=3):
        g = C.c_double(0.0)
        self._check(self.L.b200sv_gather_roofline(self.h, n_reads, iters, C.byref(g)))
        return g.value

    def set_tuning(self, fp32_eps_m=0.0, warps_per_block=0, blocks_per_sm=0):
        self._check(self.L.b200sv_set_tuning(self.h, float(fp32_eps_m), int(warps_per_block), int(blocks_per_sm)))


class _BorrowedEngine(StateValidityEngine):
    """A context owned by a MultiDeviceEngine: never destroyed on its own."""

    def close(self):
        self.h = None


class MultiDeviceEngine:
    """Several GPUs of one box behind one handle (b200sv_create_multi*): states sharded, fields built once and replicated."""

    def __init__(self, handle):
        self.L = _lib.load()
        self.h = handle
        self.n_devices = self.L.b200sv_multi_device_count(self.h)
        self.contexts = [_BorrowedEngine(C.c_void_p(self.L.b200sv_multi_ctx(self.h, i))) for i in range(self.n_devices)]
        self.n_group_vars = self.contexts[0].n_group_vars

    @classmethod
    def from_urdf(cls, devices, urdf_xml, srdf_xml, group, size, origin, resolution, max_distance, tolerance=0.0):
        L = _lib.load()
        h = C.c_void_p()
        cfg = field_cfg(size, origin, resolution, max_distance, tolerance)
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = L.b200sv_create_multi_from_urdf(dev, len(devices), urdf_xml.encode(), (srdf_xml or "").encode(),
                                             group.encode(), C.byref(cfg), C.byref(h))
        if rc != 0:
            raise B200svError(rc, (L.b200sv_last_error(None) or b"").decode())
        return cls(h)

    def _check(self, rc):
        if rc != 0:
            raise B200svError(rc, (self.L.b200sv_multi_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            for c in self.contexts:
                c.close()
            self.L.b200sv_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field_add_points(self, points, which=FIELD_WORLD):
        p = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        self._check(self.L.b200sv_multi_field_add_points(self.h, int(which), _dp(p), p.shape[0]))

    def field_set_propagation(self, mode, which=FIELD_WORLD):
        self._check(self.L.b200sv_multi_field_set_propagation(self.h, int(which), int(mode)))

    def field_sync(self, which=FIELD_WORLD):
        self._check(self.L.b200sv_multi_field_sync(self.h, int(which)))

    def field_reset(self, which=FIELD_WORLD):
        self._check(self.L.b200sv_multi_field_reset(self.h, int(which)))

    def check_batch_ptr(self, q_ptr, n, flags, bitmap_ptr):
        self._check(self.L.b200sv_check_batch_multi(self.h, C.c_void_p(q_ptr), n, int(flags), C.c_void_p(bitmap_ptr)))

    def check_batch(self, q, flags=CHECK_ALL):
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.n_group_vars)
        n = q.shape[0]
        bm = np.zeros((n + 7) // 8 + 4, np.uint8)
        self.check_batch_ptr(q.ctypes.data, n, flags, bm.ctypes.data)
        return np.unpackbits(bm, bitorder="little")[:n].astype(bool)

    def stats(self):
        s = _lib.Stats()
        self._check(self.L.b200sv_multi_get_stats(self.h, C.byref(s)))
        return dict(n_states=s.n_states, n_valid=s.n_valid, n_rechecked=s.n_rechecked, check_ms=s.check_ms,
                    edt_ms=s.edt_ms)


# ---------------------------------------------------------------------------------------------------------
# Mirror of the reference's interface for this path (collision_detection::CollisionRequest/Result,
# CollisionEnvDistanceField::check*): same names and argument meaning, single states are batches of one.
@dataclass
class CollisionRequest:  # collision_detection/include/moveit/collision_detection/collision_common.hpp:147-186
    group_name: str = ""
    contacts: bool = False
    max_contacts: int = 1
    max_contacts_per_pair: int = 1


@dataclass
class Contact:  # collision_common.hpp:80-145 (the fields the distance-field detector fills)
    pos: np.ndarray = None
    body_name_1: str = ""
    body_name_2: str = ""


@dataclass
class CollisionResult:  # collision_common.hpp:334-370
    collision: bool = False
    contact_count: int = 0
    contacts: dict = field(default_factory=dict)

    def clear(self):
        self.collision, self.contact_count, self.contacts = False, 0, {}


class CollisionEnvB200:
    """Python twin of the C++ adapter in moveit2_b200/adapter/: CollisionEnvDistanceField's check entry points
    (collision_env_distance_field.cpp:235-272, 1382-1500) plus the batched one the north star adds."""

    def __init__(self, engine: StateValidityEngine, group_name: str):
        self.engine, self.group_name = engine, group_name

    def _body_name(self, i):
        names = self.engine.link_names()
        gi = self.engine.model_arrays()["glink_index"]
        first = {int(g): k for k, g in reversed(list(enumerate(gi)))}
        return names[gi[i]] if first[int(gi[i])] == i else "attached_%d" % i  # attached bodies carry no name in the ABI

    def _single(self, res, state, flags):
        q = np.asarray(state, np.float64).reshape(1, -1)
        if self._req.contacts:
            # contact maps as the reference fills them (collision_env_distance_field.cpp:313-331, 596-621, 1604-1623)
            left = max(int(self._req.max_contacts) - res.contact_count, 0)
            if left == 0:
                return
            col, cnt, con = self.engine.check_contacts(q, flags, left, int(self._req.max_contacts_per_pair))
            res.collision = res.collision or bool(col[0])
            for k in range(int(cnt[0])):
                c = con[0, k]
                n1 = self._body_name(int(c["body_1"]))
                n2 = {-1: "self", -2: "environment"}.get(int(c["body_2"])) or self._body_name(int(c["body_2"]))
                res.contacts.setdefault((n1, n2), []).append(Contact(pos=np.array(c["pos"]), body_name_1=n1, body_name_2=n2))
                res.contact_count += 1
            return
        valid = self.engine.check_batch(q, flags)[0]
        res.collision = res.collision or (not bool(valid))

    def _gate(self, req):
        # an unknown group checks nothing (collision_env_distance_field.cpp:716-720 logs and returns)
        self._req = req
        return req.group_name == self.group_name

    def checkSelfCollision(self, req, res, state):
        if self._gate(req):
            self._single(res, state, CHECK_SELF | CHECK_INTRA)

    def checkRobotCollision(self, req, res, state):
        if self._gate(req):
            self._single(res, state, CHECK_WORLD)

    def checkCollision(self, req, res, state):
        if self._gate(req):
            self._single(res, state, CHECK_ALL)

    def checkCollisionBatch(self, req, states, flags=CHECK_ALL):
        """N group-state vectors -> validity bitmap (bit i of byte i//8, LSB first; 1 = collision-free)."""
        if not self._gate(req):
            n = np.asarray(states).reshape(-1, self.engine.n_group_vars).shape[0]
            return np.full((n + 7) // 8, 0xFF, np.uint8)
        return self.engine.check_batch_bitmap(states, flags)

    def distanceSelf(self, *_):  # stubs in the reference too (collision_env_distance_field.hpp:109-124, 179-199)
        return 0.0

    def distanceRobot(self, *_):
        return 0.0
