"""ORACLE (test infrastructure): ctypes front for the C restatement of PropagationDistanceField, plus the
reference's .df occupancy file format (distance_field/src/propagation_distance_field.cpp:635-760) in numpy."""
import ctypes
import zlib

import numpy as np

from . import lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


class PropagationDistanceField:
    """Mirrors distance_field::PropagationDistanceField's constructor and point API (propagation_distance_field.cpp:48-58;
    propagate_negative_distances = the signed field)."""

    def __init__(self, size_x, size_y, size_z, resolution, origin_x, origin_y, origin_z, max_distance,
                 propagate_negative_distances=False):
        self._L = lib()
        self.args = (size_x, size_y, size_z, resolution, origin_x, origin_y, origin_z, max_distance)
        self.propagate_negative = bool(propagate_negative_distances)
        self.h = ctypes.c_void_p(self._L.orc_pdf_create_signed(*[float(a) for a in self.args], int(self.propagate_negative)))
        n = np.zeros(3, np.int32)
        m = np.zeros(1, np.int32)
        self._L.orc_pdf_dims(self.h, _ip(n), _ip(m))
        self.num_cells = tuple(int(v) for v in n)
        self.max_distance_sq = int(m[0])
        self.resolution = float(resolution)
        self.max_distance = float(max_distance)

    def __del__(self):
        try:
            self._L.orc_pdf_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def _pts(points):
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        return p, p.shape[0]

    def add_points(self, points):
        p, n = self._pts(points)
        self._L.orc_pdf_add_points(self.h, _dp(p), n)

    def add_voxels(self, ijk):
        v = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        self._L.orc_pdf_add_voxels(self.h, _ip(v), v.shape[0])

    def remove_points(self, points):
        p, n = self._pts(points)
        self._L.orc_pdf_remove_points(self.h, _dp(p), n)

    def update_points(self, old_points, new_points):
        a, na = self._pts(old_points)
        b, nb = self._pts(new_points)
        self._L.orc_pdf_update_points(self.h, _dp(a), na, _dp(b), nb)

    def reset(self):
        self._L.orc_pdf_reset(self.h)

    def d2(self):
        """int32 squared cell distances, shape (nx, ny, nz) -- index x*ny*nz + y*nz + z (voxel_grid.hpp:425)."""
        out = np.zeros(self.num_cells, np.int32)
        self._L.orc_pdf_get_d2(self.h, _ip(out))
        return out

    def neg_d2(self):
        """int32 negative_distance_square_ per voxel (0 outside obstacles; signed fields only)."""
        out = np.zeros(self.num_cells, np.int32)
        self._L.orc_pdf_get_neg_d2(self.h, _ip(out))
        return out

    def get_distance(self, x, y, z):
        return self._L.orc_pdf_get_distance(self.h, float(x), float(y), float(z))

    def get_cell_distance(self, x, y, z):
        return self._L.orc_pdf_get_cell_distance(self.h, int(x), int(y), int(z))

    def get_distance_gradient(self, x, y, z):
        g = np.zeros(3)
        inb = ctypes.c_int(0)
        d = self._L.orc_pdf_get_distance_gradient(self.h, float(x), float(y), float(z), _dp(g), ctypes.byref(inb))
        return d, g, bool(inb.value)

    def world_to_grid(self, x, y, z):
        c = np.zeros(3, np.int32)
        ok = self._L.orc_grid_world_to_grid(self._L.orc_pdf_grid(self.h), float(x), float(y), float(z), _ip(c))
        return tuple(int(v) for v in c), bool(ok)

    def grid_to_world(self, x, y, z):
        w = np.zeros(3)
        self._L.orc_grid_grid_to_world(self._L.orc_pdf_grid(self.h), int(x), int(y), int(z), _dp(w))
        return w


def read_df_occupancy(raw):
    """Parse a .df stream written by PropagationDistanceField::writeToStream (:635-673): 7 ASCII header lines,
    then zlib data with x-major, y, then z packed 8 voxels per byte, bit zi = LSB-first.
    Returns (header dict, occupied voxel index array [k,3] in the order readFromStream pushes them)."""
    hdr = {}
    pos = 0
    for key in ("resolution", "size_x", "size_y", "size_z", "origin_x", "origin_y", "origin_z"):
        end = raw.index(b"\n", pos)
        k, v = raw[pos:end].decode().split()
        assert k == key + ":", (k, key)
        hdr[key] = float(v)
        pos = end + 1
    data = zlib.decompress(raw[pos:])
    oo = 1.0 / hdr["resolution"]
    n = [int(hdr["size_x"] * oo), int(hdr["size_y"] * oo), int(hdr["size_z"] * oo)]
    zb = (n[2] + 7) // 8
    bits = np.frombuffer(data, np.uint8)[: n[0] * n[1] * zb].reshape(n[0], n[1], zb)
    occ = np.unpackbits(bits, axis=2, bitorder="little")[:, :, : n[2]]
    return hdr, n, np.argwhere(occ).astype(np.int32)


def write_df_occupancy(hdr, occ):
    """Inverse of read_df_occupancy (writeToStream :635-673). occ: bool (nx, ny, nz)."""
    head = "".join("%s: %s\n" % (k, repr(float(hdr[k])).rstrip("0").rstrip(".") if float(hdr[k]) != int(hdr[k])
                                 else str(int(hdr[k])))
                   for k in ("resolution", "size_x", "size_y", "size_z", "origin_x", "origin_y", "origin_z"))
    packed = np.packbits(occ.astype(np.uint8), axis=2, bitorder="little")
    return head.encode() + zlib.compress(packed.tobytes())
