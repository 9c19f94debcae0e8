"""ORACLE package -- test infrastructure, NOT product code.

CPU restatement of the reference's distance-field state-validity path (see oracle/csrc/oracle.c and
oracle/robot_model.py for the file:line citations).  Allowed importers: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline and --impl reference legs.  moveit2_b200/ must never import this package.

Parity status: PINNED on the reference's own golden data (tests/test_oracle_df.py, tests/test_oracle_fk.py):
test_small.df / test_big.df occupancy, TestAddRemovePoints / removal == fresh build properties, VoxelGrid
sizing, OneRobot FK poses and mimic joints.  NOT pinned (no golden data exists in the reference tree):
sphere decompositions of mesh/box/cylinder links (un-vendored geometric_shapes) and the PR2 booleans of
test_collision_distance_field.cpp (PR2 description is external).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "oracle.c")
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    """gcc recipe for the C restatement (outputs only under oracle/_build/)."""
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    # x86-64-v3 (AVX2) keeps a library built in the authoring container runnable on the GPU box's host CPU;
    # set ORACLE_MARCH=native for baseline timing on the box that runs it (bench.py does).
    march = os.environ.get("ORACLE_MARCH", "x86-64-v3")
    cmd = ["gcc", "-O3", "-march=" + march, "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-Wall",
           "-o", _LIB, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        path = _LIB if os.path.exists(_LIB) and not os.access(_SRC, os.R_OK) else build()
        _lib = ctypes.CDLL(path)
        _declare(_lib)
    return _lib


def _declare(L):
    c = ctypes
    dp, ip, vp = c.POINTER(c.c_double), c.POINTER(c.c_int32), c.c_void_p
    L.orc_pdf_create.restype = vp
    L.orc_pdf_create.argtypes = [c.c_double] * 8
    L.orc_pdf_create_signed.restype = vp
    L.orc_pdf_create_signed.argtypes = [c.c_double] * 8 + [c.c_int]
    L.orc_pdf_get_neg_d2.argtypes = [vp, ip]
    L.orc_pdf_destroy.argtypes = [vp]
    L.orc_pdf_reset.argtypes = [vp]
    L.orc_pdf_add_points.argtypes = [vp, dp, c.c_size_t]
    L.orc_pdf_add_voxels.argtypes = [vp, ip, c.c_size_t]
    L.orc_pdf_remove_points.argtypes = [vp, dp, c.c_size_t]
    L.orc_pdf_update_points.argtypes = [vp, dp, c.c_size_t, dp, c.c_size_t]
    L.orc_pdf_dims.argtypes = [vp, ip, ip]
    L.orc_pdf_get_d2.argtypes = [vp, ip]
    L.orc_pdf_get_cell_distance.restype = c.c_double
    L.orc_pdf_get_cell_distance.argtypes = [vp, c.c_int, c.c_int, c.c_int]
    L.orc_pdf_get_distance.restype = c.c_double
    L.orc_pdf_get_distance.argtypes = [vp, c.c_double, c.c_double, c.c_double]
    L.orc_pdf_get_distance_gradient.restype = c.c_double
    L.orc_pdf_get_distance_gradient.argtypes = [vp, c.c_double, c.c_double, c.c_double, dp, c.POINTER(c.c_int)]
    L.orc_pdf_grid.restype = vp
    L.orc_pdf_grid.argtypes = [vp]
    L.orc_grid_world_to_grid.restype = c.c_int
    L.orc_grid_world_to_grid.argtypes = [vp, c.c_double, c.c_double, c.c_double, ip]
    L.orc_grid_grid_to_world.argtypes = [vp, c.c_int, c.c_int, c.c_int, dp]
    L.orc_fk_batch.argtypes = [vp, dp, dp, c.c_size_t, dp]
    L.orc_check_batch.restype = c.c_int
    L.orc_check_batch.argtypes = [vp, vp, vp, vp, dp, dp, c.c_size_t, c.c_int, c.c_int, c.c_int,
                                  c.POINTER(c.c_uint8)]
    L.orc_state_spheres.argtypes = [vp, vp, dp, dp, dp]
    L.orc_state_margin.restype = c.c_double
    L.orc_state_margin.argtypes = [vp, vp, vp, vp, dp, dp, c.c_int]
    L.orc_collision_gradients.argtypes = [vp, vp, vp, vp, dp, dp, c.c_size_t, dp, dp, c.POINTER(c.c_uint8), dp]
    L.orc_collision_gradients.restype = None
    L.orc_collision_gradients_lf.argtypes = [vp, vp, vp, vp, dp, dp, c.c_size_t, dp, dp, c.POINTER(c.c_uint8), dp,
                                             c.POINTER(vp), c.POINTER(c.c_uint8)]
    L.orc_collision_gradients_lf.restype = None
    L.orc_check_contacts.argtypes = [vp, vp, vp, vp, dp, dp, c.c_size_t, c.c_int, c.c_uint, c.c_uint,
                                     c.POINTER(c.c_uint8), ip, vp]
    L.orc_check_contacts.restype = None
    L.orc_num_threads.restype = c.c_int
