"""ctypes binding of include/b200sv.h.  Loading fails loudly when libb200sv.so has not been built: there is no
Python or CPU implementation of the path to fall back to."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SV_LIB", os.path.join(HERE, "libb200sv.so"))  # override: experimental variants

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_PARSE = 0, -1, -2, -3, -4, -5
DEVICE_NONE = -1
FIELD_WORLD, FIELD_SELF = 0, 1
CHECK_WORLD, CHECK_SELF, CHECK_INTRA, CHECK_ALL, CHECK_FP64 = 1, 2, 4, 7, 256


class Robot(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("n_vars", C.c_int32), ("parent", C.POINTER(C.c_int32)),
                ("joint_type", C.POINTER(C.c_int32)), ("joint_axis", C.POINTER(C.c_double)),
                ("joint_origin", C.POINTER(C.c_double)), ("first_var", C.POINTER(C.c_int32)),
                ("base_positions", C.POINTER(C.c_double)), ("n_group_vars", C.c_int32),
                ("group_var_index", C.POINTER(C.c_int32)), ("n_mimic", C.c_int32),
                ("mimic_dst", C.POINTER(C.c_int32)), ("mimic_src", C.POINTER(C.c_int32)),
                ("mimic_factor", C.POINTER(C.c_double)), ("mimic_offset", C.POINTER(C.c_double))]


class CollisionModel(C.Structure):
    _fields_ = [("n_glinks", C.c_int32), ("glink_index", C.POINTER(C.c_int32)),
                ("self_enabled", C.POINTER(C.c_uint8)), ("intra_enabled", C.POINTER(C.c_uint8)),
                ("sphere_start", C.POINTER(C.c_int32)), ("sphere_xyzr", C.POINTER(C.c_double)),
                ("bounding_sphere", C.POINTER(C.c_double)), ("body_id", C.POINTER(C.c_int32))]


class FieldCfg(C.Structure):
    _fields_ = [("size", C.c_double * 3), ("origin", C.c_double * 3), ("resolution", C.c_double),
                ("max_propagation_distance", C.c_double), ("collision_tolerance", C.c_double),
                ("use_signed_distance_field", C.c_int32)]


class Shape(C.Structure):
    _fields_ = [("type", C.c_int32), ("lattice_frame", C.c_int32), ("dims", C.c_double * 3), ("pose", C.c_double * 12)]


SHAPE_SPHERE, SHAPE_CYLINDER, SHAPE_BOX = 1, 2, 4
LATTICE_WORLD, LATTICE_BODY = 0, 1
PROPAGATION_EXACT, PROPAGATION_REFERENCE = 0, 1


class Convex(C.Structure):
    _fields_ = [("n_planes", C.c_int32), ("lattice_frame", C.c_int32), ("planes", C.POINTER(C.c_double)),
                ("bounding_sphere", C.c_double * 4), ("pose", C.c_double * 12)]


class Info(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("n_vars", C.c_int32), ("n_group_vars", C.c_int32),
                ("n_glinks", C.c_int32), ("n_spheres", C.c_int32), ("n_link_pairs", C.c_int32),
                ("num_cells", C.c_int32 * 3), ("max_distance_sq", C.c_int32), ("device", C.c_int32),
                ("field_bytes_per_voxel", C.c_int32), ("sm_count", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_states", C.c_uint64), ("n_valid", C.c_uint64), ("n_rechecked", C.c_uint64),
                ("check_ms", C.c_float), ("edt_ms", C.c_float)]


class PropagationStats(C.Structure):
    _fields_ = [("rounds", C.c_uint64), ("hazard_rounds", C.c_uint64), ("queue_entries", C.c_uint64),
                ("backward_offers", C.c_uint64), ("queued_for_next_call", C.c_uint64)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libb200sv.so is not built (%s missing): run `python -m moveit2_b200.build`. "
                           "There is no CPU fallback for this path." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, dp, ip, u8p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    sig = {
        "b200sv_create": (C.c_int, [C.c_int, C.POINTER(Robot), C.POINTER(CollisionModel), C.POINTER(FieldCfg),
                                    C.POINTER(vp)]),
        "b200sv_create_from_urdf": (C.c_int, [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(FieldCfg),
                                              C.POINTER(vp)]),
        "b200sv_destroy": (None, [vp]),
        "b200sv_last_error": (C.c_char_p, [vp]),
        "b200sv_get_info": (C.c_int, [vp, C.POINTER(Info)]),
        "b200sv_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "b200sv_get_robot_arrays": (C.c_int, [vp, ip, ip, dp, dp, ip, dp, ip]),
        "b200sv_get_model_arrays": (C.c_int, [vp, ip, u8p, u8p, ip, dp, dp]),
        "b200sv_get_link_names": (C.c_int, [vp, C.c_char_p, C.c_size_t]),
        "b200sv_get_group_bounds": (C.c_int, [vp, dp, dp]),
        "b200sv_get_sphere_thresholds": (C.c_int, [vp, ip]),
        "b200sv_field_reset": (C.c_int, [vp, C.c_int]),
        "b200sv_field_set_propagation": (C.c_int, [vp, C.c_int, C.c_int]),
        "b200sv_field_propagation_stats": (C.c_int, [vp, C.c_int, C.POINTER(PropagationStats)]),
        "b200sv_field_begin_update": (C.c_int, [vp, C.c_int]),
        "b200sv_field_end_update": (C.c_int, [vp, C.c_int]),
        "b200sv_field_add_points": (C.c_int, [vp, C.c_int, dp, C.c_size_t]),
        "b200sv_field_remove_points": (C.c_int, [vp, C.c_int, dp, C.c_size_t]),
        "b200sv_field_add_sphere": (C.c_int, [vp, C.c_int, dp, C.c_double]),
        "b200sv_field_add_shape": (C.c_int, [vp, C.c_int, C.POINTER(Shape), C.c_int]),
        "b200sv_field_add_convex": (C.c_int, [vp, C.c_int, C.POINTER(Convex), C.c_int]),
        "b200sv_field_add_octree_leaves": (C.c_int, [vp, C.c_int, dp, C.c_size_t]),
        "b200sv_field_update_points": (C.c_int, [vp, C.c_int, dp, C.c_size_t, dp, C.c_size_t]),
        "b200sv_field_add_points_device": (C.c_int, [vp, C.c_int, vp, C.c_size_t, vp]),
        "b200sv_field_get_d2": (C.c_int, [vp, C.c_int, ip]),
        "b200sv_field_set_d2": (C.c_int, [vp, C.c_int, ip]),
        "b200sv_field_get_neg_d2": (C.c_int, [vp, C.c_int, ip]),
        "b200sv_field_set_neg_d2": (C.c_int, [vp, C.c_int, ip]),
        "b200sv_field_write_df": (C.c_int, [vp, C.c_int, u8p, C.c_size_t, C.POINTER(C.c_size_t)]),
        "b200sv_field_read_df": (C.c_int, [vp, C.c_int, u8p, C.c_size_t]),
        "b200sv_get_self_points": (C.c_int, [vp, dp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "b200sv_fk_batch": (C.c_int, [vp, dp, C.c_size_t, C.c_int, dp]),
        "b200sv_check_batch": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, vp]),
        "b200sv_check_batch_device": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, vp, vp]),
        "b200sv_sphere_centers": (C.c_int, [vp, dp, dp]),
        "b200sv_collision_gradients": (C.c_int, [vp, dp, C.c_size_t, dp, dp, u8p, dp, dp]),
        "b200sv_check_contacts": (C.c_int, [vp, dp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, u8p, ip, vp]),
        "b200sv_check_motions": (C.c_int, [vp, dp, dp, C.c_size_t, C.c_double, u8p, dp, C.c_uint32, C.c_uint32, u8p,
                                           ip, C.POINTER(C.c_uint64)]),
        "b200sv_check_batch_f32": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, vp]),
        "b200sv_fk_batch_device": (C.c_int, [vp, vp, C.c_size_t, C.c_int, vp, vp]),
        "b200sv_set_acm_masks": (C.c_int, [vp, u8p, u8p]),
        "b200sv_set_base_state": (C.c_int, [vp, dp, C.c_int]),
        "b200sv_set_self_link_points": (C.c_int, [vp, C.c_int32, dp, C.c_size_t]),
        "b200sv_field_export_size": (C.c_int, [vp, C.POINTER(C.c_size_t)]),
        "b200sv_field_export_device": (C.c_int, [vp, C.c_int, vp, vp]),
        "b200sv_field_import_device": (C.c_int, [vp, C.c_int, vp, vp]),
        "b200sv_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(Robot), C.POINTER(CollisionModel),
                                          C.POINTER(FieldCfg), C.POINTER(vp)]),
        "b200sv_create_multi_from_urdf": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_char_p, C.c_char_p,
                                                    C.POINTER(FieldCfg), C.POINTER(vp)]),
        "b200sv_destroy_multi": (None, [vp]),
        "b200sv_multi_last_error": (C.c_char_p, [vp]),
        "b200sv_multi_device_count": (C.c_int, [vp]),
        "b200sv_multi_ctx": (vp, [vp, C.c_int]),
        "b200sv_multi_field_sync": (C.c_int, [vp, C.c_int]),
        "b200sv_multi_field_reset": (C.c_int, [vp, C.c_int]),
        "b200sv_multi_field_set_propagation": (C.c_int, [vp, C.c_int, C.c_int]),
        "b200sv_multi_field_add_points": (C.c_int, [vp, C.c_int, dp, C.c_size_t]),
        "b200sv_check_batch_multi": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, vp]),
        "b200sv_multi_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "b200sv_set_link_field": (C.c_int, [vp, C.c_int32, dp, dp, ip]),
        "b200sv_build_link_fields": (C.c_int, [vp]),
        "b200sv_get_link_field": (C.c_int, [vp, C.c_int32, ip, dp, ip]),
        "b200sv_set_gradient_link_mask": (C.c_int, [vp, u8p]),
        "b200sv_get_fp32_budget": (C.c_int, [vp, dp, dp]),
        "b200sv_debug_fast_centres": (C.c_int, [vp, dp, C.c_size_t, dp]),
        "b200sv_gather_roofline": (C.c_int, [vp, C.c_size_t, C.c_int, dp]),
        "b200sv_set_tuning": (C.c_int, [vp, C.c_double, C.c_int, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here means the library and the header disagree: fail loudly
        fn.restype = res
        fn.argtypes = args
    L._b200sv_symbols = sorted(sig)
    _lib = L
    return L
