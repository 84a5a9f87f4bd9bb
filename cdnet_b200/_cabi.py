"""ctypes binding of libcdnet_b200.so (include/cdnet_b200.h).

There is no CPU fallback: if the shared library is missing and cannot be built with nvcc, or the
visible GPU is not a compute-capability-10.x part, importing the product path raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CDNET_B200_LIB: another build of the SAME library (tools/build_variant.py: compile-time tunings side by side on one box)
LIB_PATH = os.environ.get("CDNET_B200_LIB") or os.path.join(_HERE, "libcdnet_b200.so")

c_int, c_size_t, c_void_p = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p

# symbol -> (restype, argtypes); kept in the order of include/cdnet_b200.h
SIGNATURES = {
    "cdnet_version": (ctypes.c_char_p, []),
    "cdnet_device_ok": (c_int, [c_int]),
    "cdnet_ddm_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_ddm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_circshift": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cdnet_ccl_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_ccl": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_label_values": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_fill_holes_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_fill_holes": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_remove_small_mask_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_remove_small_mask": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_remove_small_labels_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_remove_small_labels": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_label_dilate": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cdnet_edt_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_edt": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_ws_postproc_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_shard_ws_relabel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cdnet_shard_ws_process": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_size_t, c_void_p]),
    "cdnet_ws_postproc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_size_t, c_void_p]),
    "cdnet_dam_postproc_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_dam_postproc": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_dcm_voting2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cdnet_plain_postproc_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_plain_postproc": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_center_points_workspace_bytes": (c_size_t, [c_int] * 4),
    "cdnet_center_points": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_encode_targets_workspace_bytes": (c_size_t, [c_int] * 3),
    "cdnet_label_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cdnet_encode_targets": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdnet_encode_targets_i32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdnet_label_stats_i32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cdnet_tta_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_int, c_void_p]),
    "cdnet_label_to_vector": (c_int, [c_void_p, c_int, c_void_p, c_int, c_size_t, c_int, c_void_p]),
    "cdnet_align_angle": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "cdnet_angle_to_vector": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "cdnet_vector_to_label": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "cdnet_direction_one_hot_workspace_bytes": (c_size_t, [c_int]),
    "cdnet_direction_one_hot": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_size_t,
                                        c_void_p, c_size_t, c_void_p]),
    "cdnet_ternary_label": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cdnet_shard_ddm_codes": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cdnet_shard_point_max": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdnet_shard_boost": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int, c_int, c_void_p]),
    "cdnet_shard_label_stage1": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "cdnet_shard_label_stage2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                         c_void_p]),
    "cdnet_shard_label_stage3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cdnet_shard_label_stage4": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                         c_void_p]),
    "cdnet_shard_relabel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cdnet_seam_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cdnet_seam_export": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_int, c_void_p]),
    "cdnet_seam_solve": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "cdnet_seam_ids_export": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "cdnet_seam_ids_apply": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "cdnet_label_pairs_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cdnet_label_pairs": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_int, c_void_p, c_size_t, c_void_p]),
    "cdnet_remap_labels": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_size_t, c_void_p]),
    "cdnet_launch_count": (ctypes.c_ulonglong, []),
    "cdnet_profile_enable": (None, [c_int]),
    "cdnet_profile_report": (c_int, [ctypes.c_char_p, c_size_t]),
}

E_BADARG, E_WORKSPACE = 1, 2
S_DDM_CONSTANT, S_WS_OVERFLOW, S_NO_BACKGROUND, S_CLASS_RANGE, S_SHARD_OVERFLOW = 1, 2, 16, 32, 64
S_WS_CONTESTED_SHIFT = 8  # status >> 8: watershed pixels two equal-priority age-0 markers compete for (cdnet_b200.h)

_lib = None


class CdnetError(RuntimeError):
    pass


def lib():
    """Loads (building first if the .so is absent and nvcc is available).  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        try:
            _build.build()
        except Exception as e:  # no fallback: the CUDA library IS the product
            raise CdnetError("libcdnet_b200.so is missing and could not be built (%s); run "
                             "`python -m cdnet_b200.build`" % (e,))
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise CdnetError("%s: CUDA error %d" % (what, -rc))
    raise CdnetError("%s: %s" % (what, {E_BADARG: "bad argument", E_WORKSPACE: "workspace too small"}.get(rc, rc)))
