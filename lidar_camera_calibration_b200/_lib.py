"""Loads libilcc2_b200.so (the CUDA product library) through ctypes and declares the C ABI of include/ilcc2_b200.h.

There is no fallback: if the library is missing the import fails with the build command, and every compute
entry point of the library itself fails when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ILCC2_LIB overrides the library path (A/B runs of build variants); the default is the in-tree build.
LIB_PATH = os.environ.get("ILCC2_LIB") or os.path.join(_HERE, "libilcc2_b200.so")

EXPORTS = [
    "ilcc2_version", "ilcc2_default_params", "ilcc2_set_chessboard_param", "ilcc2_create", "ilcc2_destroy",
    "ilcc2_last_error", "ilcc2_extract_corners", "ilcc2_fetch_indices", "ilcc2_fetch_board_points",
    "ilcc2_set_profiling", "ilcc2_stage_times", "ilcc2_launch_count", "ilcc2_save_corners2txt", "ilcc2_shard_range",
    "ilcc2_pnp_create", "ilcc2_pnp_destroy", "ilcc2_pnp_last_error", "ilcc2_pnp_launch_count", "ilcc2_axis_roughly",
    "ilcc2_pnp_order_corners", "ilcc2_pnp_solve", "ilcc2_compose_extrinsic", "ilcc2_extrinsic2bin",
    "ilcc2_read_cam_corners", "ilcc2_read_lidar_corners", "ilcc2_read_camera_yaml", "ilcc2_project_points",
    "ilcc2_set_gather_targets", "ilcc2_create_sized", "ilcc2_gather_setup_nccl", "ilcc2_gather_sync",
    "ilcc2_gather_is_fused", "ilcc2_gather_teardown", "ilcc2_set_lanes",
]


class Params(C.Structure):
    _fields_ = [
        ("roi_half_x", C.c_float), ("roi_half_y", C.c_float), ("roi_half_z", C.c_float),
        ("cluster_tolerance", C.c_float),
        ("min_cluster_size", C.c_int32), ("max_cluster_size", C.c_int32),
        ("ransac_threshold", C.c_double),
        ("ransac_max_iterations", C.c_int32),
        ("ransac_probability", C.c_double),
        ("hist_bins", C.c_int32),
        ("gray_rate", C.c_double), ("huber_delta", C.c_double), ("grid_length", C.c_double),
        ("cells_x", C.c_int32), ("cells_y", C.c_int32),
        ("min_board_points", C.c_int32), ("max_solver_iterations", C.c_int32),
    ]


class Diag(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_roi", C.c_int32), ("nn_index", C.c_int32), ("n_clusters", C.c_int32),
        ("click_in_cluster", C.c_int32), ("cluster_rank", C.c_int32), ("n_cluster_pts", C.c_int32),
        ("ransac_iterations", C.c_int32), ("n_ransac_inliers", C.c_int32), ("n_board", C.c_int32),
        ("n_opt", C.c_int32), ("hyp_chosen", C.c_int32),
        ("iters", (C.c_int32 * 2) * 2), ("term", (C.c_int32 * 2) * 2), ("n_evals", C.c_int32 * 2),
        ("plane", C.c_float * 4), ("pca", C.c_float * 16), ("eigvals", C.c_float * 3),
        ("hist_min", C.c_double), ("hist_max", C.c_double), ("hist_mean", C.c_double),
        ("r_low", C.c_double), ("r_high", C.c_double), ("gray", C.c_double * 2),
        ("theta_t", (C.c_double * 3) * 2), ("cost", C.c_double * 2), ("cost_a", C.c_double * 2),
    ]


class PnpSummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("iterations", C.c_int32),
                ("termination", C.c_int32), ("n_evals", C.c_int32), ("n_pairs", C.c_int32)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C lidar_camera_calibration_b200/csrc` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.ilcc2_version.restype = C.c_int
    L.ilcc2_default_params.argtypes = [C.POINTER(Params)]
    L.ilcc2_set_chessboard_param.restype = C.c_int
    L.ilcc2_set_chessboard_param.argtypes = [C.POINTER(Params), C.c_char_p]
    L.ilcc2_create.restype = C.c_int
    L.ilcc2_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.ilcc2_create_sized.restype = C.c_int
    L.ilcc2_create_sized.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.ilcc2_destroy.restype = C.c_int
    L.ilcc2_destroy.argtypes = [C.c_void_p]
    L.ilcc2_last_error.restype = C.c_char_p
    L.ilcc2_last_error.argtypes = [C.c_void_p]
    L.ilcc2_extract_corners.restype = C.c_int
    L.ilcc2_extract_corners.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ilcc2_fetch_indices.restype = C.c_int
    L.ilcc2_fetch_indices.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.ilcc2_fetch_board_points.restype = C.c_int
    L.ilcc2_fetch_board_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.ilcc2_set_lanes.restype = C.c_int
    L.ilcc2_set_lanes.argtypes = [C.c_void_p, C.c_int]
    L.ilcc2_set_profiling.restype = C.c_int
    L.ilcc2_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.ilcc2_stage_times.restype = C.c_int
    L.ilcc2_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.c_int]
    L.ilcc2_launch_count.restype = C.c_int64
    L.ilcc2_launch_count.argtypes = [C.c_void_p]
    L.ilcc2_save_corners2txt.restype = C.c_int
    L.ilcc2_save_corners2txt.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.ilcc2_shard_range.restype = None
    L.ilcc2_shard_range.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ilcc2_set_gather_targets.restype = C.c_int
    L.ilcc2_set_gather_targets.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int]
    L.ilcc2_gather_setup_nccl.restype = C.c_int
    L.ilcc2_gather_setup_nccl.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.ilcc2_gather_sync.restype = C.c_int
    L.ilcc2_gather_sync.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
    L.ilcc2_gather_is_fused.restype = C.c_int
    L.ilcc2_gather_is_fused.argtypes = [C.c_void_p]
    L.ilcc2_gather_teardown.restype = C.c_int
    L.ilcc2_gather_teardown.argtypes = [C.c_void_p]
    L.ilcc2_pnp_create.restype = C.c_int
    L.ilcc2_pnp_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.ilcc2_pnp_destroy.restype = C.c_int
    L.ilcc2_pnp_destroy.argtypes = [C.c_void_p]
    L.ilcc2_pnp_last_error.restype = C.c_char_p
    L.ilcc2_pnp_last_error.argtypes = [C.c_void_p]
    L.ilcc2_pnp_launch_count.restype = C.c_int64
    L.ilcc2_pnp_launch_count.argtypes = [C.c_void_p]
    L.ilcc2_axis_roughly.restype = C.c_int
    L.ilcc2_axis_roughly.argtypes = [C.c_char_p, C.c_void_p]
    L.ilcc2_pnp_order_corners.restype = C.c_int
    L.ilcc2_pnp_order_corners.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ilcc2_pnp_solve.restype = C.c_int
    L.ilcc2_pnp_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double,
                                  C.c_int, C.c_void_p, C.c_void_p, C.POINTER(PnpSummary), C.c_int, C.c_void_p]
    L.ilcc2_project_points.restype = C.c_int
    L.ilcc2_project_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ilcc2_compose_extrinsic.restype = C.c_int
    L.ilcc2_compose_extrinsic.argtypes = [C.c_void_p] * 4
    L.ilcc2_extrinsic2bin.restype = C.c_int
    L.ilcc2_extrinsic2bin.argtypes = [C.c_char_p, C.c_void_p]
    L.ilcc2_read_cam_corners.restype = C.c_int
    L.ilcc2_read_cam_corners.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.ilcc2_read_lidar_corners.restype = C.c_int
    L.ilcc2_read_lidar_corners.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    L.ilcc2_read_camera_yaml.restype = C.c_int
    L.ilcc2_read_camera_yaml.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = L
    return L
