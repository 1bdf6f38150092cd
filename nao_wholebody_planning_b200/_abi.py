"""ctypes view of include/nwb.h (the C ABI) and loader of the in-tree CUDA library.

The product path has no CPU fallback: if libnwb.so is missing or cannot be loaded the import of
any operator fails loudly with instructions to build it (`python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libnwb.so"

NJ = 21
NFRAMES = 29
ADVANCED, REACHED, TRAPPED = 0, 1, 2
REASON_COLLISION, REASON_UNSTABLE = 1, 2
QUIRK_DB_KEEPS_INFEASIBLE, QUIRK_SIGNED_FOOT_ERROR, QUIRK_DB_INDEX_FLOOR = 1, 2, 4
PAIRS_GROUP, PAIRS_ALL = 0, 1
F64, F32 = 0, 1


class NwbParams(C.Structure):
    """struct nwb_params (include/nwb.h)."""
    _fields_ = [
        ("device", C.c_int32),
        ("quirks", C.c_int32),
        ("srdf_trim", C.c_int32),
        ("scale_origin", C.c_int32),
        ("scale_sp", C.c_double),
        ("ds_tolerance", C.c_double),
        ("ik_eps", C.c_double),
        ("ik_maxiter", C.c_int32),
        ("scg_max_samples", C.c_int32),
        ("scg_max_ik_iterations", C.c_int32),
        ("planner_max_iterations", C.c_int32),
        ("planner_step_factor", C.c_double),
        ("pinv_eps", C.c_double),
        ("shortcut_waypoints", C.c_int32),
        ("shortcut_max_rounds", C.c_int32),
        ("r_arm_weight_approach", C.c_double),
        ("l_arm_weight_approach", C.c_double),
        ("legs_weight_approach", C.c_double),
        ("r_arm_weight_manipulate", C.c_double),
        ("l_arm_weight_manipulate", C.c_double),
        ("legs_weight_manipulate", C.c_double),
        ("manipulate_waypoints", C.c_int32),
        ("object_interp_points", C.c_int32),
        ("manipulation_velocity", C.c_double),
        ("time_parameterization", C.c_int32),
        ("iptp_max_iterations", C.c_int32),
        ("iptp_max_time_change", C.c_double),
        ("joint_max_velocity", C.c_double * NJ),
        ("joint_max_acceleration", C.c_double * NJ),
    ]


class NwbPlanResult(C.Structure):
    """struct nwb_plan_result (include/nwb.h)."""
    _fields_ = [(n, C.c_int32) for n in ("start_valid", "goal_valid", "success", "iterations", "num_nodes_generated",
                                         "nodes_start", "nodes_goal", "connector", "bridge_other", "raw_path_len")]


class GenDsRequest(C.Structure):
    _fields_ = [("max_samples", C.c_int32), ("max_ik_iterations", C.c_int32)]


class GenDsResponse(C.Structure):
    _fields_ = [("num_configs_generated", C.c_int32)]


class Point3(C.Structure):
    """struct nwb_point3: geometry_msgs/Point, geometry_msgs/Vector3."""
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


class Trajectory(C.Structure):
    """struct nwb_trajectory: moveit_msgs/DisplayTrajectory flattened into caller-provided arrays."""
    _fields_ = [("capacity", C.c_int32), ("n_points", C.c_int32), ("positions", C.POINTER(C.c_double)),
                ("time_from_start", C.POINTER(C.c_double))]


class GoalConfigRequest(C.Structure):      # Compute_Goal_Config.srv:1-4
    _fields_ = [("right_hand_position", Point3), ("right_hand_direction", Point3)]


class GoalConfigResponse(C.Structure):     # Compute_Goal_Config.srv:6-8
    _fields_ = [("wb_goal_config", C.c_double * NJ), ("num_goal_configs_found", C.c_int32)]


class MotionPlanRequest(C.Structure):      # Compute_Motion_Plan.srv:1-12
    _fields_ = [("right_hand_position", Point3), ("right_hand_direction", Point3), ("wb_start_config", C.POINTER(C.c_double)),
                ("wb_start_config_len", C.c_int32), ("execute_trajectory", C.c_uint8)]


class MotionPlanResponse(C.Structure):     # Compute_Motion_Plan.srv:14-29
    _fields_ = [("wb_goal_config", C.c_double * NJ), ("time_goal_config", C.c_double), ("motion_plan", Trajectory),
                ("num_nodes_generated", C.c_uint32), ("search_time", C.c_double), ("success", C.c_uint8)]


class ManipulationPlanResponse(C.Structure):   # Compute_Linear_Manipulation_Plan.srv:20-46 == Circular:27-53
    _fields_ = [("motion_plan_approach", Trajectory), ("motion_plan_manipulation", Trajectory),
                ("wb_goal_config_approach", C.c_double * NJ), ("time_goal_config_approach", C.c_double),
                ("wb_goal_config_manipulate", C.c_double * NJ), ("time_goal_config_manipulate", C.c_double),
                ("num_nodes_generated_approach", C.c_uint32), ("num_nodes_generated_manipulate", C.c_uint32),
                ("search_time_approach", C.c_double), ("search_time_manipulate", C.c_double),
                ("success_approach", C.c_uint8), ("success_manipulation", C.c_uint8)]


class LinearManipulationRequest(C.Structure):  # Compute_Linear_Manipulation_Plan.srv:1-18
    _fields_ = [("right_hand_position", Point3), ("right_hand_direction", Point3), ("wb_start_config", C.POINTER(C.c_double)),
                ("wb_start_config_len", C.c_int32), ("relative_angle", C.c_double), ("object_translation_length", C.c_double),
                ("execute_trajectory", C.c_uint8)]


class CircularManipulationRequest(C.Structure):  # Compute_Circular_Manipulation_Plan.srv:1-25
    _fields_ = [("right_hand_position", Point3), ("right_hand_direction", Point3), ("wb_start_config", C.POINTER(C.c_double)),
                ("wb_start_config_len", C.c_int32), ("relative_angle", C.c_double), ("rotation_radius", C.c_double),
                ("rotation_angle", C.c_double), ("clearance_handle", C.c_double), ("execute_trajectory", C.c_uint8)]


# NAO V4 URDF velocity limits in planner joint order (from memory, UNVERIFIED; see include/nwb.h)
NAO_MAX_VELOCITY = [4.16174, 6.40239, 6.40239, 6.40239, 4.16174, 8.26797, 7.19407, 8.26797, 7.19407, 24.6229,
                    8.26797, 7.19407, 8.26797, 7.19407, 24.6229, 4.16174, 4.16174, 6.40239, 6.40239, 6.40239, 4.16174]


def default_params(**over):
    """Reference defaults (planning_config/planning_parameters.yaml, hard-coded constants)."""
    p = NwbParams(device=0, quirks=QUIRK_SIGNED_FOOT_ERROR | QUIRK_DB_INDEX_FLOOR, srdf_trim=1, scale_origin=1,
                  scale_sp=0.8, ds_tolerance=0.01, ik_eps=1e-3, ik_maxiter=100, scg_max_samples=5000,
                  scg_max_ik_iterations=5, planner_max_iterations=8000, planner_step_factor=0.1, pinv_eps=1e-5,
                  shortcut_waypoints=100, shortcut_max_rounds=500, r_arm_weight_approach=1.0, l_arm_weight_approach=1.0,
                  legs_weight_approach=1.0, r_arm_weight_manipulate=1.0, l_arm_weight_manipulate=0.0,
                  legs_weight_manipulate=1.0, manipulate_waypoints=4, object_interp_points=20, manipulation_velocity=0.02,
                  time_parameterization=1, iptp_max_iterations=100, iptp_max_time_change=0.01)
    p.joint_max_velocity[:] = NAO_MAX_VELOCITY
    p.joint_max_acceleration[:] = [1.0] * NJ
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class NwbError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """Load libnwb.so (built in-tree by __graft_entry__.build()).  Never falls back to a CPU path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path or os.environ.get("NWB_LIB", LIB_PATH))
    if not p.exists():
        raise NwbError(
            f"{p} not found: the CUDA library is not built. Run `python -c \"import __graft_entry__ as g; g.build()\"` "
            "from the repo root. There is no CPU fallback.")
    lib = C.CDLL(str(p))
    _declare(lib)
    if path is None:
        _lib = lib
    return lib


def _declare(lib):
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pd, pu8, pf, pi32 = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32)
    sigs = {
        "nwb_params_default": (C.c_int, [C.POINTER(NwbParams)]),
        "nwb_abi_version": (C.c_int, []),
        "nwb_ctx_create": (C.c_int, [C.POINTER(NwbParams), C.POINTER(vp)]),
        "nwb_ctx_destroy": (None, [vp]),
        "nwb_last_error": (C.c_char_p, [vp]),
        "nwb_ctx_synchronize": (C.c_int, [vp]),
        "nwb_scale_support_polygon": (C.c_int, [vp, dbl]),
        "nwb_get_joint_limits": (C.c_int, [vp, pd, pd]),
        "nwb_ctx_stream": (vp, [vp]),
        "nwb_ctx_set_stream": (C.c_int, [vp, vp]),
        "nwb_scene_clear": (C.c_int, [vp]),
        "nwb_scene_add_box": (C.c_int, [vp, pd, pd, pd]),
        "nwb_scene_add_boxes": (C.c_int, [vp, vp, i64]),
        "nwb_scene_num_boxes": (C.c_int, [vp]),
        "nwb_is_feasible_batch": (C.c_int, [vp, vp, i64, vp, vp, vp]),
        "nwb_is_feasible_batch_f32": (C.c_int, [vp, vp, i64, vp, vp, vp]),
        "nwb_is_feasible_batch_dev": (C.c_int, [vp, vp, C.c_int, i64, vp, vp, vp]),
        "nwb_is_state_colliding_batch": (C.c_int, [vp, vp, i64, vp, vp]),
        "nwb_fk_batch": (C.c_int, [vp, vp, i64, vp, vp]),
        "nwb_frame_name": (C.c_char_p, [C.c_int]),
        "nwb_tree_create": (C.c_int, [vp, i64, C.POINTER(vp)]),
        "nwb_tree_destroy": (None, [vp]),
        "nwb_tree_clear": (C.c_int, [vp]),
        "nwb_tree_size": (i64, [vp]),
        "nwb_tree_add": (C.c_int, [vp, vp, vp, i64]),
        "nwb_tree_add_dev": (C.c_int, [vp, vp, vp, i64]),
        "nwb_tree_get": (C.c_int, [vp, i64, i64, vp, vp]),
        "nwb_nearest_neighbour_batch": (C.c_int, [vp, vp, i64, vp, vp]),
        "nwb_nearest_neighbour_batch_dev": (C.c_int, [vp, vp, i64, vp, vp]),
        "nwb_nn_scan_owner": (C.c_int64, [vp, i64, i64]),
        "nwb_nn_debug_tensor_tile": (C.c_int, [vp, vp, i64, vp]),
        "nwb_steer_project_batch": (C.c_int, [vp, vp, vp, i64, vp, dbl, vp, vp, vp]),
        "nwb_steer_project_batch_dev": (C.c_int, [vp, vp, vp, i64, vp, dbl, vp, vp, vp]),
        "nwb_connect_batch": (C.c_int, [vp, vp, vp, i64, vp, dbl, i32, vp, vp, vp]),
        "nwb_is_path_valid_batch": (C.c_int, [vp, vp, vp, i64, i32, vp, vp]),
        "nwb_is_path_valid_batch_dev": (C.c_int, [vp, vp, vp, i64, i32, vp, vp]),
        "nwb_sample_stable_configs": (C.c_int, [vp, C.c_uint64, C.c_uint64, i64, i32, vp, i64, C.POINTER(i64), vp, vp]),
        "nwb_sample_stable_configs_dev": (C.c_int, [vp, C.c_uint64, C.c_uint64, i64, i32, vp, vp, i64, vp, vp]),
        "nwb_solve_query_batch": (C.c_int, [vp, vp, vp, i64, vp, i32, vp, vp, i32, dbl, i32, vp, vp, i32]),
        "nwb_shortcut_path": (C.c_int, [vp, vp, i32, vp, i32, vp]),
        "nwb_interpolate_path": (C.c_int, [vp, i32, i32, vp, i32]),
        "nwb_dev_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "nwb_dev_free": (C.c_int, [vp, vp]),
        "nwb_ipc_export": (C.c_int, [vp, vp, vp]),
        "nwb_ipc_import": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "nwb_ipc_close": (C.c_int, [vp, vp]),
        "nwb_set_gather_peers": (C.c_int, [vp, C.POINTER(vp), i32, i64]),
        "nwb_write_configs": (C.c_int, [C.c_char_p, vp, i64]),
        "nwb_load_ds_database": (C.c_int, [C.c_char_p, vp, i64, C.POINTER(i64)]),
        "nwb_generate_ds_database": (C.c_int, [vp, C.POINTER(GenDsRequest), C.POINTER(GenDsResponse), C.c_char_p, C.c_uint64]),
        "nwb_arm_ik_batch": (C.c_int, [vp, vp, i64, vp, vp]),
        "nwb_goal_arm_ik_batch": (C.c_int, [vp, vp, i64, vp, C.c_uint64, C.c_uint64, vp, vp, vp]),
        "nwb_hand_trajectory": (C.c_int, [i32, vp, vp, i32, vp, vp]),
        "nwb_enforce_ma_batch": (C.c_int, [vp, vp, i32, i32, vp, vp, vp, i64, vp, vp]),
        "nwb_goal_sample_batch": (C.c_int, [vp, C.c_uint64, C.c_uint64, i64, vp, i32, vp, vp, vp]),
        "nwb_sample_goal_configs": (C.c_int, [vp, C.c_uint64, vp, i32, i32, vp, vp, vp, C.POINTER(i32), C.POINTER(i64)]),
        "nwb_solve_query_constrained_batch": (C.c_int, [vp, vp, vp, i64, vp, i32, vp, vp, i32, dbl, i32, vp, i32, vp, vp, i32]),
        "nwb_ctx_set_ds_database": (C.c_int, [vp, vp, i64]),
        "nwb_ctx_load_ds_database": (C.c_int, [vp, C.c_char_p]),
        "nwb_ctx_ds_database_size": (i64, [vp]),
        "nwb_time_parameterize": (C.c_int, [vp, vp, i32, vp]),
        "nwb_compute_goal_config": (C.c_int, [vp, C.POINTER(GoalConfigRequest), C.POINTER(GoalConfigResponse), C.c_uint64]),
        "nwb_compute_motion_plan": (C.c_int, [vp, C.POINTER(MotionPlanRequest), C.POINTER(MotionPlanResponse), C.c_uint64]),
        "nwb_compute_motion_plan_batch": (C.c_int, [vp, C.POINTER(MotionPlanRequest), C.POINTER(MotionPlanResponse), i64, vp]),
        "nwb_compute_linear_manipulation_plan": (C.c_int, [vp, C.POINTER(LinearManipulationRequest),
                                                           C.POINTER(ManipulationPlanResponse), C.c_uint64]),
        "nwb_compute_circular_manipulation_plan": (C.c_int, [vp, C.POINTER(CircularManipulationRequest),
                                                             C.POINTER(ManipulationPlanResponse), C.c_uint64]),
        "nwb_host_alloc": (vp, [C.c_size_t]),
        "nwb_host_free": (None, [vp]),
        "nwb_ffma_peak": (C.c_int, [vp, C.c_int, pd, pd]),
        "nwb_last_kernel_ms": (C.c_int, [vp, pd, C.POINTER(i64)]),
        "nwb_set_kernel_timing": (C.c_int, [vp, C.c_int]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    lib._nwb_sigs = sigs
