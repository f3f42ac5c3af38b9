"""ctypes mirrors of the structs in include/asc.h (kept in one place; checked by tests/test_abi.py)."""
import ctypes as C

ASC_MAX_DOF = 20

ASC_OK, ASC_ERR_INVALID, ASC_ERR_CUDA, ASC_ERR_INDETERMINATE, ASC_ERR_COMM, ASC_ERR_NOT_CONVERGED = range(6)
ASC_BATCH_LM, ASC_BATCH_DOGLEG = 0, 1
ASC_POSE3_FIRST_ORDER_CAYLEY, ASC_POSE3_FIRST_ORDER_EXPMAP, ASC_POSE3_FULL_EXPMAP = 0, 1, 2


class AscParams(C.Structure):
    _fields_ = [
        ("sigma_px", C.c_double),
        ("sigma_landmark", C.c_double),
        ("cauchy_k", C.c_double),
        ("sigma_encoder", C.c_double * ASC_MAX_DOF),
        ("sigma_extrinsic", C.c_double * 6),
        ("use_dead_band", C.c_int32),
        ("dead_band", C.c_double),
        ("pose3_chart", C.c_int32),
        ("max_iterations", C.c_int32),
        ("relative_error_tol", C.c_double),
        ("absolute_error_tol", C.c_double),
        ("error_tol", C.c_double),
        ("lambda_initial", C.c_double),
        ("lambda_factor", C.c_double),
        ("lambda_upper_bound", C.c_double),
        ("lambda_lower_bound", C.c_double),
        ("min_model_fidelity", C.c_double),
        ("delta_initial", C.c_double),
        ("pcg_rel_tol", C.c_double),
        ("pcg_max_iters", C.c_int32),
        ("device", C.c_int32),
        ("pcg_cluster_max", C.c_int32),
        ("pcg_cluster_covis", C.c_double),
        ("pcg_coarse_levels", C.c_int32),
        ("pcg_ml_degree", C.c_int32),
        ("pcg_ml_agg", C.c_int32),
        ("pcg_ml_top_dim", C.c_int32),
    ]


class AscIterLog(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32),
        ("inner_steps", C.c_int32),
        ("accepted", C.c_int32),
        ("pcg_iterations", C.c_int32),
        ("error", C.c_double),
        ("lambda_or_delta", C.c_double),
        ("linearize_ms", C.c_double),
        ("solve_ms", C.c_double),
        ("prepare_ms", C.c_double),
        ("coarse_ms", C.c_double),
        ("pcg_ms", C.c_double),
    ]


class AscLinearStats(C.Structure):
    _fields_ = [
        ("error", C.c_double),
        ("linearize_ms", C.c_double),
        ("assemble_ms", C.c_double),
        ("cheirality", C.c_int64),
    ]


class AscSimConfig(C.Structure):
    _fields_ = [
        ("trajectory_size", C.c_int32),
        ("num_landmarks_x", C.c_int32),
        ("num_landmarks_y", C.c_int32),
        ("landmark_size_x", C.c_double),
        ("landmark_size_y", C.c_double),
        ("landmark_z_min", C.c_double),
        ("landmark_z_max", C.c_double),
        ("max_range", C.c_double),
        ("encoder_noise", C.c_double),
        ("velocity_noise", C.c_double),
        ("seed", C.c_int32),
        ("min_observations", C.c_int32),
        ("sim_extrinsic", C.c_double * 12),
        ("extrinsic_guess", C.c_double * 12),
        ("segment_length", C.c_int32),
    ]
