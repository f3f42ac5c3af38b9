/*
 * asc.h -- C ABI of the B200-native batch optimiser for arm_slam_calib graphs.
 *
 * This is the drop-in boundary.  Everything the reference does
 * between "here is the factor graph and the current estimate" and "here is the new
 * estimate" -- ArmSlamCalib::OptimizeStep(), src/ArmSlamCalib.cpp:429-533, i.e. the
 * gtsam::LevenbergMarquardtOptimizer / gtsam::DoglegOptimizer calls at :443-449 and
 * :460-465 -- is reachable through the entry points below.  The header-only facade in
 * include/gtsam_facade/ packs a GTSAM-shaped graph into these calls.
 *
 * Conventions
 *   - plain C, caller-owned host buffers, opaque handle, int status codes, no exceptions
 *   - all floating point is IEEE double (GTSAM Vector/Matrix are double)
 *   - rigid transforms are 12 doubles, row-major 3x4 [R | t]
 *   - Pose3 tangent order is [omega(3), v(3)] (GTSAM convention)
 *   - indices are int32, 0-based
 *   - there is NO CPU fallback: every compute entry point fails with ASC_ERR_CUDA when
 *     no sm_100 device is usable
 */
#ifndef ASC_H_
#define ASC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASC_MAX_DOF 20 /* array bound of asc_params; kernels are instantiated for 2..12 DOF (asc_set_chain rejects the rest with ASC_ERR_INVALID); data/dof_experiments.py:17 sweeps 4..19 */

typedef struct asc_handle asc_handle;

enum asc_status {
    ASC_OK = 0,
    ASC_ERR_INVALID = 1,       /* bad argument / call order */
    ASC_ERR_CUDA = 2,          /* no device, launch or allocation failure */
    ASC_ERR_INDETERMINATE = 3, /* gtsam::IndeterminantLinearSystemException, ArmSlamCalib.cpp:451 */
    ASC_ERR_COMM = 4,          /* NCCL failure */
    ASC_ERR_NOT_CONVERGED = 5  /* the PCG solve hit pcg_max_iters above pcg_rel_tol (no GTSAM analogue: Cholesky is direct) */
};

/* optimization_mode, include/arm_slam_calib/ArmSlamCalib.h:59-64 (ISAM is out of scope) */
enum asc_mode { ASC_BATCH_LM = 0, ASC_BATCH_DOGLEG = 1 };

/* [DEP] GTSAM build flags that change Pose3 retract/localCoordinates */
enum asc_pose3_chart {
    ASC_POSE3_FIRST_ORDER_CAYLEY = 0, /* GTSAM 3.2 default build */
    ASC_POSE3_FIRST_ORDER_EXPMAP = 1, /* GTSAM_ROT3_EXPMAP */
    ASC_POSE3_FULL_EXPMAP = 2         /* GTSAM_POSE3_EXPMAP */
};

/* Noise / robust-loss parameters; defaults = ArmSlamCalib::Params(), ArmSlamCalib.h:69-105 */
typedef struct asc_params {
    double sigma_px;                 /* projectionNoiseLevel, ArmSlamCalib.h:86 (1.0) */
    double sigma_landmark;           /* landmarkNoiseLevel, ArmSlamCalib.h:83 (50.0) */
    double cauchy_k;                 /* cauchy_multiplier, ArmSlamCalib.cpp:197-198 (3.0) */
    double sigma_encoder[ASC_MAX_DOF]; /* encoderNoiseLevel per joint, ArmSlamCalib.cpp:672-677 (0.1) */
    double sigma_extrinsic[6];       /* PriorFactor<Pose3> sigmas in tangent order, ArmSlamCalib.cpp:684 */
    int32_t use_dead_band;           /* useDeadBand, ArmSlamCalib.h:94 (0) */
    double dead_band;                /* deadBandSize, ArmSlamCalib.h:96 (0.05) */
    int32_t pose3_chart;             /* enum asc_pose3_chart */
    /* optimiser parameters: GTSAM 3.2 defaults */
    int32_t max_iterations;          /* 100 */
    double relative_error_tol;       /* 1e-5 */
    double absolute_error_tol;       /* 1e-5 */
    double error_tol;                /* 0 */
    double lambda_initial;           /* 1e-5 */
    double lambda_factor;            /* 10 */
    double lambda_upper_bound;       /* 1e5 */
    double lambda_lower_bound;       /* 0 */
    double min_model_fidelity;       /* 1e-3 */
    double delta_initial;            /* Dogleg trust region, 1.0 */
    /* linear solver (replaces SEQUENTIAL_CHOLESKY, ArmSlamCalib.cpp:461) */
    double pcg_rel_tol;              /* ||r||_M^-1 / ||b||_M^-1 stopping ratio (1e-12) */
    int32_t pcg_max_iters;           /* 2000 */
    int32_t device;                  /* CUDA device ordinal, -1 = current */
    /* block-Jacobi preconditioner blocks = runs of consecutive, co-visible keyframes ("clusters") */
    int32_t pcg_cluster_max;         /* max keyframes per cluster (10); <=1: one block per keyframe */
    double pcg_cluster_covis;        /* start a new cluster when the share of common landmarks with the
                                        previous keyframe drops below this (0.25) */
    /* second preconditioner level: coarse space = one joint-offset vector per cluster + the extrinsic; its block-sparse operator is
       solved by a multilevel V-cycle (Chebyshev smoothers on N x N block rows, aggregation, small dense top level) written for
       this library (csrc/coarse_ml.cuh) -- no size limit, no cuSOLVER/cuBLAS */
    int32_t pcg_coarse_levels;       /* max smoothed levels of the V-cycle (8); 0 switches the coarse level off */
    int32_t pcg_ml_degree;           /* Chebyshev degree of the pre- and post-smoother (4) */
    int32_t pcg_ml_agg;              /* nodes per aggregate (8); enlarged up to 64 when that reaches the dense top level in one step;
                                        negative: exactly |value|, never enlarged */
    int32_t pcg_ml_top_dim;          /* aggregate until the coarsest level has at most this many unknowns (1024, the maximum); it is inverted densely */
} asc_params;

/* One record per outer iterate() of the optimiser */
typedef struct asc_iter_log {
    int32_t iteration;     /* 1-based */
    int32_t inner_steps;   /* damped solves (LM) or trust-region tries (Dogleg) in this iterate() */
    int32_t accepted;      /* 1 if the estimate moved */
    int32_t pcg_iterations;/* summed over inner steps */
    double error;          /* graph.error(values) after the iteration */
    double lambda_or_delta;/* lambda (LM) or Delta (Dogleg) after the iteration */
    double linearize_ms;   /* device time of the linearisation + assembly kernels */
    double solve_ms;       /* device time of everything after the linearisation (all inner steps) */
    /* per-stage device times inside solve_ms, summed over the inner steps (the role of the reference's Timer.cpp:41-89 tables) */
    double prepare_ms;     /* landmark-block inverses, cluster blocks and their inverses, reduced right-hand side */
    double coarse_ms;      /* coarse-level refresh: block-sparse operator, multilevel set-up, border elimination */
    double pcg_ms;         /* the PCG loop */
} asc_iter_log;

typedef struct asc_linear_stats {
    double error;          /* 0.5 * sum ||whitened, reweighted residual||^2 over all factors */
    double linearize_ms;   /* device time of the FK + projection-linearise kernel(s) only */
    double assemble_ms;    /* device time of landmark/pose block assembly */
    int64_t cheirality;    /* number of projection factors that hit the z<0.01 guard */
} asc_linear_stats;

void asc_default_params(asc_params* p);
const char* asc_last_error(void);
const char* asc_version(void);

int asc_create(const asc_params* params, asc_handle** out);
int asc_destroy(asc_handle* h);
/* Replace the optimiser-control fields (iteration limits, tolerances, lambda/Delta schedule).  Noise sigmas, the
 * Cauchy k and the cluster settings are baked into the packed problem and take effect at the next asc_set_problem. */
int asc_set_params(asc_handle* h, const asc_params* params);

/*
 * Serial revolute chain; replaces the DART skeleton/group the factors hold
 * (RobotProjectionFactor.h:33-35, ArmSlamCalib.cpp:224-272).  Joint i:
 *   T_world<-child(i) = T_world<-parent(i) * T_p2j[i] * Rot(axis[i], q[i]) * inverse(T_c2j[i])
 * The camera link is child(N-1).  t_base (nullable) places the chain root.
 */
int asc_set_chain(asc_handle* h, int32_t n_dof, const double* t_p2j /*[N][12]*/,
                  const double* axis /*[N][3]*/, const double* t_c2j /*[N][12]*/,
                  const uint8_t* continuous /*[N], RobotConfig.h:72-77*/,
                  const double* t_base /*[12] or NULL*/);

/* Cal3_S2(fx, fy, 0, cx, cy), ArmSlamCalib.cpp:670 */
int asc_set_camera(asc_handle* h, double fx, double fy, double cx, double cy);

/*
 * The factor graph, flattened.  Observation m is RobotProjectionFactor(uv[m], q_{pose_idx[m]},
 * l_{lm_idx[m]}, K0) (ArmSlamCalib.cpp:1625-1646); pose p has EncoderFactor(enc[p]) iff
 * has_enc[p] (ArmSlamCalib.cpp:299-305; NULL = all); landmark j has PriorFactor<Point3>
 * (lm_prior[j]) (ArmSlamCalib.cpp:1648-1651); K0 has PriorFactor<Pose3>(x_prior) (:685).
 * The order of the observation list is preserved in every per-observation output.
 */
int asc_set_problem(asc_handle* h, int32_t n_poses, int32_t n_landmarks, int64_t n_obs,
                    const int32_t* pose_idx, const int32_t* lm_idx, const double* uv /*[M][2]*/,
                    const double* enc /*[P][N]*/, const uint8_t* has_enc /*[P] or NULL*/,
                    const double* lm_prior /*[L][3]*/, const double* x_prior /*[12]*/);

/* Values: q_i (RobotConfig), l_j (Point3), K0 (Pose3) */
/*
 * Optional DriftFactor links (include/arm_slam_calib/DriftFactor.h:44-61; added by AddDriftFactor when use_drift_noise is
 * set, src/ArmSlamCalib.cpp:317-320, :1653-1668): has_drift[p] != 0 <=> DriftFactor(q_{p-1}, q_p) with measurement
 * drift_meas[p][N] = q1Encoders - q0Encoders and Diagonal noise sigma_drift[N] (drift_noise_level, default 0.05).
 * residual (q_p - q_{p-1}) - drift_meas[p], J0 = -I, J1 = +I: a block-tridiagonal term in the keyframe part of the normal
 * equations.  Call after asc_set_problem (which drops earlier links); has_drift == NULL removes them.
 */
int asc_set_drift(asc_handle* h, const uint8_t* has_drift /*[P]*/, const double* drift_meas /*[P][N]*/, const double* sigma_drift /*[N]*/);
/*
 * Optional VelocityFactor links (include/arm_slam_calib/VelocityFactor.h:40-62; AddVelocityFactor when use_velocity_noise is set,
 * src/ArmSlamCalib.cpp:322-325, :1671-1686): has_velocity[p] != 0 <=> VelocityFactor(q_{p-1}, q_p, velocities[p-1]) with
 * velocity_meas[p][N] = qDot and Diagonal noise sigma_velocity[N] (velocity_noise_level, default 0.01).  residual
 * (q_p - q_{p-1}) - qDot; its central-difference Jacobians of a linear function are exactly J0 = -I, J1 = +I, so it is the
 * DriftFactor's analytic twin and shares its kernels.  Same call rules as asc_set_drift; both kinds may be present.
 */
int asc_set_velocity(asc_handle* h, const uint8_t* has_velocity /*[P]*/, const double* velocity_meas /*[P][N]*/, const double* sigma_velocity /*[N]*/);
int asc_set_values(asc_handle* h, const double* q /*[P][N]*/, const double* l /*[L][3]*/,
                   const double* x /*[12]*/);
int asc_get_values(asc_handle* h, double* q, double* l, double* x);

/*
 * Incremental mode (the drivers call OptimizeStep every one or two frames on a growing graph, sim_genrobot.cpp:97-104; SimulationStep
 * adds the configuration, its EncoderFactor and the new projection factors, ArmSlamCalib.cpp:307-327): append keyframes, landmarks and
 * observations to the packed problem.  New observations index the GROWN sets (poses 0 .. P + n_new_poses - 1, landmarks 0 .. L +
 * n_new_landmarks - 1) and are appended after the existing ones; the current estimate of the existing variables is kept (warm start),
 * new variables start at q_new / l_new.  Link factors must be set again afterwards.
 */
int asc_append(asc_handle* h, int32_t n_new_poses, int32_t n_new_landmarks, int64_t n_new_obs, const int32_t* pose_idx,
               const int32_t* lm_idx, const double* uv /*[n_new_obs][2]*/, const double* enc_new /*[n_new_poses][N]*/,
               const uint8_t* has_enc_new /*[n_new_poses] or NULL*/, const double* lm_prior_new /*[n_new_landmarks][3]*/,
               const double* q_new /*[n_new_poses][N]*/, const double* l_new /*[n_new_landmarks][3]*/);

/* NonlinearFactorGraph::error(values) */
int asc_error(asc_handle* h, double* error);

/* NonlinearFactorGraph::linearize(values) + block assembly; exposed for the
 * "factor linearisations/s" metric and per-kernel parity */
int asc_linearize(asc_handle* h, asc_linear_stats* stats);

/* {LevenbergMarquardt,Dogleg}Optimizer::optimize(); the estimate is updated in place */
int asc_optimize(asc_handle* h, int32_t mode, asc_iter_log* log, int32_t log_capacity,
                 int32_t* n_iterations, double* final_error);

/* key of the variable named by the last ASC_ERR_INDETERMINATE:
 * kind 'l' / 'q' / 'K', index as in Symbol(kind, index) (ArmSlamCalib.h:284-286) */
int asc_indeterminate_key(asc_handle* h, char* kind, int64_t* index);

/* ---- parity / inspection accessors (valid after asc_linearize) ------------------- */
/* whitened + reweighted per-observation blocks, original observation order */
int asc_get_projection_blocks(asc_handle* h, double* b /*[M][2]*/, double* jq /*[M][2][N]*/,
                              double* jl /*[M][2][3]*/);
/* undamped landmark blocks: C_j (3x3 full), g_l (3), W_Kj (6x3) */
int asc_get_landmark_blocks(asc_handle* h, double* c /*[L][9]*/, double* gl /*[L][3]*/,
                            double* wk /*[L][18]*/);
/* undamped pose blocks: B_pp (NxN full), B_pK (Nx6), g_p (N); and B_KK (6x6), g_K (6) */
int asc_get_pose_blocks(asc_handle* h, double* bpp /*[P][N][N]*/, double* bpk /*[P][N][6]*/,
                        double* gp /*[P][N]*/, double* bkk /*[36]*/, double* gk /*[6]*/);
/* Solve (H + lambda I) delta = g with the device Schur/PCG path; delta_q [P][N], delta_l [L][3],
 * delta_x [6].  Returns the PCG iteration count. */
int asc_solve_damped(asc_handle* h, double lambda, double* delta_q, double* delta_l,
                     double* delta_x, int32_t* pcg_iterations);

/*
 * Marginal covariance of the extrinsic K0 at the current values: the K0 block of (A^T A)^-1 of the linearised graph --
 * what isam2.marginalCovariance(ExtrinsicSymbol()) hands to DrawState in the reference (src/ArmSlamCalib.cpp:490-493,
 * :976-998) and what gtsam::Marginals(graph, values).marginalCovariance(K0) returns on the batch path.  Computed from
 * the operators of the solve: (S^-1)_KK of the UNDAMPED reduced system, six PCG solves with unit right-hand sides.
 * cov36: 6x6 row-major, tangent order [omega, v].  ASC_ERR_INDETERMINATE if the system is singular.
 */
int asc_extrinsic_marginal(asc_handle* h, double* cov36, int32_t* pcg_iterations);

/* ---- measurement support (bench.py) ----------------------------------------------- */
/* device-resident snapshot of the current values / restore it (keeps the timed region free of H2D copies) */
int asc_save_values(asc_handle* h);
int asc_restore_values(asc_handle* h);
/* number of kernels this handle has launched so far */
int64_t asc_launch_count(asc_handle* h);
/* Sizes of the preconditioner of the packed problem (for byte accounting): clusters, coarse dimension (0 = no coarse level),
 * total doubles in the cluster blocks. */
int asc_solver_info(asc_handle* h, int32_t* n_clusters, int32_t* coarse_dim, int64_t* cluster_block_doubles);
/* Hierarchy of the multilevel coarse solve: number of levels (the last one is the dense top level; 0 = no coarse level), nodes and
 * non-zero N x N blocks of each level (arrays of `capacity` entries), Chebyshev degree of the smoothers. */
int asc_coarse_info(asc_handle* h, int32_t* n_levels, int32_t* level_nodes, int64_t* level_blocks, int32_t capacity, int32_t* degree);
/* debugging: barrier timestamps (globaltimer ns; [barrier][CTA][arrive, release]) of the last V-cycle launch when ASC_ML_TRACE is set */
int asc_debug_ml_trace(asc_handle* h, unsigned long long* out, int32_t capacity);
/* CUDA-event stopwatch recorded on the stream this handle launches its kernels on */
int asc_timer_start(asc_handle* h);
int asc_timer_stop(asc_handle* h, double* elapsed_ms);
/* Bracket up to max_samples launches of one kernel with CUDA events on the launching stream while the normal
 * entry points run: 1 linearise (K1+K2b), 2 landmark blocks (K2a), 3 Schur pass A, 4 pass B, 5 pass C, 6 error (K1e) */
int asc_profile_begin(asc_handle* h, int32_t kernel_id, int32_t max_samples);
int asc_profile_read(asc_handle* h, int32_t* n_samples, double* mean_ms, double* min_ms);

/* ---- multi-GPU: landmark sharding --------------------------- */
/* Size in bytes of the opaque communicator id (ncclUniqueId). */
int asc_comm_id_size(void);
/* rank 0 creates the id; the caller broadcasts the bytes (e.g. torch.distributed). */
int asc_comm_create_id(void* id_bytes);
/* Collective: every rank calls with the same id.  After this, asc_set_problem expects only this
 * rank's shard of landmarks/observations (see asc_shard_by_landmark) with q/enc replicated. */
int asc_comm_init(asc_handle* h, int32_t rank, int32_t world_size, const void* id_bytes);
/* Host-side partitioner: contiguous landmark ranges balanced by observation count.
 * lm_begin[r], lm_begin[r+1] bound rank r's landmark ids. */
int asc_shard_by_landmark(int32_t n_landmarks, int64_t n_obs, const int32_t* lm_idx,
                          int32_t world_size, int32_t* lm_begin /*[world_size+1]*/);

/* ---- synthetic problems (host side; ArmSlamCalib.cpp:535-602, :654-706, :1773-1887) ---- */
typedef struct asc_sim_config {
    int32_t trajectory_size;   /* trajectory_length, ArmSlamCalib.h:70 */
    int32_t num_landmarks_x, num_landmarks_y; /* ArmSlamCalib.h:71-72 */
    double landmark_size_x, landmark_size_y;  /* ArmSlamCalib.h:73-74 */
    double landmark_z_min, landmark_z_max;    /* utils::Rand(-5, 5), ArmSlamCalib.cpp:662 */
    double max_range;          /* [NEW] visibility thinning, 0 = none */
    double encoder_noise;      /* encoderNoiseLevel */
    double velocity_noise;     /* velocityNoiseLevel */
    int32_t seed;              /* ros param "seed", ArmSlamCalib.cpp:193,:542 */
    int32_t min_observations;  /* admission rule observations.size() > 3, ArmSlamCalib.cpp:1860 */
    double sim_extrinsic[12];  /* simExtrinsic (ground truth) */
    double extrinsic_guess[12];/* extrinsicInitialGuess */
    int32_t segment_length;    /* [NEW] 0 = one random walk (reference); k>0 = restart the reference's walk from a
                                  random in-limits configuration every k steps, so long trajectories keep exploring
                                  instead of saturating at the joint limits */
} asc_sim_config;

typedef struct asc_sim_problem asc_sim_problem;

void asc_sim_default_config(asc_sim_config* c);
/* chain fixtures: writes N x {t_p2j[12], axis[3], t_c2j[12]}; returns N or <0 */
int asc_chain_mico(double* t_p2j, double* axis, double* t_c2j);
/* RobotGenerator::GenerateRobot(numDof, scale, growth, randomLength, randomAngle, false, 6),
 * src/RobotGenerator.cpp:22-147; consumes rand() like the reference (call srand first) */
int asc_chain_generate(int32_t n_dof, double scale, double growth, double random_length,
                       double random_angle, double* t_p2j, double* axis, double* t_c2j);

int asc_sim_create(const asc_sim_config* cfg, int32_t n_dof, const double* t_p2j,
                   const double* axis, const double* t_c2j, double fx, double fy, double cx,
                   double cy, asc_sim_problem** out);
void asc_sim_destroy(asc_sim_problem* s);
/* sizes: P, L (landmarks in graph), M, L_sim (all simulated landmarks) */
int asc_sim_sizes(const asc_sim_problem* s, int32_t* n_poses, int32_t* n_landmarks,
                  int64_t* n_obs, int32_t* n_sim_landmarks);
/* copies out the flattened graph; lm_id[j] is the reference landmark id of compact index j;
 * factor_order[m] is the position of observation m in the reference's graph (AddFactor order) */
int asc_sim_export(const asc_sim_problem* s, int32_t* pose_idx, int32_t* lm_idx, double* uv,
                   double* enc, double* q_truth, double* lm_truth, int32_t* lm_id,
                   int64_t* factor_order);

/* host forward kinematics of the chain (used by the generator; exported for tests) */
/* srand() of the libc stream utils::Rand draws from (sim_genrobot.cpp:60) */
void asc_srand(unsigned seed);

int asc_chain_fk(int32_t n_dof, const double* t_p2j, const double* axis, const double* t_c2j,
                 const double* q, double* t_link /*[12]*/);

#ifdef __cplusplus
}
#endif
#endif /* ASC_H_ */
