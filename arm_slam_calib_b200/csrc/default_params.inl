// asc_default_params: shared by libasc_b200.so (asc_api.cu) and the host-only libasc_host.so (host_only.cpp).
// Defaults = ArmSlamCalib::Params() (ArmSlamCalib.h:69-105) and the GTSAM 3.2 optimiser defaults.
void asc_default_params(asc_params* p) {
    std::memset(p, 0, sizeof *p);
    p->sigma_px = 1.0;            // ArmSlamCalib.h:86
    p->sigma_landmark = 50.0;     // :83
    p->cauchy_k = 3.0;            // ArmSlamCalib.cpp:197
    for (int i = 0; i < ASC_MAX_DOF; ++i) p->sigma_encoder[i] = 0.1;   // ArmSlamCalib.h:79
    for (int i = 0; i < 6; ++i) p->sigma_extrinsic[i] = 0.1;           // :81-82
    p->use_dead_band = 0;
    p->dead_band = 0.05;
    p->pose3_chart = ASC_POSE3_FIRST_ORDER_CAYLEY;
    p->max_iterations = 100;
    p->relative_error_tol = 1e-5;
    p->absolute_error_tol = 1e-5;
    p->error_tol = 0.0;
    p->lambda_initial = 1e-5;
    p->lambda_factor = 10.0;
    p->lambda_upper_bound = 1e5;
    p->lambda_lower_bound = 0.0;
    p->min_model_fidelity = 1e-3;
    p->delta_initial = 1.0;
    p->pcg_rel_tol = 1e-12;
    p->pcg_max_iters = 2000;
    p->device = -1;
    p->pcg_cluster_max = 10;
    p->pcg_cluster_covis = 0.25;
    p->pcg_coarse_levels = 8;
    p->pcg_ml_degree = 4;
    p->pcg_ml_agg = 8;
    p->pcg_ml_top_dim = 1024;
}

