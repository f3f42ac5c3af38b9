import sys
import numpy as np
sys.path.insert(0, ".")
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200.sim import pose12, rodrigues_y
from oracle import Oracle
n_dof = int(sys.argv[1])
chain = asc.chain_generate(n_dof, seed=n_dof)
R = rodrigues_y(1.57)
pb = asc.make_sim_problem(chain=chain, trajectory_size=60, seed=n_dof, srand=False, encoder_noise=0.03, sim_extrinsic=pose12(R=R), extrinsic_guess=pose12(R=R, t=[0.1, 0.1, 0.1]))
print("P L M", pb.n_poses, pb.n_landmarks, pb.n_obs)
for variant in ["default", "tol1e-14", "nocoarse", "fp64coarse", "percluster1"]:
    prm = asc.default_params()
    for i in range(20): prm.sigma_encoder[i] = 0.03
    if variant == "tol1e-14": prm.pcg_rel_tol = 1e-14
    if variant == "nocoarse": prm.pcg_coarse_levels = 0
    if variant == "deg2": prm.pcg_ml_degree = 2
    if variant == "percluster1": prm.pcg_cluster_max = 1
    orc = Oracle(pb, prm)
    ref = orc.optimize(pb.q0, pb.l0, pb.x0, mode=0, solver=0)
    opt = asc.BatchOptimizer(pb, prm)
    q, l, x = opt.optimize(0)
    print(variant, "iters", opt.iterations(), ref["iterations"])
    for a, r in zip(opt.log, ref["log"]):
        print("  it %2d inner %d/%d pcg %5d err %.10g ref %.10g rel %.2e lam %.1e" % (a.iteration, a.inner_steps, r.inner_steps, a.pcg_iterations, a.error, r.error, abs(a.error - r.error) / r.error, a.lambda_or_delta))
    opt.close()
