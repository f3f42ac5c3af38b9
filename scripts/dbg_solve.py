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
prm = asc.default_params()
for i in range(20): prm.sigma_encoder[i] = 0.03
orc = Oracle(pb, prm)
# walk a few LM iterations with the oracle to get a mid-run state
prm3 = asc.default_params(); prm3.max_iterations = 3
for i in range(20): prm3.sigma_encoder[i] = 0.03
st = Oracle(pb, prm3).optimize(pb.q0, pb.l0, pb.x0, mode=0, solver=0)
vals = (st["q"], st["l"], st["x"])
for lam in (1e-6, 1e-4, 1e-2):
    rc, dq, dl, dx, _ = orc.solve_damped(*vals, lam, solver=0)
    rc, pq, pl, px, pit = orc.solve_damped(*vals, lam, solver=1)
    nrm = np.sqrt((dq**2).sum() + (dl**2).sum() + (dx**2).sum())
    def dist(a, b, c): return np.sqrt(((a-dq)**2).sum() + ((b-dl)**2).sum() + ((c-dx)**2).sum()) / nrm
    print("lam %g |delta| %.3g  cpu-pcg(%d its) rel diff %.2e" % (lam, nrm, pit, dist(pq, pl, px)))
    for tol in (1e-12, 1e-14):
        p2 = asc.default_params(); p2.pcg_rel_tol = tol
        for i in range(20): p2.sigma_encoder[i] = 0.03
        opt = asc.BatchOptimizer(pb, p2)
        opt.set_values(*vals)
        gq, gl, gx, its = opt.solve_damped(lam)
        print("    gpu tol %g: %d its rel diff %.2e (q %.2e l %.2e x %.2e)" % (tol, its, dist(gq, gl, gx), np.abs(gq-dq).max()/np.abs(dq).max(), np.abs(gl-dl).max()/np.abs(dl).max(), np.abs(gx-dx).max()/np.abs(dx).max()))
        opt.close()
