"""Dev probe: run one optimize() on a named workload and print the per-iteration log."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 100
t = time.time()
pb = getattr(configs, name)()
print("%s: P=%d L=%d M=%d N=%d  (generated in %.1fs)" % (name, pb.n_poses, pb.n_landmarks, pb.n_obs, pb.n_dof, time.time() - t), flush=True)
prm = configs.params_for(name, asc.default_params)
prm.max_iterations = maxit
if len(sys.argv) > 4:
    prm.pcg_rel_tol = float(sys.argv[4])
if len(sys.argv) > 5:
    prm.pcg_ml_degree = int(sys.argv[5])
if len(sys.argv) > 6:
    prm.pcg_cluster_max = int(sys.argv[6])
t = time.time()
opt = asc.BatchOptimizer(pb, prm)
print("setup %.2fs" % (time.time() - t), flush=True)
st = opt.linearize()
print("linearize: error %.9g  lin %.3f ms  asm %.3f ms  cheir %d" % (st.error, st.linearize_ms, st.assemble_ms, st.cheirality))
st = opt.linearize()
print("linearize: error %.9g  lin %.3f ms  asm %.3f ms  -> %.3g factor linearisations/s" % (st.error, st.linearize_ms, st.assemble_ms, pb.n_obs / (st.linearize_ms * 1e-3)))
t = time.time()
q, l, x = opt.optimize(mode)
dt = time.time() - t
print("optimize: %d iterations, error %.9g, %.3fs wall -> %.2f iters/s" % (opt.iterations(), opt.error(), dt, opt.iterations() / dt))
for g in opt.log:
    print("  it %2d inner %d acc %d pcg %5d  err %.9g  lam/delta %.3g  lin %.2f ms  solve %.2f ms (%.1f us/pcg-it)" % (
        g.iteration, g.inner_steps, g.accepted, g.pcg_iterations, g.error, g.lambda_or_delta, g.linearize_ms, g.solve_ms,
        1e3 * g.solve_ms / max(g.pcg_iterations, 1)))
qt = pb.meta["q_truth"]
print("mean |q - truth|: before %.4g after %.4g" % (np.abs(pb.q0 - qt).mean(), np.abs(q - qt).mean()))
print("extrinsic t:", x.reshape(3, 4)[:, 3])
