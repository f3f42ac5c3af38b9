"""Dev probe: run one optimize() on a named workload (bench.py names: c2, c3, c4, c4_12, c5, c2x<k>) and print the per-iteration log.
usage: python scripts/probe.py <workload> [max_iterations] [name=value asc_params overrides ...]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import arm_slam_calib_b200 as asc
from bench import make_workload

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
maxit = int(sys.argv[2]) if len(sys.argv) > 2 and "=" not in sys.argv[2] else 100
t = time.time()
pb, prm, mode, desc, eq, name = make_workload(name, 1)
print("%s: P=%d L=%d M=%d N=%d  (generated in %.1fs)" % (name, pb.n_poses, pb.n_landmarks, pb.n_obs, pb.n_dof, time.time() - t), flush=True)
prm.max_iterations = maxit
for k in sys.argv[2:]:
    if "=" in k:
        n_, v = k.split("=")
        setattr(prm, n_, type(getattr(prm, n_))(float(v)))
t = time.time()
opt = asc.BatchOptimizer(pb, prm)
print("setup %.2fs  clusters/coarse %s  hierarchy %s" % (time.time() - t, opt.solver_info(), opt.coarse_info()), flush=True)
st = opt.linearize()
st = opt.linearize()
print("linearize: error %.9g  lin %.3f ms  asm %.3f ms  cheir %d -> %.3g factor linearisations/s (whole linearisation)" % (
    st.error, st.linearize_ms, st.assemble_ms, st.cheirality, pb.n_obs / ((st.linearize_ms + st.assemble_ms) * 1e-3)))
t = time.time()
try:
    q, l, x = opt.optimize(mode)
except Exception as e:
    print("optimize raised:", e)
    q, l, x = opt.values()
dt = time.time() - t
print("optimize: %d iterations, error %.9g, %.3fs wall -> %.2f iters/s" % (opt.iterations(), opt.error(), dt, opt.iterations() / dt))
for g in opt.log:
    print("  it %2d inner %d acc %d pcg %5d  err %.9g  lam/delta %.3g  lin %.2f ms  solve %.2f ms = prepare %.2f + coarse %.2f + pcg %.2f (%.1f us/pcg-it) + rest" % (
        g.iteration, g.inner_steps, g.accepted, g.pcg_iterations, g.error, g.lambda_or_delta, g.linearize_ms, g.solve_ms, g.prepare_ms, g.coarse_ms, g.pcg_ms,
        1e3 * g.pcg_ms / max(g.pcg_iterations, 1)))
qt = pb.meta["q_truth"]
print("mean |q - truth|: before %.4g after %.4g" % (np.abs(pb.q0 - qt).mean(), np.abs(q - qt).mean()))
print("extrinsic t:", x.reshape(3, 4)[:, 3])
