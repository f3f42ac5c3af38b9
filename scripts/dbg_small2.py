"""Dev: PCG iteration counts of the damped solves at the state after 3 LM iterations of the 40-pose problem."""
import sys
sys.path.insert(0, ".")
import numpy as np
import arm_slam_calib_b200 as asc
from oracle import Oracle
pb = asc.make_sim_problem(trajectory_size=40, seed=3)
prm = asc.default_params()
for k in sys.argv[1:]:
    name, v = k.split("=")
    setattr(prm, name, type(getattr(prm, name))(float(v)))
prm.max_iterations = 3
opt = asc.BatchOptimizer(pb, prm)
q, l, x = opt.optimize(0)
print("after 3 its: error", opt.error(), "lambda", opt.log[-1].lambda_or_delta)
orc = Oracle(pb, prm)
for rep in range(2):
    for lam in (1e-4, 1e-3, 1e-2, 1e-1):
        try:
            dq, dl, dx, its = opt.solve_damped(lam)
            rc, oq, ol, ox, oits = orc.solve_damped(q, l, x, lam, solver=0)
            print("lambda %g: gpu its %d  rel diff vs exact %.2e" % (lam, its, np.abs(dq - oq).max() / np.abs(oq).max()))
        except Exception as e:
            print("lambda %g: %s" % (lam, e))
    opt.set_values(q, l, x)   # fresh linearisation: no warm start for the first lambda
