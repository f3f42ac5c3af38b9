"""Dev: per-iteration GPU vs oracle log on the 40-pose problem (facade / smoke problem)."""
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
opt = asc.BatchOptimizer(pb, prm)
print("coarse", opt.coarse_info(), "solver", opt.solver_info())
q, l, x = opt.optimize(0)
ref = Oracle(pb, prm).optimize(pb.q0, pb.l0, pb.x0, mode=0, solver=0)
print("gpu its", opt.iterations(), "oracle its", ref["iterations"])
for i in range(max(len(opt.log), len(ref["log"]))):
    a = opt.log[i] if i < len(opt.log) else None
    b = ref["log"][i] if i < len(ref["log"]) else None
    print(i + 1, (a.inner_steps, a.accepted, "%.12g" % a.error, a.lambda_or_delta, a.pcg_iterations) if a else None,
          (b.inner_steps, b.accepted, "%.12g" % b.error, b.lambda_or_delta) if b else None,
          "rel %.2e" % (abs(a.error - b.error) / b.error) if a and b else "")
