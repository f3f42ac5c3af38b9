"""Dev: how loose can the PCG tolerance be before per-iteration error parity (1e-6 rel) with exact Cholesky breaks?"""
import sys
import numpy as np
sys.path.insert(0, ".")
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs
from oracle import Oracle
for name in ["c1", "c1off"]:
    pb = configs.c1(extrinsic_offset=(name == "c1off"))
    prm = asc.default_params()
    ref = Oracle(pb, prm).optimize(pb.q0, pb.l0, pb.x0, mode=0, solver=0)
    for tol in [1e-12, 1e-10, 1e-8, 1e-6, 1e-4]:
        prm = asc.default_params(); prm.pcg_rel_tol = tol
        opt = asc.BatchOptimizer(pb, prm)
        q, l, x = opt.optimize(0)
        n = min(len(opt.log), len(ref["log"]))
        rel = max(abs(a.error - b.error) / b.error for a, b in zip(opt.log[:n], ref["log"][:n]))
        print("%s tol %g: iters %d (ref %d) pcg total %d  max rel err diff %.2e  |dq| %.2e |dx| %.2e" % (
            name, tol, len(opt.log), len(ref["log"]), sum(g.pcg_iterations for g in opt.log), rel,
            np.abs(q - ref["q"]).max(), np.abs(x - ref["x"]).max()))
        opt.close()
