"""Dev: device time of one V-cycle (CUDA events on the launching stream, un-graphed, main stream) on C2."""
import sys
sys.path.insert(0, ".")
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs
pb = configs.c2()
prm = asc.default_params()
for k in sys.argv[1:]:
    name, v = k.split("=")
    setattr(prm, name, type(getattr(prm, name))(float(v)))
opt = asc.BatchOptimizer(pb, prm)
print("hierarchy", opt.coarse_info())
opt.linearize()
import os
kids = ((8, "vcycle"),) if os.environ.get("VC_ONLY") else ((8, "vcycle"), (3, "pass A"), (4, "pass B"), (5, "pass C"), (9, "cluster update"))
prm.pcg_max_iters = 64
for kid, name in kids:
    opt.linearize()
    opt.profile_begin(kid, 64)
    its = -1
    try:
        dq, dl, dx, its = opt.solve_damped(1e-5)
    except Exception as e:
        print("   (solve raised: %s)" % str(e)[:60])
    n, mean, mn = opt.profile_read()
    print("%-14s samples %d mean %.1f us min %.1f us  (pcg its %d)" % (name, n, 1e3 * mean, 1e3 * mn, its))
