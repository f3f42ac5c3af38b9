import sys, time
sys.path.insert(0, ".")
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs
pb = configs.c2()
prm = asc.default_params(); prm.max_iterations = 3
opt = asc.BatchOptimizer(pb, prm)
opt.optimize(0)
for i in range(2):
    t0 = time.perf_counter(); opt.set_problem(pb); t1 = time.perf_counter(); opt.set_values(pb.q0, pb.l0, pb.x0); t2 = time.perf_counter()
    opt.optimize(0); t3 = time.perf_counter(); opt.values(); t4 = time.perf_counter()
    print("set_problem %.1f ms  set_values %.1f ms  optimize %.1f ms  get_values %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)))
