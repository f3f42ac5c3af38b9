"""Dev: per-barrier arrival / release timestamps of the persistent V-cycle kernel (ASC_ML_TRACE=1)."""
import os, sys
os.environ["ASC_ML_TRACE"] = "1"
os.environ["ASC_NO_SIDE_VCYCLE"] = "1"
sys.path.insert(0, ".")
import numpy as np
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs
from arm_slam_calib_b200.capi import _ptr
pb = configs.c2()
prm = asc.default_params()
prm.pcg_max_iters = 40
opt = asc.BatchOptimizer(pb, prm)
try:
    opt.solve_damped(1e-5)
except Exception as e:
    print("solve:", str(e)[:80])
ctas = int(os.environ.get("ASC_ML_CTAS", "128"))
buf = np.zeros(64 * 148 * 2, dtype=np.uint64)
assert opt.L.asc_debug_ml_trace(opt.h, _ptr(buf), buf.size) == 0
t = buf[:32 * ctas * 2].reshape(32, ctas, 2).astype(np.int64)
nb = int((t[:, 0, 0] > 0).sum())
t0 = t[0, :, 0].min()
print("barriers traced:", nb)
prev_release = None
for b in range(nb):
    arr, rel = t[b, :, 0] - t0, t[b, :, 1] - t0
    work = "" if prev_release is None else " work(arrive - prev release): min %5.2f med %5.2f max %5.2f us" % tuple(np.percentile(arr - prev_release, [0, 50, 100]) / 1e3)
    print("barrier %2d: arrive min %7.2f med %7.2f max %7.2f | release min %7.2f max %7.2f us |%s  slowest CTAs %s" % (
        b, arr.min() / 1e3, np.median(arr) / 1e3, arr.max() / 1e3, rel.min() / 1e3, rel.max() / 1e3, work, np.argsort(-arr)[:4].tolist()))
    prev_release = rel
