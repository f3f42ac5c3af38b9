"""Dev probe: C2 linearisation only (K0, K1+K2b, K2a, K1e), a few calls, event timings printed.  Short enough for ncu."""
import sys
sys.path.insert(0, ".")
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pb = getattr(configs, name)()
opt = asc.BatchOptimizer(pb, configs.params_for(name, asc.default_params))
for kid in (1, 7, 2, 6):
    opt.profile_begin(kid, 64)
    for _ in range(reps):
        if kid == 6:
            opt.graph_error()
        else:
            st = opt.linearize()
    n, mean_ms, min_ms = opt.profile_read()
    print("kernel %d: %d samples mean %.1f us min %.1f us" % (kid, n, 1e3 * mean_ms, 1e3 * min_ms))
print("error %.12g cheir %d" % (st.error, st.cheirality))
opt.close()
