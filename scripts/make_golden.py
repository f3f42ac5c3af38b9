"""Generates tests/golden/c1_golden.npz from the CPU oracle (exact Cholesky solver).  The reference has no golden
vectors of its own; this file freezes OUR restatement so later changes are caught."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs
from oracle import Oracle

pb = configs.c1()
orc = Oracle(pb, asc.default_params())
out = dict(pose_idx=pb.pose_idx, lm_idx=pb.lm_idx, lm_id=pb.meta["lm_id"], factor_order=pb.meta["factor_order"], uv=pb.uv, enc=pb.enc)
for mode, key in ((0, "lm"), (1, "dl")):
    r = orc.optimize(pb.q0, pb.l0, pb.x0, mode=mode, solver=0)
    out[key + "_errors"] = np.array([g.error for g in r["log"]])
    out[key + "_lambda"] = np.array([g.lambda_or_delta for g in r["log"]])
    out[key + "_q"], out[key + "_l"], out[key + "_x"] = r["q"], r["l"], r["x"]
    print(key, r["iterations"], r["error"])
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
