"""Error metrics and result files of the reference, in its own text formats, so that data/plot_*.py read them unchanged.

Host-side (numpy) mirror of the step right AFTER the optimisation path in the reference:
`ArmSlamCalib::ComputeError` (src/ArmSlamCalib.cpp:1360-1412), `CalculateCurrentErrors` (:2061-2076), `PrintSimErrors`
(:2216-2294), `ComputeLatestJointAngleOffsets` / `ComputeLatestGroundTruthSimOffsets` (:723-765), `Diff` (:1230-1242) and
the per-step log lines of the shipped driver (src/sim_genrobot.cpp:86-175).  Nothing here is on the GPU path: these are
O(P + L + M) reductions over the values `asc_get_values` returns.
"""
import os

import numpy as np


# ------------------------------------------------------------------ kinematics on the host (DART convention)
def _T(m12):
    T = np.eye(4)
    T[:3, :] = np.asarray(m12, dtype=np.float64).reshape(3, 4)
    return T


def _rot(axis, th):
    a = np.asarray(axis, dtype=np.float64)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def link_pose(pb, q):
    """World pose (4x4) of the camera link: T_base * prod_i T_p2j[i] * Rot(axis[i], q[i]) * inverse(T_c2j[i])."""
    T = np.eye(4) if pb.t_base is None else _T(pb.t_base)
    for i in range(pb.n_dof):
        R = np.eye(4)
        R[:3, :3] = _rot(pb.axis[i], q[i])
        T = T @ _T(pb.t_p2j[i]) @ R @ np.linalg.inv(_T(pb.t_c2j[i]))
    return T


def camera_pose(pb, q, x12):
    """ArmSlamCalib::GetCameraPose(q, extrinsic): T_link(q) * extrinsic."""
    return link_pose(pb, q) @ _T(x12)


def quaternion_xyzw(R):
    """Rotation matrix -> (x, y, z, w), the branch structure of Eigen::Quaternion(Matrix3) that Rot3::toQuaternion uses."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
    out = np.zeros(4)
    out[i] = 0.25 * s
    out[j] = (R[j, i] + R[i, j]) / s
    out[k] = (R[k, i] + R[i, k]) / s
    out[3] = (R[k, j] - R[j, k]) / s
    return out


# ------------------------------------------------------------------ metrics
def diff(pb, q1, q2):
    """ArmSlamCalib::Diff (:1230-1242): q1 - q2, wrapped to (-pi, pi] on continuous joints."""
    out = np.asarray(q1, dtype=np.float64) - np.asarray(q2, dtype=np.float64)
    if pb.continuous is not None:
        c = np.asarray(pb.continuous, dtype=bool)
        out = np.where(c, np.arctan2(np.sin(out), np.cos(out)), out)
    return out


def compute_error(truth, current):
    """ArmSlamCalib::ComputeError (:1360-1412).  truth / current = (q [P][N], l [L][3], x12): mean landmark distance, mean norm of the
    joint-angle difference per configuration, distance between the extrinsic translations."""
    (qt, lt, xt), (q, l, x) = truth, current
    lt, l = np.asarray(lt).reshape(-1, 3), np.asarray(l).reshape(-1, 3)
    out = dict(landmarkError=0.0, extrinsicError=0.0, jointAngleError=0.0, cameraPoseError=0.0)
    if len(l):
        out["landmarkError"] = float(np.linalg.norm(lt - l, axis=1).mean())
    if len(q):
        out["jointAngleError"] = float(np.linalg.norm(np.asarray(qt) - np.asarray(q), axis=1).mean())
    out["extrinsicError"] = float(np.linalg.norm(np.asarray(xt).reshape(3, 4)[:, 3] - np.asarray(x).reshape(3, 4)[:, 3]))
    return out


def reprojection_errors(pb, q, l, x12):
    """ArmSlamCalib::CalculateCurrentErrors (:2061-2076): |evaluateError| of every projection factor at the given values, in pixels
    (unwhitened; a point behind the z < 0.01 guard counts as the factor does, with the projection (2 fx, 2 fx),
    RobotProjectionFactor.h:213-220)."""
    q, l = np.asarray(q, dtype=np.float64), np.asarray(l, dtype=np.float64).reshape(-1, 3)
    cams = {}
    out = np.empty(pb.n_obs)
    for m in range(pb.n_obs):
        p = int(pb.pose_idx[m])
        if p not in cams:
            cams[p] = np.linalg.inv(camera_pose(pb, q[p], x12))
        pc = cams[p] @ np.append(l[pb.lm_idx[m]], 1.0)
        if pc[2] < 0.01:
            uv = np.array([2 * pb.fx, 2 * pb.fx])
        else:
            uv = np.array([pb.fx * pc[0] / pc[2] + pb.cx, pb.fy * pc[1] / pc[2] + pb.cy])
        out[m] = np.linalg.norm(uv - pb.uv[m])
    return out


def median_reprojection_error(pb, q, l, x12):
    """The driver's statistic: sorted errors, element size/2 (sim_genrobot.cpp:128-134); None with three factors or fewer."""
    e = np.sort(reprojection_errors(pb, q, l, x12))
    return float(e[len(e) // 2]) if len(e) > 3 else None


def latest_joint_angle_offsets(pb, q_current, q_initial, t):
    """ComputeLatestJointAngleOffsets (:740-765): Diff(current q_t, initial q_t); t == 0 means the last configuration."""
    if t == 0:
        t = len(q_current) - 1
    return diff(pb, q_current[t], q_initial[t])


def latest_ground_truth_sim_offsets(pb, trajectory, encoders, t):
    """ComputeLatestGroundTruthSimOffsets (:723-738): Diff(true q_t, encoder reading at t); t == 0 means the last step."""
    if t == 0:
        t = len(trajectory) - 1
    return diff(pb, trajectory[t], encoders[t])


# ------------------------------------------------------------------ files
def _row(values):
    return " ".join("%.6g" % v for v in values)   # std::ofstream default formatting: 6 significant digits


def print_sim_errors(directory, postfix, pb, q, x12, trajectory, encoders, sim_extrinsic):
    """ArmSlamCalib::PrintSimErrors (:2216-2294): joint_error, extrinsic_error, position_error, kinematic_error `<postfix>.txt`.
    Quirk kept: the quaternion written next to the SIMULATED extrinsic translation is the CURRENT estimate's (:2225)."""
    q = np.asarray(q, dtype=np.float64)
    X, Xs = _T(x12), _T(sim_extrinsic)
    cur_q = quaternion_xyzw(X[:3, :3])
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "extrinsic_error%s.txt" % postfix), "w") as f:
        f.write(_row(list(Xs[:3, 3]) + list(cur_q)) + " " + _row(list(X[:3, 3]) + list(cur_q)) + " ")
    with open(os.path.join(directory, "joint_error%s.txt" % postfix), "w") as fj, \
            open(os.path.join(directory, "position_error%s.txt" % postfix), "w") as fp, \
            open(os.path.join(directory, "kinematic_error%s.txt" % postfix), "w") as fk:
        for i in range(len(q)):
            gt, enc, qi = trajectory[i], encoders[i], q[i]
            fj.write(_row(gt) + " " + _row(qi) + " " + _row(enc) + "\n")
            pose_gt = camera_pose(pb, gt, sim_extrinsic)[:3, 3]
            pose_i = camera_pose(pb, qi, x12)[:3, 3]
            fp.write(_row(list(pose_gt) + list(pose_i)) + "\n")
            kin_i = camera_pose(pb, qi, sim_extrinsic)[:3, 3]
            kin_enc = camera_pose(pb, enc, sim_extrinsic)[:3, 3]
            fk.write(_row(list(pose_gt) + list(kin_i) + list(kin_enc)) + "\n")


class RunLog:
    """The four per-step streams sim_genrobot writes (sim_genrobot.cpp:86-175), file names included (`offsets.txt<postfix>.txt`, ...):
    joint offsets (estimate - initial | truth - encoders), median reprojection error, extrinsic (simulated | current, as
    translation + xyzw quaternion), and `landmarkError extrinsicError jointAngleError` of ComputeError."""

    def __init__(self, directory, postfix=""):
        os.makedirs(directory, exist_ok=True)
        self.offsets = open(os.path.join(directory, "offsets.txt%s.txt" % postfix), "w")
        self.reproj = open(os.path.join(directory, "reproj_error.txt%s.txt" % postfix), "w")
        self.extrinsic = open(os.path.join(directory, "extrinsic_errors.txt%s.txt" % postfix), "w")
        self.error = open(os.path.join(directory, "error%s.txt" % postfix), "w")

    def step(self, pb, t, q, l, x12, q_initial, trajectory, encoders, truth, sim_extrinsic):
        """One driver iteration after (an optional) OptimizeStep: values (q, l, x12) = currentEstimate, truth = (q, l, x12)."""
        a = latest_joint_angle_offsets(pb, q, q_initial, t)
        b = latest_ground_truth_sim_offsets(pb, trajectory, encoders, t)
        self.offsets.write(_row(a[:6]) + " " + _row(b[:6]) + "\n")     # the driver writes six joints (:113-129)
        med = median_reprojection_error(pb, q, l, x12)
        if med is not None:
            self.reproj.write("%.6g\n" % med)
        X, Xs = _T(x12), _T(sim_extrinsic)
        self.extrinsic.write(_row(list(Xs[:3, 3]) + list(quaternion_xyzw(Xs[:3, :3]))) + " " + _row(list(X[:3, 3]) + list(quaternion_xyzw(X[:3, :3]))) + "\n")
        e = compute_error(truth, (q, l, x12))
        self.error.write("%.6g %.6g %.6g\n" % (e["landmarkError"], e["extrinsicError"], e["jointAngleError"]))

    def close(self):
        for f in (self.offsets, self.reproj, self.extrinsic, self.error):
            f.close()
