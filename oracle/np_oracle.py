"""Independent numpy restatement of the path (TEST INFRASTRUCTURE ONLY), used to pin the C oracle.

Deliberately different in construction from asc_oracle.c so that agreement means something:
  - forward kinematics by homogeneous 4x4 products with scipy-free Rodrigues via matrix exponential series terms
  - Jacobians by central differences of the forward model in the tangent space (no analytic Jacobian code)
  - the damped step from the FULL dense normal equations (no Schur complement, no landmark elimination)
Reference behaviour restated: RobotProjectionFactor.h:196-259 (+ [DEP] pinhole camera), EncoderFactor.h:60-90,
PriorFactor<Pose3>/<Point3> at ArmSlamCalib.cpp:685/:1650, Robust(Cauchy) whitening.
"""
import numpy as np


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def rot(axis, th):
    K = hat(np.asarray(axis, float))
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def T44(m12):
    T = np.eye(4)
    T[:3, :] = np.asarray(m12, float).reshape(3, 4)
    return T


def fk(pb, q):
    T = np.eye(4) if pb.t_base is None else T44(pb.t_base)
    for i in range(pb.n_dof):
        R = np.eye(4)
        R[:3, :3] = rot(pb.axis[i], q[i])
        T = T @ T44(pb.t_p2j[i]) @ R @ np.linalg.inv(T44(pb.t_c2j[i]))
    return T


def cayley(w):
    K = hat(w) / 2.0
    return np.linalg.solve(np.eye(3) - K, np.eye(3) + K)


def retract_pose(x12, xi):
    """GTSAM 3.2 default chart: Rot3 Cayley, translation first order."""
    X = T44(x12)
    O = np.eye(4)
    O[:3, :3] = X[:3, :3] @ cayley(xi[:3])
    O[:3, 3] = X[:3, 3] + X[:3, :3] @ xi[3:]
    return O[:3, :].reshape(12)


def local_pose(x0, x1):
    A, B = T44(x0), T44(x1)
    R = A[:3, :3].T @ B[:3, :3]
    K = (R - np.eye(3)) @ np.linalg.inv(R + np.eye(3))   # = hat(w)/2 for the Cayley map
    w = 2.0 * np.array([K[2, 1], K[0, 2], K[1, 0]])
    return np.concatenate([w, A[:3, :3].T @ (B[:3, 3] - A[:3, 3])])


def project(pb, q, l, x12):
    cam = fk(pb, q) @ T44(x12)
    pc = np.linalg.inv(cam) @ np.append(l, 1.0)
    if pc[2] < 0.01:
        return np.array([2 * pb.fx, 2 * pb.fx]), True
    return np.array([pb.fx * pc[0] / pc[2] + pb.cx, pb.fy * pc[1] / pc[2] + pb.cy]), False


def whitened_residuals(pb, prm, q, l, x12):
    """Stacked whitened + Cauchy-reweighted residual of every factor, order: projections, encoders, landmark priors, K prior."""
    k2 = prm.cauchy_k ** 2
    out = []
    for m in range(pb.n_obs):
        uv, _ = project(pb, q[pb.pose_idx[m]], l[pb.lm_idx[m]], x12)
        r = (uv - pb.uv[m]) / prm.sigma_px
        out.append(np.sqrt(k2 / (k2 + r @ r)) * r)
    for p in range(pb.n_poses):
        if pb.has_enc is None or pb.has_enc[p]:
            d = q[p] - pb.enc[p]
            if pb.continuous is not None:
                c = np.asarray(pb.continuous, bool)
                d = np.where(c, np.arctan2(np.sin(d), np.cos(d)), d)
            out.append(d / np.array(prm.sigma_encoder[:pb.n_dof]))
    for j in range(pb.n_landmarks):
        r = (l[j] - pb.lm_prior[j]) / prm.sigma_landmark
        out.append(np.sqrt(k2 / (k2 + r @ r)) * r)
    out.append(local_pose(pb.x_prior, x12) / np.array(prm.sigma_extrinsic[:6]))
    out.extend(drift_residuals(pb, q))
    return np.concatenate(out)


def drift_residuals(pb, q):
    """DriftFactor (DriftFactor.h:44-61): (q1 - q0) - (q1Encoders - q0Encoders); VelocityFactor (VelocityFactor.h:40-62):
    (q1 - q0) - qDot.  Diagonal sigma; linear, no robust weight."""
    out = []
    for has, meas, sigma in (("has_drift", "drift_meas", "sigma_drift"), ("has_vel", "vel_meas", "sigma_vel")):
        if getattr(pb, has, None) is not None:
            sg = np.broadcast_to(np.asarray(getattr(pb, sigma), dtype=np.float64), (pb.n_dof,))
            for p in range(1, pb.n_poses):
                if getattr(pb, has)[p]:
                    out.append(((q[p] - q[p - 1]) - getattr(pb, meas)[p]) / sg)
    return out


def error(pb, prm, q, l, x12):
    r = whitened_residuals(pb, prm, q, l, x12)
    return 0.5 * float(r @ r)


def frozen_weight_residuals(pb, prm, q, l, x12, q0, l0, x0):
    """Residuals at (q,l,x) with the robust weights frozen at the linearisation point (q0,l0,x0): what
    Robust::WhitenSystem linearises (the weight multiplies A and b but is not differentiated)."""
    k2 = prm.cauchy_k ** 2
    out = []
    for m in range(pb.n_obs):
        uv0, _ = project(pb, q0[pb.pose_idx[m]], l0[pb.lm_idx[m]], x0)
        r0 = (uv0 - pb.uv[m]) / prm.sigma_px
        uv, _ = project(pb, q[pb.pose_idx[m]], l[pb.lm_idx[m]], x12)
        out.append(np.sqrt(k2 / (k2 + r0 @ r0)) * (uv - pb.uv[m]) / prm.sigma_px)
    for p in range(pb.n_poses):
        if pb.has_enc is None or pb.has_enc[p]:
            out.append((q[p] - pb.enc[p]) / np.array(prm.sigma_encoder[:pb.n_dof]))
    for j in range(pb.n_landmarks):
        r0 = (l0[j] - pb.lm_prior[j]) / prm.sigma_landmark
        out.append(np.sqrt(k2 / (k2 + r0 @ r0)) * (l[j] - pb.lm_prior[j]) / prm.sigma_landmark)
    out.append(local_pose(pb.x_prior, x12) / np.array(prm.sigma_extrinsic[:6]))
    out.extend(drift_residuals(pb, q))
    return np.concatenate(out)


def dense_jacobian(pb, prm, q, l, x12, h=1e-6):
    """d(frozen-weight residual)/d(tangent) by central differences; tangent order [q (P*N), l (L*3), K (6)]."""
    P, N, L = pb.n_poses, pb.n_dof, pb.n_landmarks
    n = P * N + L * 3 + 6
    cols = []
    for i in range(n):
        def at(s):
            qq, ll, xx = q.copy(), l.copy(), x12.copy()
            if i < P * N:
                qq.reshape(-1)[i] += s
            elif i < P * N + 3 * L:
                ll.reshape(-1)[i - P * N] += s
            else:
                d = np.zeros(6); d[i - P * N - 3 * L] = s
                xx = retract_pose(x12, d)
            return frozen_weight_residuals(pb, prm, qq, ll, xx, q, l, x12)
        cols.append((at(h) - at(-h)) / (2 * h))
    A = np.stack(cols, axis=1)
    # [DEP] gtsam::PriorFactor<Pose3>::evaluateError returns H = eye(6) (not the derivative of localCoordinates):
    # the one place where the linearisation is knowingly not the true Jacobian
    nd = sum(len(r) for r in drift_residuals(pb, q))   # drift rows come after the K0 prior rows
    k0 = A.shape[0] - nd - 6
    A[k0:k0 + 6, :] = 0.0
    A[k0:k0 + 6, n - 6:] = np.diag(1.0 / np.array(prm.sigma_extrinsic[:6]))
    return A


def damped_step(pb, prm, q, l, x12, lam):
    """(A^T A + lam I) delta = -A^T r from the full dense system."""
    A = dense_jacobian(pb, prm, q, l, x12)
    r = frozen_weight_residuals(pb, prm, q, l, x12, q, l, x12)
    H = A.T @ A + lam * np.eye(A.shape[1])
    d = np.linalg.solve(H, -A.T @ r)
    P, N, L = pb.n_poses, pb.n_dof, pb.n_landmarks
    return d[:P * N].reshape(P, N), d[P * N:P * N + 3 * L].reshape(L, 3), d[P * N + 3 * L:]
