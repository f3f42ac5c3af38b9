"""Result metrics / files in the reference's text formats (arm_slam_calib_b200/results.py; src/ArmSlamCalib.cpp:1360-1412, :2061-2076,
:2216-2294; src/sim_genrobot.cpp:86-175).  Host-only: the kinematics are checked against the oracle's."""
import os

import numpy as np

from conftest import perturbed_values


def test_host_kinematics_and_reprojection_match_the_oracle(asc, oracle_cls, small_problem):
    from arm_slam_calib_b200 import results as R
    pb = small_problem
    orc = oracle_cls(pb, asc.default_params())
    q, l, x = perturbed_values(pb)
    for p in (0, pb.n_poses // 2, pb.n_poses - 1):
        t_link, _, _ = orc.fk(q[p])
        assert np.abs(R.link_pose(pb, q[p])[:3, :] - t_link.reshape(3, 4)).max() < 1e-12
    e = R.reprojection_errors(pb, q, l, x)
    for m in range(0, pb.n_obs, max(1, pb.n_obs // 50)):
        uv, ch = orc.project(q[pb.pose_idx[m]], l[pb.lm_idx[m]], x, jac=False)
        assert abs(e[m] - np.linalg.norm(uv - pb.uv[m])) < 1e-8
    assert R.median_reprojection_error(pb, q, l, x) == np.sort(e)[len(e) // 2]
    # exact projections at the truth (ArmSlamCalib.cpp:1784-1800): the median error vanishes there
    truth = (pb.meta["q_truth"], pb.meta["lm_truth"], pb.meta["sim_extrinsic"])
    assert R.median_reprojection_error(pb, *truth) < 1e-9


def test_compute_error_and_quaternion(asc, small_problem):
    from arm_slam_calib_b200 import results as R
    pb = small_problem
    truth = (pb.meta["q_truth"], pb.meta["lm_truth"], pb.meta["sim_extrinsic"])
    cur = (pb.q0, pb.l0 + np.array([0.0, 3.0, 4.0]), pb.x0)
    e = R.compute_error(truth, cur)
    assert abs(e["landmarkError"] - 5.0) < 1e-12
    assert abs(e["jointAngleError"] - np.linalg.norm(pb.meta["q_truth"] - pb.q0, axis=1).mean()) < 1e-15
    assert abs(e["extrinsicError"] - np.linalg.norm(pb.meta["sim_extrinsic"].reshape(3, 4)[:, 3] - pb.x0.reshape(3, 4)[:, 3])) < 1e-15
    for ax, th in (([0, 0, 1], 0.3), ([1, 0, 0], 3.0), ([0, 1, 0], -2.5), ([0.6, 0.0, 0.8], 3.1)):
        Rm = R._rot(ax, th)
        x, y, z, w = R.quaternion_xyzw(Rm)
        assert abs(x * x + y * y + z * z + w * w - 1) < 1e-12
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(Rq - Rm).max() < 1e-12


def test_files_have_the_reference_layout(asc, small_problem, tmp_path):
    """What data/plot_*.py read with numpy.loadtxt: column counts of PrintSimErrors and of the driver's per-step streams."""
    from arm_slam_calib_b200 import results as R
    pb = small_problem
    N, P = pb.n_dof, pb.n_poses
    q, l, x = perturbed_values(pb)
    qt, lt, xs = pb.meta["q_truth"], pb.meta["lm_truth"], pb.meta["sim_extrinsic"]
    d = str(tmp_path)
    R.print_sim_errors(d, "_7", pb, q, x, qt, pb.enc, xs)
    joints = np.loadtxt(os.path.join(d, "joint_error_7.txt"))
    assert joints.shape == (P, 3 * N)                       # truth | estimate | encoders (:2247-2267)
    assert np.abs(joints[:, :N] - qt).max() < 1e-5 and np.abs(joints[:, N:2 * N] - q).max() < 1e-5 and np.abs(joints[:, 2 * N:] - pb.enc).max() < 1e-5
    assert np.loadtxt(os.path.join(d, "position_error_7.txt")).shape == (P, 6)
    kin = np.loadtxt(os.path.join(d, "kinematic_error_7.txt"))
    assert kin.shape == (P, 9)
    ext = np.loadtxt(os.path.join(d, "extrinsic_error_7.txt"))
    assert ext.shape == (14,)                               # sim t, q | current t, q (:2223-2239)
    assert np.abs(ext[7:10] - x.reshape(3, 4)[:, 3]).max() < 1e-5
    # plot_dof_experiments.py:75-76 reads the joint file as  median |data[:, 0:d] - data[:, d:2d]|
    assert np.isfinite(np.median(np.abs(joints[:, 0:N] - joints[:, N:2 * N])))
    log = R.RunLog(d, "_7")
    for t in (2, 4):
        log.step(pb, t, q, l, x, pb.q0, qt, pb.enc, (qt, lt, xs), xs)
    log.close()
    assert np.loadtxt(os.path.join(d, "offsets.txt_7.txt")).shape == (2, 12)
    assert np.loadtxt(os.path.join(d, "reproj_error.txt_7.txt")).shape == (2,)
    assert np.loadtxt(os.path.join(d, "extrinsic_errors.txt_7.txt")).shape == (2, 14)
    assert np.loadtxt(os.path.join(d, "error_7.txt")).shape == (2, 3)
