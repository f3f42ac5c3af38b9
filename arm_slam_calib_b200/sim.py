"""Host-side synthetic problems (wraps the C++ generator in csrc/sim_generator.cpp)."""
import ctypes as C

import numpy as np

from . import _ctypes_defs as D
from .capi import Problem, _ptr, _vp
from .capi import hostlib as lib   # generator calls go to the host-only library


def chain_mico():
    t_p2j = np.zeros((6, 12)); axis = np.zeros((6, 3)); t_c2j = np.zeros((6, 12))
    n = lib().asc_chain_mico(_ptr(t_p2j), _ptr(axis), _ptr(t_c2j))
    assert n == 6
    return t_p2j, axis, t_c2j


def chain_generate(n_dof, seed=None, scale=0.5, growth=0.7, random_length=0.01, random_angle=0.1):
    """RobotGenerator::GenerateRobot(n, 0.5, 0.7, 0.01, 0.1, false, 6) as sim_genrobot.cpp:74 calls it.
    seed: srand(seed) first (sim_genrobot.cpp:60); None keeps the libc stream where it is."""
    if seed is not None:
        lib().asc_srand(seed)
    t_p2j = np.zeros((n_dof, 12)); axis = np.zeros((n_dof, 3)); t_c2j = np.zeros((n_dof, 12))
    n = lib().asc_chain_generate(n_dof, scale, growth, random_length, random_angle, _ptr(t_p2j), _ptr(axis), _ptr(t_c2j))
    if n != n_dof:
        raise ValueError("unsupported DOF %d" % n_dof)
    return t_p2j, axis, t_c2j


def rodrigues_y(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def pose12(R=None, t=None):
    m = np.zeros((3, 4))
    m[:, :3] = np.eye(3) if R is None else R
    if t is not None:
        m[:, 3] = t
    return m.reshape(12)


def make_sim_problem(chain=None, trajectory_size=250, nx=25, ny=25, size_x=10.0, size_y=10.0, z_range=(-5.0, 5.0), max_range=0.0,
                     encoder_noise=0.1, seed=0, segment_length=0, sim_extrinsic=None, extrinsic_guess=None, fx=640.0, fy=640.0, cx=320.0, cy=240.0,
                     srand=True):
    """CreateSimulation + SimulationStep over the whole trajectory (src/ArmSlamCalib.cpp:654-706, :307-327),
    one batch graph.  Defaults are ArmSlamCalib::Params (ArmSlamCalib.h:69-105) with trajectory_length=250
    (BASELINE config C1)."""
    L = lib()
    if chain is None:
        chain = chain_mico()
    t_p2j, axis, t_c2j = chain
    n_dof = axis.shape[0]
    cfg = D.AscSimConfig()
    L.asc_sim_default_config(C.byref(cfg))
    cfg.trajectory_size = trajectory_size
    cfg.num_landmarks_x, cfg.num_landmarks_y = nx, ny
    cfg.landmark_size_x, cfg.landmark_size_y = size_x, size_y
    cfg.landmark_z_min, cfg.landmark_z_max = z_range
    cfg.max_range = max_range
    cfg.encoder_noise = encoder_noise
    cfg.seed = seed
    cfg.segment_length = segment_length
    if sim_extrinsic is not None:
        cfg.sim_extrinsic[:] = list(np.asarray(sim_extrinsic, dtype=np.float64).reshape(12))
    if extrinsic_guess is not None:
        cfg.extrinsic_guess[:] = list(np.asarray(extrinsic_guess, dtype=np.float64).reshape(12))
    if srand:
        L.asc_srand(seed)
    s = _vp()
    rc = L.asc_sim_create(C.byref(cfg), n_dof, _ptr(t_p2j), _ptr(axis), _ptr(t_c2j), fx, fy, cx, cy, C.byref(s))
    if rc:
        raise RuntimeError(L.asc_last_error().decode())
    try:
        P, Lm, M, Ls = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int32()
        L.asc_sim_sizes(s, C.byref(P), C.byref(Lm), C.byref(M), C.byref(Ls))
        P, Lm, M = P.value, Lm.value, M.value
        pose_idx = np.zeros(M, np.int32); lm_idx = np.zeros(M, np.int32); uv = np.zeros((M, 2))
        enc = np.zeros((P, n_dof)); q_truth = np.zeros((P, n_dof)); lm_truth = np.zeros((Lm, 3))
        lm_id = np.zeros(Lm, np.int32); order = np.zeros(M, np.int64)
        L.asc_sim_export(s, _ptr(pose_idx), _ptr(lm_idx), _ptr(uv), _ptr(enc), _ptr(q_truth), _ptr(lm_truth), _ptr(lm_id), _ptr(order))
    finally:
        L.asc_sim_destroy(s)
    guess = np.array(list(cfg.extrinsic_guess))
    return Problem(n_dof=n_dof, t_p2j=t_p2j, axis=axis, t_c2j=t_c2j, fx=fx, fy=fy, cx=cx, cy=cy, pose_idx=pose_idx, lm_idx=lm_idx,
                   uv=uv, enc=enc, lm_prior=lm_truth.copy(), x_prior=guess.copy(), q0=enc.copy(), l0=lm_truth.copy(), x0=guess.copy(),
                   meta=dict(q_truth=q_truth, lm_truth=lm_truth, lm_id=lm_id, factor_order=order, n_sim_landmarks=Ls.value,
                             sim_extrinsic=np.array(list(cfg.sim_extrinsic))))


def with_drift(pb, sigma=0.05):
    """The same problem with DriftFactor(q_{p-1}, q_p) between every pair of consecutive configurations, as SimulationStep adds them
    when use_drift_noise is set (src/ArmSlamCalib.cpp:317-320, :1653-1668): measurement = encoders[p] - encoders[p-1],
    Diagonal sigma = drift_noise_level (0.05, ArmSlamCalib.h:84)."""
    import dataclasses
    import numpy as np
    P, N = pb.enc.shape
    hd = np.ones(P, dtype=np.uint8); hd[0] = 0
    meas = np.zeros((P, N)); meas[1:] = pb.enc[1:] - pb.enc[:-1]
    return dataclasses.replace(pb, has_drift=hd, drift_meas=meas, sigma_drift=np.full(N, float(sigma)))


def with_velocity(pb, sigma=0.01, seed=0):
    """The same problem with VelocityFactor(q_{p-1}, q_p, velocities[p-1]) between consecutive configurations, as SimulationStep adds
    them when use_velocity_noise is set (src/ArmSlamCalib.cpp:322-325, :1671-1686).  The measured joint velocity is synthesised as
    the true increment plus N(0, sigma) noise (the reference records qDot + N(0, velocity_noise_level), :588); Diagonal sigma =
    velocity_noise_level (0.01, ArmSlamCalib.h:80)."""
    import dataclasses
    import numpy as np
    P, N = pb.enc.shape
    qt = pb.meta.get("q_truth", pb.enc)
    rng = np.random.default_rng(seed)
    hv = np.ones(P, dtype=np.uint8); hv[0] = 0
    meas = np.zeros((P, N)); meas[1:] = (qt[1:] - qt[:-1]) + rng.normal(0.0, sigma, size=(P - 1, N))
    return dataclasses.replace(pb, has_vel=hv, vel_meas=meas, sigma_vel=np.full(N, float(sigma)))
