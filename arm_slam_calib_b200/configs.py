"""Named workloads = BASELINE.json `configs` made concrete.  All synthetic."""
import numpy as np

from .sim import chain_generate, chain_mico, make_sim_problem, pose12, rodrigues_y


def c1(seed=0, extrinsic_offset=False):
    """C1: Mico-like 6-DOF, 250 configurations, reference defaults (ArmSlamCalib.h:69-105)."""
    guess = pose12(t=[0.1, 0.1, 0.1]) if extrinsic_offset else None
    return make_sim_problem(trajectory_size=250, seed=seed, extrinsic_guess=guess)


def c2(seed=0, scale=1.0):
    """C2 / C3: Mico chain, 10k poses, ~100k landmarks, ~2M observations.
    [NEW] knobs: dense landmark cloud around the arm, visibility range 0.27 m, and the reference's
    random walk restarted every 10 steps so 10k poses keep exploring.  `scale` multiplies the pose count and
    the landmark count together (C5 = scale 20) and shrinks the visibility range by scale^(-1/3), so that the number of
    observations per pose (~200) and per landmark (~20) stays what it is at scale 1: the arm's workspace does not grow, the
    landmark cloud gets denser and the camera sees a smaller piece of it."""
    n = int(round(900 * np.sqrt(scale)))
    return make_sim_problem(trajectory_size=int(round(10000 * scale)), nx=n, ny=n, size_x=1.3, size_y=1.3, z_range=(-0.9, 1.6),
                            max_range=0.27 * float(scale) ** (-1.0 / 3.0), segment_length=10, seed=seed, extrinsic_guess=pose12(t=[0.02, 0.02, 0.02]))


def c4(n_dof=7, seed=0, n_poses=50000, landmarks_per_side=2000):
    """C4: sim_genrobot random serial chain (sim_genrobot.cpp:57-81): GenerateRobot(N,0.5,0.7,0.01,0.1,false,6),
    sigma_enc 0.03, true extrinsic rodrigues(0,1.57,0), guess offset 0.1 m.  [NEW] landmark cloud / visibility range tuned once so that
    the graph has the sizes BASELINE.json names: 50k poses, ~480k landmarks, ~10.2M observations (203 per pose, 21 per landmark).
    With 50k poses looking into the same metre-sized workspace that needs a 0.2 m visibility range, so the reference's 0.1 m extrinsic
    guess offset (2 % of ITS 5 m deep scene) would put a fifth of the landmarks behind the guessed camera (measured: 2.2M of 10.2M
    factors hit the cheirality guard and LM rejects every step); the offset is scaled with the scene to 0.02 m, as in C2."""
    chain = chain_generate(n_dof, seed=seed)
    R = rodrigues_y(1.57)
    return make_sim_problem(chain=chain, trajectory_size=n_poses, nx=landmarks_per_side, ny=landmarks_per_side, size_x=1.6, size_y=1.6,
                            z_range=(-1.6, 1.6), max_range=0.2, segment_length=10, encoder_noise=0.03, seed=seed, srand=False,
                            sim_extrinsic=pose12(R=R), extrinsic_guess=pose12(R=R, t=[0.02, 0.02, 0.02]))


def c5(seed=0):
    """C5: the C2 generator at 20x: 200k poses, ~2M landmarks, ~40M observations."""
    return c2(seed=seed, scale=20.0)


def params_for(name, asc_default_params):
    p = asc_default_params()
    if name == "c4":
        for i in range(len(p.sigma_encoder)):
            p.sigma_encoder[i] = 0.03   # sim_genrobot.cpp:63
    return p
