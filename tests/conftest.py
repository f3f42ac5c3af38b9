import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def asc():
    import __graft_entry__ as g
    g.build()
    import arm_slam_calib_b200 as a
    return a


@pytest.fixture(scope="session")
def oracle_cls(asc):
    from oracle import Oracle
    return Oracle


@pytest.fixture(scope="session")
def c1_problem(asc):
    """BASELINE config C1: Mico-like 6-DOF fixture, 250 configurations, reference defaults."""
    return asc.make_sim_problem(trajectory_size=250, seed=0)


@pytest.fixture(scope="session")
def c1_offset_problem(asc):
    """C1 with sim_genrobot's 0.1 m extrinsic guess offset so the extrinsic moves (sim_genrobot.cpp:67-68)."""
    from arm_slam_calib_b200.sim import pose12
    return asc.make_sim_problem(trajectory_size=250, seed=0, extrinsic_guess=pose12(t=[0.1, 0.1, 0.1]))


@pytest.fixture(scope="session")
def small_problem(asc):
    return asc.make_sim_problem(trajectory_size=40, seed=3)


def perturbed_values(pb, seed=1, sq=0.02, sl=0.05, sx=0.02):
    """Values away from the initial estimate so every factor type has a non-trivial residual."""
    from oracle import Oracle
    rng = np.random.default_rng(seed)
    q = pb.q0 + sq * rng.standard_normal(pb.q0.shape)
    l = pb.l0 + sl * rng.standard_normal(pb.l0.shape)
    x = Oracle.pose3_retract(0, pb.x0, sx * rng.standard_normal(6))
    return q, l, x
