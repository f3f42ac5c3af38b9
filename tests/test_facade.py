"""The GTSAM-shaped C++ facade (include/gtsam_facade/gtsam_facade.h): tests/cpp/facade_test.cpp builds the graph through
AddFactor/AddValue-style calls and runs the body of ArmSlamCalib::OptimizeStep (src/ArmSlamCalib.cpp:438-467) unchanged."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def facade_exe(asc, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("facade") / "facade_test")
    libdir = os.path.join(ROOT, "arm_slam_calib_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), "-o", exe, "-L", libdir, "-l:libasc_b200.so",
                    "-Wl,-rpath," + libdir], check=True)
    return exe


def test_facade_compiles_and_fails_loudly_without_a_gpu(facade_exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([facade_exe, "BatchLM", "12"], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode_name,mode", [("BatchLM", 0), ("BatchDogLeg", 1), ("BatchDogleg", 1)])
def test_optimize_step_through_the_facade_matches_the_oracle(asc, oracle_cls, facade_exe, mode_name, mode):
    r = subprocess.run([facade_exe, mode_name, "40"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tok = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0].split()
    iters, err, kx, ky, kz, found = int(tok[1]), float(tok[2]), float(tok[3]), float(tok[4]), float(tok[5]), int(tok[6])
    pb = asc.make_sim_problem(trajectory_size=40, seed=3)
    ref = oracle_cls(pb, asc.default_params()).optimize(pb.q0, pb.l0, pb.x0, mode=mode, solver=0)
    assert iters == ref["iterations"] and found == pb.n_landmarks
    assert abs(err - ref["error"]) <= 1e-6 * ref["error"]
    assert np.abs(np.array([kx, ky, kz]) - ref["x"].reshape(3, 4)[:, 3]).max() < 1e-5


@pytest.mark.gpu
def test_marginals_through_the_facade_match_the_oracle(asc, oracle_cls, facade_exe):
    r = subprocess.run([facade_exe, "BatchLM", "40", "marginal"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    diag = np.array([float(t) for t in [l for l in r.stdout.splitlines() if l.startswith("MARGINAL")][0].split()[1:]])
    pb = asc.make_sim_problem(trajectory_size=40, seed=3)
    orc = oracle_cls(pb, asc.default_params())
    ref = orc.optimize(pb.q0, pb.l0, pb.x0, mode=0, solver=0)
    rc, cov = orc.extrinsic_marginal(ref["q"], ref["l"], ref["x"])
    assert rc == 0
    assert np.abs(diag - np.diag(cov)).max() < 1e-4 * np.abs(np.diag(cov)).max()


@pytest.mark.gpu
def test_drift_factors_through_the_facade_match_the_oracle(asc, oracle_cls, facade_exe):
    """use_drift_noise: DriftFactor between consecutive configurations (src/ArmSlamCalib.cpp:317-320, :1653-1668)."""
    r = subprocess.run([facade_exe, "BatchLM", "40", "drift"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tok = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0].split()
    iters, err = int(tok[1]), float(tok[2])
    pb = asc.with_drift(asc.make_sim_problem(trajectory_size=40, seed=3))
    ref = oracle_cls(pb, asc.default_params()).optimize(pb.q0, pb.l0, pb.x0, mode=0, solver=0)
    assert iters == ref["iterations"]
    assert abs(err - ref["error"]) <= 1e-6 * ref["error"]


@pytest.mark.gpu
def test_remaining_gtsam_surface_through_the_facade(facade_exe):
    """saveGraph (src/ArmSlamCalib.cpp:1910-1913), NonlinearFactorGraph::clone (:2209), Rot3::toQuaternion (:2224), Point3 ops (:1377)."""
    r = subprocess.run([facade_exe, "BatchLM", "20", "dot"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tok = [l for l in r.stdout.splitlines() if l.startswith("DOT")][0].split()
    dot_bytes, n_factors, qnorm, moved = int(tok[1]), int(tok[2]), float(tok[3]), float(tok[4])
    assert dot_bytes > 1000 and n_factors > 20 and abs(qnorm - 1.0) < 1e-6 and moved >= 0.0
