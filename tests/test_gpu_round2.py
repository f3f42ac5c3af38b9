"""Round-2 parity rows (VERDICT r1 "next round" item 1): per-iteration GPU-vs-oracle comparison at the benchmark size for both
optimisers, the EncoderFactor variants, the Pose3 charts, the multilevel coarse solve against the exact reduced system, and the
multi-rank path as a test.

Tolerances: per-iteration total error 1e-6 relative, final joint angles / extrinsic 1e-5 (north star); at C2/C3 size the oracle
solves by PCG too (an exact dense Cholesky of a 60 006-dim reduced system is out of reach on the CPU), with its own -- different --
preconditioner, so the comparison checks the operators and the controllers, not the preconditioner."""
import dataclasses
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import perturbed_values

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def c2(asc):
    from arm_slam_calib_b200 import configs
    return configs.c2()


def _compare_logs(gpu_log, ref_log, tol, first=None):
    if first is None:
        assert len(gpu_log) == len(ref_log)
    else:
        assert len(gpu_log) >= first and len(ref_log) >= first
        gpu_log, ref_log = gpu_log[:first], ref_log[:first]
    for a, b in zip(gpu_log, ref_log):
        assert a.inner_steps == b.inner_steps and a.accepted == b.accepted, (a.iteration, a.inner_steps, b.inner_steps)
        assert abs(a.error - b.error) <= tol * b.error, (a.iteration, a.error, b.error)
        assert abs(a.lambda_or_delta - b.lambda_or_delta) <= 1e-9 * b.lambda_or_delta


def test_c2_first_lm_iterations_match_the_oracle(asc, oracle_cls, c2):
    """BASELINE configs[1] (10k poses / 2M observations), BatchLM: the first three LM iterations -- the benchmark's step --
    against the CPU oracle on the same arrays, per iteration to 1e-6 relative."""
    prm = asc.default_params()
    prm.max_iterations = 3
    opt = asc.BatchOptimizer(c2, prm)
    q, l, x = opt.optimize(asc.ASC_BATCH_LM)
    ref = oracle_cls(c2, prm).optimize(c2.q0, c2.l0, c2.x0, mode=0, solver=1)
    print("C2 LM   gpu errors %s\n        cpu errors %s" % ([g.error for g in opt.log], [g.error for g in ref["log"]]))
    print("        gpu pcg %s  cpu pcg %s" % ([g.pcg_iterations for g in opt.log], [g.pcg_iterations for g in ref["log"]]))
    _compare_logs(opt.log, ref["log"], 1e-6)
    assert np.abs(q - ref["q"]).max() < 1e-5 and np.abs(x - ref["x"]).max() < 1e-5
    opt.close()


def test_c3_first_dogleg_iterations_match_the_oracle(asc, oracle_cls, c2):
    """BASELINE configs[2] (the same arrays, BatchDogleg): three Dogleg iterations per iteration to 1e-6.  The oracle evaluates
    M(delta) as 0.5 |A delta - b|^2 over all factors; the GPU uses the Gauss-Newton identities n^T H n = n^T g, g^T H n = g^T g --
    agreement of error AND Delta at size shows rho and the trust-region updates are the same."""
    prm = asc.default_params()
    prm.max_iterations = 3
    prm.pcg_max_iters = 20000      # undamped Gauss-Newton system: the oracle's cluster-Jacobi PCG needs more than the default 2000
    opt = asc.BatchOptimizer(c2, prm)
    q, l, x = opt.optimize(asc.ASC_BATCH_DOGLEG)
    ref = oracle_cls(c2, prm).optimize(c2.q0, c2.l0, c2.x0, mode=1, solver=1)
    print("C3 DL   gpu errors %s deltas %s\n        cpu errors %s deltas %s" % ([g.error for g in opt.log], [g.lambda_or_delta for g in opt.log],
                                                                            [g.error for g in ref["log"]], [g.lambda_or_delta for g in ref["log"]]))
    _compare_logs(opt.log, ref["log"], 1e-6)
    assert np.abs(q - ref["q"]).max() < 1e-5 and np.abs(x - ref["x"]).max() < 1e-5
    opt.close()


def _check_against_oracle(asc, oracle_cls, pb, prm, modes=(0, 1), blocks=True, tol=1e-6, first=None):
    opt = asc.BatchOptimizer(pb, prm)
    orc = oracle_cls(pb, prm)
    vals = perturbed_values(pb)
    opt.set_values(*vals)
    e_gpu, e_cpu = opt.graph_error(), orc.error(*vals)
    assert abs(e_gpu - e_cpu) <= 1e-11 * abs(e_cpu), (e_gpu, e_cpu)
    if blocks:
        o = orc.linearize(*vals)
        st = opt.linearize()
        assert abs(st.error - o["error"]) <= 1e-11 * o["error"]
        bpp, bpk, gp, bkk, gk = opt.pose_blocks()
        for a, b in ((bpp, o["bpp"]), (gp, o["gp"]), (bkk, o["bkk"]), (gk, o["gk"])):
            assert np.abs(a - b).max() <= 1e-9 * max(np.abs(b).max(), 1e-300)
        rc, dq, dl, dx, _ = orc.solve_damped(*vals, 1e-3, solver=0)
        gq, gl_, gx, its = opt.solve_damped(1e-3)
        s = np.abs(dq).max()
        assert rc == 0 and its > 0 and np.abs(gq - dq).max() < 1e-7 * s and np.abs(gx - dx).max() < 1e-7 * s
    for mode in modes:
        opt.set_values(pb.q0, pb.l0, pb.x0)
        q, l, x = opt.optimize(mode)
        ref = orc.optimize(pb.q0, pb.l0, pb.x0, mode=mode, solver=0)
        _compare_logs(opt.log, ref["log"], tol, first)
        if first is None:
            assert np.abs(q - ref["q"]).max() < 1e-5 and np.abs(x - ref["x"]).max() < 1e-5
    opt.close()
    return e_cpu


def test_continuous_joints_wrap_like_the_reference(asc, oracle_cls, small_problem):
    """EncoderFactor on continuous joints: atan2(sin(q - e), cos(q - e)) (EncoderFactor.h:76-84, RobotConfig.h:72-77).  Encoder readings
    shifted by whole turns give the same residual only when the joint is flagged continuous."""
    pb = small_problem
    cont = np.array([1, 0, 1, 0, 0, 1], dtype=np.uint8)
    enc = pb.enc.copy()
    enc[:, 0] += 2 * np.pi
    enc[:, 2] -= 4 * np.pi
    wrapped = dataclasses.replace(pb, continuous=cont, enc=enc)
    prm = asc.default_params()
    e_wrapped = _check_against_oracle(asc, oracle_cls, wrapped, prm)
    e_plain = oracle_cls(pb, prm).error(*perturbed_values(pb))
    assert abs(e_wrapped - e_plain) <= 1e-9 * e_plain          # whole turns vanish on the continuous joints
    unwrapped = dataclasses.replace(pb, enc=enc)                # same readings, joints NOT continuous: a huge encoder residual
    assert oracle_cls(unwrapped, prm).error(*perturbed_values(pb)) > 10 * e_plain
    opt = asc.BatchOptimizer(unwrapped, prm)
    opt.set_values(*perturbed_values(pb))
    assert abs(opt.graph_error() - oracle_cls(unwrapped, prm).error(*perturbed_values(pb))) <= 1e-11 * opt.graph_error()
    opt.close()


def test_dead_band_encoder_cost(asc, oracle_cls, small_problem):
    """useDeadBand (EncoderFactor.h:40-57, ArmSlamCalib.h:94,96): |d| < band -> 0, else shrunk by the band; J stays I (:99)."""
    prm = asc.default_params()
    prm.use_dead_band = 1
    prm.dead_band = 0.05
    # The dead band makes the cost non-smooth while the Jacobian stays I (EncoderFactor.h:99), so LM's accept/reject path is touchy:
    # error, blocks and the exact damped step are held to the usual bars; the optimiser runs are compared per iteration to 1e-5
    # (measured 1.3e-6 at iteration 4 between the GPU's PCG and the oracle's Cholesky, the same order as between the oracle's own
    # two solvers).
    e_band = _check_against_oracle(asc, oracle_cls, small_problem, prm, tol=1e-5)
    assert e_band < oracle_cls(small_problem, asc.default_params()).error(*perturbed_values(small_problem))


def test_encoder_factor_only_on_the_first_pose(asc, oracle_cls, small_problem):
    """use_joint_encoders = false: only pose 0 gets an EncoderFactor (ArmSlamCalib.cpp:299-305)."""
    has = np.zeros(small_problem.n_poses, dtype=np.uint8)
    has[0] = 1
    pb = dataclasses.replace(small_problem, has_enc=has)
    # Without encoder priors on poses 1.. the joint offsets are held by the projections alone and the problem has near-gauge directions:
    # LM wanders for 60-100 iterations along them and any two fp64 solvers part ways (the oracle's own Cholesky and PCG paths differ
    # by 4.5e-5 at iteration 3 and 11 % at iteration 7).  Error, blocks and the exact damped step are held to the usual bars; the
    # optimiser is compared over the two iterations that are still deterministic.
    _check_against_oracle(asc, oracle_cls, pb, asc.default_params(), modes=(0,), first=2)


@pytest.mark.parametrize("chart", [1, 2])
def test_pose3_charts(asc, oracle_cls, c1_offset_problem, chart):
    """Rot3 expmap / full Pose3 expmap builds of GTSAM: retract and localCoordinates on the device against the oracle,
    through an optimisation in which the extrinsic moves by 0.1 m."""
    prm = asc.default_params()
    prm.pose3_chart = chart
    _check_against_oracle(asc, oracle_cls, c1_offset_problem, prm, blocks=False)
    # and the prior residual (localCoordinates) away from the prior
    pb = c1_offset_problem
    opt = asc.BatchOptimizer(pb, prm)
    q, l, _ = perturbed_values(pb)
    x = oracle_cls.pose3_retract(chart, pb.x0, np.array([0.3, -0.2, 0.25, 0.1, -0.05, 0.2]))
    opt.set_values(q, l, x)
    e_cpu = oracle_cls(pb, prm).error(q, l, x)
    assert abs(opt.graph_error() - e_cpu) <= 1e-11 * e_cpu
    st = opt.linearize()
    assert abs(st.error - e_cpu) <= 1e-11 * e_cpu
    _, _, _, _, gk = opt.pose_blocks()
    assert np.abs(gk - oracle_cls(pb, prm).linearize(q, l, x)["gk"]).max() <= 1e-9 * np.abs(gk).max()
    opt.close()


def test_multilevel_coarse_solve_is_exact_preconditioning(asc, oracle_cls, c2):
    """The V-cycle only preconditions: the converged damped step at C2 size must not depend on the hierarchy (degree, aggregate
    size, depth) nor on having a coarse level at all.  Also reports the hierarchy and the iteration counts."""
    sols, its_all = [], {}
    variants = {"default": {}, "degree2": dict(pcg_ml_degree=2), "deep": dict(pcg_ml_agg=-4, pcg_ml_top_dim=120), "none": dict(pcg_coarse_levels=0)}
    for name, kw in variants.items():
        prm = asc.default_params()
        for k, v in kw.items():
            setattr(prm, k, v)
        opt = asc.BatchOptimizer(c2, prm)
        info = opt.coarse_info()
        dq, dl, dx, its = opt.solve_damped(1e-5)
        print("C2 damped solve, coarse variant %-8s levels %s degree %d -> %d PCG iterations" % (name, info["levels"], info["degree"], its))
        sols.append((dq, dx)); its_all[name] = its
        opt.close()
    s = np.abs(sols[0][0]).max()
    for dq, dx in sols[1:]:
        assert np.abs(dq - sols[0][0]).max() < 1e-7 * s and np.abs(dx - sols[0][1]).max() < 1e-7 * s
    assert its_all["default"] < its_all["none"] / 2          # the coarse level is worth at least 2x


def test_two_ranks_match_one_gpu(asc):
    """scripts/mgpu_check.py under torchrun on 2 GPUs (landmark-sharded C1 vs one GPU: bitwise-equal replicated poses, per-iteration
    error 1e-6).  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "mgpu_check.py"), "c1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MGPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_packing_does_not_depend_on_the_openmp_team_size(asc, c1_offset_problem):
    """asc_set_problem's parallel counting sorts and incidence builder must give the same packed graph whatever team OpenMP delivers
    (slices were derived from the REQUESTED thread count): the optimiser's result under OMP_NUM_THREADS=1 and under
    OMP_THREAD_LIMIT=3 / OMP_DYNAMIC is bitwise the one of the default run."""
    import json
    script = ("import sys, json; sys.path.insert(0, %r); import arm_slam_calib_b200 as asc\n"
              "from arm_slam_calib_b200.sim import pose12\n"
              "pb = asc.make_sim_problem(trajectory_size=250, seed=0, extrinsic_guess=pose12(t=[0.1, 0.1, 0.1]))\n"
              "opt = asc.BatchOptimizer(pb, asc.default_params()); q, l, x = opt.optimize(0)\n"
              "print(json.dumps([opt.error().hex(), float(q.sum()).hex(), float(l.sum()).hex(), opt.iterations()]))\n" % ROOT)
    outs = []
    for env in ({}, {"OMP_NUM_THREADS": "1"}, {"OMP_THREAD_LIMIT": "3", "OMP_DYNAMIC": "true"}, {"OMP_NUM_THREADS": "5"}):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=e, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert all(o == outs[0] for o in outs[1:]), outs
