"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Tolerances (stated per north star): factor/landmark association and indexing bit-exact; per-iteration total
error within 1e-6 relative; final joint angles / extrinsic within 1e-5 rad / 1e-5 m.  Per-kernel blocks are held
to 1e-9 relative to the block scale (fp64, different summation order and FMA contraction)."""
import numpy as np
import pytest

from conftest import perturbed_values

pytestmark = pytest.mark.gpu


def rel(a, b):
    s = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / s


@pytest.fixture(scope="module")
def problems(asc, c1_problem, c1_offset_problem, small_problem):
    return dict(c1=c1_problem, c1off=c1_offset_problem, small=small_problem)


@pytest.mark.parametrize("name", ["small", "c1", "c1off"])
def test_error_matches_oracle(asc, oracle_cls, problems, name):
    pb = problems[name]
    prm = asc.default_params()
    opt = asc.BatchOptimizer(pb, prm)
    orc = oracle_cls(pb, prm)
    for vals in [(pb.q0, pb.l0, pb.x0), perturbed_values(pb)]:
        opt.set_values(*vals)
        e_gpu, e_cpu = opt.graph_error(), orc.error(*vals)
        assert abs(e_gpu - e_cpu) <= 1e-11 * abs(e_cpu), (e_gpu, e_cpu)
    opt.close()


@pytest.mark.parametrize("name", ["small", "c1off"])
def test_linearization_blocks_match_oracle(asc, oracle_cls, problems, name):
    pb = problems[name]
    prm = asc.default_params()
    opt = asc.BatchOptimizer(pb, prm)
    orc = oracle_cls(pb, prm)
    vals = perturbed_values(pb)
    opt.set_values(*vals)
    o = orc.linearize(*vals)
    st = opt.linearize()
    assert abs(st.error - o["error"]) <= 1e-11 * o["error"]
    assert st.cheirality == o["cheirality"]
    b, jq, jl = opt.projection_blocks()
    assert rel(b, o["b"]) < 1e-10
    assert rel(jq, o["jq"]) < 1e-10
    assert rel(jl, o["jl"]) < 1e-10
    opt.linearize()
    c, gl, wk = opt.landmark_blocks()
    assert rel(c, o["c"]) < 1e-9 and rel(gl, o["gl"]) < 1e-9 and rel(wk, o["wk"]) < 1e-9
    bpp, bpk, gp, bkk, gk = opt.pose_blocks()
    assert rel(bpp, o["bpp"]) < 1e-9 and rel(bpk, o["bpk"]) < 1e-9 and rel(gp, o["gp"]) < 1e-9
    assert rel(bkk, o["bkk"]) < 1e-9 and rel(gk, o["gk"]) < 1e-9
    opt.close()


def test_cheirality_branch(asc, oracle_cls, small_problem):
    """Landmarks pushed behind the camera hit the z<0.01 guard: J = 0, residual (2fx - z) (RobotProjectionFactor.h:213-220)."""
    pb = small_problem
    prm = asc.default_params()
    opt = asc.BatchOptimizer(pb, prm)
    orc = oracle_cls(pb, prm)
    q, l, x = perturbed_values(pb)
    l = l.copy()
    l[::3] += np.array([0.0, 0.0, 40.0])   # far away: some end up behind some cameras
    l[1::3] -= np.array([30.0, 0.0, 0.0])
    opt.set_values(q, l, x)
    o = orc.linearize(q, l, x)
    assert o["cheirality"] > 0
    st = opt.linearize()
    assert st.cheirality == o["cheirality"]
    assert abs(st.error - o["error"]) <= 1e-11 * o["error"]
    b, jq, jl = opt.projection_blocks()
    assert rel(b, o["b"]) < 1e-10 and rel(jq, o["jq"]) < 1e-10 and rel(jl, o["jl"]) < 1e-10
    opt.close()


@pytest.mark.parametrize("lam", [1e-5, 1e-2, 10.0])
def test_damped_solve_matches_exact_cholesky(asc, oracle_cls, c1_offset_problem, lam):
    pb = c1_offset_problem
    prm = asc.default_params()
    opt = asc.BatchOptimizer(pb, prm)
    orc = oracle_cls(pb, prm)
    vals = perturbed_values(pb)
    opt.set_values(*vals)
    rc, dq, dl, dx, _ = orc.solve_damped(*vals, lam, solver=0)
    assert rc == 0
    gq, gl_, gx, its = opt.solve_damped(lam)
    assert its > 0
    assert rel(gq, dq) < 1e-7 and rel(gl_, dl) < 1e-7 and rel(gx, dx) < 1e-7
    opt.close()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", ["c1", "c1off"])
def test_optimize_matches_oracle_per_iteration(asc, oracle_cls, problems, name, mode):
    pb = problems[name]
    prm = asc.default_params()
    opt = asc.BatchOptimizer(pb, prm)
    orc = oracle_cls(pb, prm)
    ref = orc.optimize(pb.q0, pb.l0, pb.x0, mode=mode, solver=0)
    q, l, x = opt.optimize(mode)
    assert opt.iterations() == ref["iterations"]
    for a, b in zip(opt.log, ref["log"]):
        assert a.inner_steps == b.inner_steps and a.accepted == b.accepted
        assert abs(a.error - b.error) <= 1e-6 * b.error, (a.iteration, a.error, b.error)
        assert abs(a.lambda_or_delta - b.lambda_or_delta) <= 1e-9 * b.lambda_or_delta
    assert abs(opt.error() - ref["error"]) <= 1e-6 * ref["error"]
    assert np.abs(q - ref["q"]).max() < 1e-5          # rad
    assert np.abs(x - ref["x"]).max() < 1e-5          # m / rotation-matrix entries
    assert np.abs(l - ref["l"]).max() < 1e-4
    opt.close()


def test_run_to_run_determinism(asc, c1_offset_problem):
    pb = c1_offset_problem
    res = []
    for _ in range(2):
        opt = asc.BatchOptimizer(pb, asc.default_params())
        q, l, x = opt.optimize(0)
        res.append((q.copy(), l.copy(), x.copy(), opt.error()))
        opt.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert res[0][3] == res[1][3]


def test_repacking_on_a_reused_handle_is_bitwise_identical(asc, c1_offset_problem, small_problem):
    """asc_set_problem reuses the device arena without clearing it: results must not depend on what was there before."""
    prm = asc.default_params()
    fresh = asc.BatchOptimizer(c1_offset_problem, prm)
    q0, l0, x0 = fresh.optimize(0)
    e0 = fresh.error()
    fresh.close()
    opt = asc.BatchOptimizer(small_problem, prm)      # different sizes first
    opt.optimize(1)
    for _ in range(2):
        opt.set_problem(c1_offset_problem)
        opt.set_values(c1_offset_problem.q0, c1_offset_problem.l0, c1_offset_problem.x0)
        q, l, x = opt.optimize(0)
        assert np.array_equal(q, q0) and np.array_equal(l, l0) and np.array_equal(x, x0) and opt.error() == e0
    opt.close()


@pytest.mark.parametrize("name", ["small", "c1off"])
def test_extrinsic_marginal_matches_oracle(asc, oracle_cls, problems, name):
    """asc_extrinsic_marginal (six PCG solves of the undamped reduced system) against the oracle's exact (S^-1)_KK."""
    pb = problems[name]
    prm = asc.default_params()
    opt = asc.BatchOptimizer(pb, prm)
    orc = oracle_cls(pb, prm)
    vals = perturbed_values(pb)
    opt.set_values(*vals)
    cov, its = opt.extrinsic_marginal()
    rc, ref = orc.extrinsic_marginal(*vals)
    assert rc == 0 and its > 0
    assert np.abs(cov - cov.T).max() == 0.0
    assert np.abs(cov - ref).max() < 1e-7 * np.abs(ref).max(), (cov, ref)
    # the marginal does not disturb a following optimize()
    q, l, x = opt.optimize(asc.ASC_BATCH_LM)
    r = orc.optimize(*vals, mode=0, solver=0)
    assert opt.iterations() == r["iterations"] and abs(opt.error() - r["error"]) <= 1e-6 * r["error"]
    opt.close()


@pytest.mark.parametrize("name", ["small", "c1off"])
def test_drift_factors_match_oracle(asc, oracle_cls, problems, name):
    """DriftFactor(q_{p-1}, q_p) links (DriftFactor.h:44-61): error, keyframe blocks, exact damped step, both optimisers."""
    pb = asc.with_drift(problems[name])
    if name == "small":     # both link kinds at once: VelocityFactor (VelocityFactor.h:40-62) on top of the drift links, with a gap
        pb = asc.with_velocity(pb)
        pb.has_vel[3] = 0
    prm = asc.default_params()
    if name == "small":
        prm.max_iterations = 15     # the stiff velocity links make LM crawl towards GTSAM's convergence test; compare 15 iterations
    opt = asc.BatchOptimizer(pb, prm)
    orc = oracle_cls(pb, prm)
    base = oracle_cls(problems[name], prm)
    vals = perturbed_values(pb)
    opt.set_values(*vals)
    e_gpu, e_cpu = opt.graph_error(), orc.error(*vals)
    assert abs(e_gpu - e_cpu) <= 1e-11 * e_cpu
    assert e_cpu > base.error(*vals)           # the links do contribute
    o = orc.linearize(*vals)
    st = opt.linearize()
    assert abs(st.error - o["error"]) <= 1e-11 * o["error"]
    bpp, bpk, gp, bkk, gk = opt.pose_blocks()
    assert rel(bpp, o["bpp"]) < 1e-9 and rel(gp, o["gp"]) < 1e-9
    for lam in (1e-5, 1e-1):
        rc, dq, dl, dx, _ = orc.solve_damped(*vals, lam, solver=0)
        gq, gl_, gx, its = opt.solve_damped(lam)
        s = np.abs(dq).max()
        assert rc == 0 and np.abs(gq - dq).max() < 1e-7 * s and np.abs(gx - dx).max() < 1e-7 * s
    for mode in (asc.ASC_BATCH_LM, asc.ASC_BATCH_DOGLEG):
        opt.set_values(pb.q0, pb.l0, pb.x0)
        q, l, x = opt.optimize(mode)
        r = orc.optimize(pb.q0, pb.l0, pb.x0, mode=mode, solver=0)
        assert opt.iterations() == r["iterations"]
        for a, b in zip(opt.log, r["log"]):
            assert abs(a.error - b.error) <= 1e-6 * b.error
        assert np.abs(q - r["q"]).max() < 1e-5 and np.abs(x - r["x"]).max() < 1e-5
    opt.close()
