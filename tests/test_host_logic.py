"""Host-side logic: the synthetic generator (association/indexing), the landmark partitioner, the mode switch."""
import numpy as np
import pytest


def test_generator_is_deterministic_and_follows_the_admission_rule(asc):
    a = asc.make_sim_problem(trajectory_size=120, seed=4)
    b = asc.make_sim_problem(trajectory_size=120, seed=4)
    for k in ("pose_idx", "lm_idx", "uv", "enc"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    c = asc.make_sim_problem(trajectory_size=120, seed=5)
    assert not np.array_equal(a.enc, c.enc)
    # a landmark is in the graph only once it has > 3 observations (ArmSlamCalib.cpp:1860)
    counts = np.bincount(a.lm_idx, minlength=a.n_landmarks)
    assert counts.min() >= 4
    # compact landmark index follows the reference id order (Values iterate by key)
    assert np.all(np.diff(a.meta["lm_id"]) > 0)
    # factor order: strictly increasing positions in the reference graph, #0 is the extrinsic prior
    fo = a.meta["factor_order"]
    assert fo.min() >= 2 and np.all(np.diff(fo) > 0)
    # inside one landmark observations arrive in time order
    for j in range(min(a.n_landmarks, 20)):
        assert np.all(np.diff(a.pose_idx[a.lm_idx == j]) >= 0)


def test_observations_are_exact_projections_of_the_truth(asc, oracle_cls):
    """GetNewLandmarksSimulated projects the true landmark through the true camera, no pixel noise (ArmSlamCalib.cpp:1784-1800)."""
    pb = asc.make_sim_problem(trajectory_size=60, seed=1)
    orc = oracle_cls(pb, asc.default_params())
    rng = np.random.default_rng(0)
    for m in rng.choice(pb.n_obs, size=25, replace=False):
        uv, ch = orc.project(pb.meta["q_truth"][pb.pose_idx[m]], pb.meta["lm_truth"][pb.lm_idx[m]], pb.meta["sim_extrinsic"], jac=False)
        assert ch == 0 and np.abs(uv - pb.uv[m]).max() < 1e-9
        assert 0 < pb.uv[m][0] < 2 * pb.cx and 0 < pb.uv[m][1] < 2 * pb.cy      # visibility gate :1786


def test_generated_chain_matches_robot_generator_rules(asc):
    """RobotGenerator.cpp:88-129: axis Z, first joint rotated -1.57 about Y with no offset, last link twisted 1.57 about X,
    link scale recurrence max(0.7 s + U(-0.01, 0.01), 0.01)."""
    t_p2j, axis, t_c2j = asc.chain_generate(9, seed=0)
    assert np.array_equal(axis, np.tile([0.0, 0.0, 1.0], (9, 1)))
    T0 = t_p2j[0].reshape(3, 4)
    assert np.allclose(T0[:, 3], 0) and np.isclose(T0[0, 0], np.cos(-1.57)) and np.isclose(T0[0, 2], np.sin(-1.57))
    Tl = t_c2j[-1].reshape(3, 4)
    assert np.isclose(Tl[1, 1], np.cos(1.57)) and np.isclose(Tl[2, 1], np.sin(1.57))
    assert np.isclose(t_c2j[0].reshape(3, 4)[0, 3], -0.25)     # -scale * 0.5
    lens = [-4.0 * t_c2j[i].reshape(3, 4)[0, 3] for i in range(1, 9)]   # lastLinklength per joint
    assert np.isclose(lens[0], 0.5)
    for a, b in zip(lens, lens[1:]):
        assert abs(b - 0.7 * a) <= 0.0100001
    # same seed -> same chain; rand() stream advances otherwise
    t2, _, _ = asc.chain_generate(9, seed=0)
    assert np.array_equal(t_p2j, t2)


def test_host_fk_matches_oracle_fk(asc, oracle_cls):
    chain = asc.chain_generate(8, seed=3)
    pb = asc.make_sim_problem(chain=chain, trajectory_size=8, seed=3, srand=False)
    orc = oracle_cls(pb, asc.default_params())
    q = np.linspace(-1.0, 1.0, 8)
    out = np.zeros(12)
    rc = asc.lib().asc_chain_fk(8, pb.t_p2j.ctypes.data, pb.axis.ctypes.data, pb.t_c2j.ctypes.data, q.ctypes.data, out.ctypes.data)
    assert rc == 0
    assert np.abs(out - orc.fk(q)[0]).max() < 1e-13


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_landmark_partition_is_contiguous_complete_and_balanced(asc, world):
    from arm_slam_calib_b200.capi import shard_problem
    pb = asc.make_sim_problem(trajectory_size=150, seed=2)
    bounds = asc.shard_by_landmark(pb.n_landmarks, pb.lm_idx, world)
    assert bounds[0] == 0 and bounds[-1] == pb.n_landmarks and np.all(np.diff(bounds) >= 0)
    seen = np.zeros(pb.n_obs, bool)
    loads = []
    for r in range(world):
        sh = shard_problem(pb, r, world)
        sel = sh.meta["obs_sel"]
        assert not (seen & sel).any()
        seen |= sel
        loads.append(sel.sum())
        assert sh.lm_idx.min(initial=0) >= 0 and (sh.n_obs == 0 or sh.lm_idx.max() < sh.n_landmarks)
        assert np.array_equal(sh.enc, pb.enc)            # poses replicated
    assert seen.all()
    if world > 1:
        assert max(loads) <= 1.5 * pb.n_obs / world + np.bincount(pb.lm_idx).max()


def test_empty_and_ragged_inputs_are_rejected_or_accepted_like_the_packer_says(asc):
    with pytest.raises(asc.AscError):
        asc.shard_by_landmark(3, np.array([0, 5], np.int32), 2)     # out-of-range landmark index
    b = asc.shard_by_landmark(0, np.zeros(0, np.int32), 4)           # empty graph
    assert list(b) == [0, 0, 0, 0, 0]


def test_optimization_mode_strings(asc):
    """ArmSlamCalib.cpp:408-422: code reads "BatchDogLeg", README says "BatchDogleg"; unknown strings keep the previous mode."""
    f = asc.optimization_mode_from_string
    assert f("BatchLM") == asc.ASC_BATCH_LM
    assert f("BatchDogLeg") == asc.ASC_BATCH_DOGLEG and f("BatchDogleg") == asc.ASC_BATCH_DOGLEG
    assert f("nonsense", previous=asc.ASC_BATCH_LM) == asc.ASC_BATCH_LM
    with pytest.raises(NotImplementedError):
        f("ISAM")
