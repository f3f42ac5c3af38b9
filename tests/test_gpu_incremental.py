"""Incremental mode: the reference's drivers call OptimizeStep every second frame on the growing graph with the
previous estimate as the initial values (src/sim_genrobot.cpp:97-104).  asc_append grows the packed graph on the device; every
OptimizeStep must equal (a) a fresh optimiser packed from scratch on the same prefix graph with the same initial values, bitwise, and
(b) the oracle taking the same step from the same warm start.  The early prefix graphs (12-28 poses, ~30 landmarks) are so poorly
conditioned that the oracle's OWN two solvers (dense Cholesky and PCG) differ by up to 2.6e-5 in the per-iteration error there (1e-9
on the full 40-pose graph), so (b) is held to 1e-4 per iteration; (a) is exact."""
TOL = 1e-4
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def prefix_graphs(pb, steps):
    """The graph after `t` simulation steps for every t in `steps`, as index sets into the full problem: observations of poses < t whose
    landmark has been admitted by then (more than 3 observations so far, ArmSlamCalib.cpp:1860); landmarks are numbered in order of
    admission so that ids are stable while the graph grows."""
    lm_first = {}
    out = []
    for t in steps:
        seen = pb.pose_idx < t
        cnt = np.bincount(pb.lm_idx[seen], minlength=pb.n_landmarks)
        for j in np.nonzero(cnt > 3)[0]:
            if j not in lm_first:
                lm_first[int(j)] = len(lm_first)
        obs = np.nonzero(seen & np.isin(pb.lm_idx, list(lm_first)))[0]
        out.append((t, obs, dict(lm_first)))
    return out


def sub_problem(pb, t, obs, lm_map):
    inv = np.array(sorted(lm_map, key=lm_map.get), dtype=np.int64)            # new id -> old id
    remap = np.full(pb.n_landmarks, -1, dtype=np.int32)
    remap[inv] = np.arange(len(inv), dtype=np.int32)
    return dataclasses.replace(pb, pose_idx=pb.pose_idx[obs].copy(), lm_idx=remap[pb.lm_idx[obs]], uv=pb.uv[obs].copy(), enc=pb.enc[:t].copy(),
                               lm_prior=pb.lm_prior[inv].copy(), q0=pb.q0[:t].copy(), l0=pb.l0[inv].copy(), meta={}), inv


def test_incremental_loop_matches_fresh_packing_and_the_oracle(asc, oracle_cls):
    pb = asc.make_sim_problem(trajectory_size=40, seed=3)
    steps = [12, 20, 28, 40]
    graphs = prefix_graphs(pb, steps)
    prm = asc.default_params()
    prm.max_iterations = 6          # every OptimizeStep of the loop: a few LM iterations from the previous estimate
    t0, obs0, map0 = graphs[0]
    sub0, inv0 = sub_problem(pb, t0, obs0, map0)
    inc = asc.BatchOptimizer(sub0, prm)
    q, l, x = inc.optimize(asc.ASC_BATCH_LM)
    ref = oracle_cls(sub0, prm).optimize(sub0.q0, sub0.l0, sub0.x0, mode=0, solver=0)
    for a, b in zip(inc.log, ref["log"]):
        assert abs(a.error - b.error) <= TOL * b.error
    prev_obs, prev_t, prev_nl = obs0, t0, len(inv0)
    for t, obs, lm_map in graphs[1:]:
        sub, inv = sub_problem(pb, t, obs, lm_map)
        # what SimulationStep added since the last OptimizeStep: observations not in the previous graph (in list order), new keyframes, new landmarks
        new_mask = ~np.isin(obs, prev_obs)
        new_obs = obs[new_mask]
        remap = np.full(pb.n_landmarks, -1, dtype=np.int32)
        remap[inv] = np.arange(len(inv), dtype=np.int32)
        n_new_l = len(inv) - prev_nl
        inc.append(pb.pose_idx[new_obs], remap[pb.lm_idx[new_obs]], pb.uv[new_obs], pb.enc[prev_t:t], pb.lm_prior[inv[prev_nl:]], pb.q0[prev_t:t], pb.l0[inv[prev_nl:]])
        # the graph the library now holds = old observations then the new ones
        order = np.concatenate([prev_obs, new_obs])
        held = dataclasses.replace(sub, pose_idx=pb.pose_idx[order].copy(), lm_idx=remap[pb.lm_idx[order]], uv=pb.uv[order].copy())
        q0 = np.concatenate([q, pb.q0[prev_t:t]]); l0 = np.concatenate([l, pb.l0[inv[prev_nl:]]])
        x_start = x.copy()
        fresh = asc.BatchOptimizer(held, prm)
        fresh.set_values(q0, l0, x)
        assert inc.graph_error() == fresh.graph_error()
        fq, fl, fx = fresh.optimize(asc.ASC_BATCH_LM)
        q, l, x = inc.optimize(asc.ASC_BATCH_LM)
        assert np.array_equal(q, fq) and np.array_equal(l, fl) and np.array_equal(x, fx) and inc.error() == fresh.error()
        fresh.close()
        # the oracle takes the same OptimizeStep from the same warm start (the capped, unconverged steps of two implementations drift apart
        # by ~1e-6 per step if each continues from its own estimate; that drift is not what this test is about)
        ref = oracle_cls(held, prm).optimize(q0, l0, x_start, mode=0, solver=0)
        assert inc.iterations() == ref["iterations"]
        for a, b in zip(inc.log, ref["log"]):
            assert a.inner_steps == b.inner_steps and abs(a.error - b.error) <= TOL * b.error, (t, a.iteration, a.error, b.error)
        assert np.abs(q - ref["q"]).max() < 1e-3 and np.abs(x - ref["x"]).max() < 1e-3
        prev_obs, prev_t, prev_nl = order, t, len(inv)
    inc.close()
