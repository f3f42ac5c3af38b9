"""world_size-2 test of the landmark-sharded path on CPU (gloo): each rank linearises only its shard (with the
oracle, since there is no GPU here), rank 0 alone adds the encoder / extrinsic priors, and the all-reduced pose
blocks, K blocks and error must equal the unsharded linearisation -- the exact exchange step the CUDA path does
with ncclAllReduce."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import arm_slam_calib_b200 as asc
    from arm_slam_calib_b200.capi import shard_problem
    from oracle import Oracle
    pb = asc.make_sim_problem(trajectory_size=60, seed=6)
    prm = asc.default_params()
    full = Oracle(pb, prm).linearize(pb.q0, pb.l0, pb.x0)
    sh = shard_problem(pb, rank, world)
    lo, hi = sh.meta["lm_range"]
    if rank != 0:   # priors are added by rank 0 only: give the other ranks no encoder factors and an uninformative K prior
        import dataclasses
        sh = dataclasses.replace(sh, has_enc=np.zeros(pb.n_poses, np.uint8))
        for i in range(6):
            prm.sigma_extrinsic[i] = 1e150
    loc = Oracle(sh, prm).linearize(pb.q0, pb.l0[lo:hi], pb.x0)
    buf = np.concatenate([loc["bpp"].ravel(), loc["bpk"].ravel(), loc["gp"].ravel(), loc["bkk"].ravel(), loc["gk"].ravel(), [loc["error"]]])
    t = torch.from_numpy(buf)
    dist.all_reduce(t)
    ref = np.concatenate([full["bpp"].ravel(), full["bpk"].ravel(), full["gp"].ravel(), full["bkk"].ravel(), full["gk"].ravel(), [full["error"]]])
    err = float(np.abs(t.numpy() - ref).max() / np.abs(ref).max())
    # landmark blocks are local to the shard and must equal the matching slice
    lerr = float(max(np.abs(loc["c"] - full["c"][lo:hi]).max(initial=0.0), np.abs(loc["wk"] - full["wk"][lo:hi]).max(initial=0.0)))
    out.put((rank, err, lerr, hi - lo))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_linearisation_all_reduces_to_the_full_one(asc):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for _, err, lerr, n in res:
        assert err < 1e-12 and lerr < 1e-9 and n > 0
