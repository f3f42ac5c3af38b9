"""torchrun-launched check of the landmark-sharded path: N ranks vs one GPU on the same problem.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py [c1|c2]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs
from arm_slam_calib_b200.capi import shard_problem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
pb = configs.c1(extrinsic_offset=True) if name == "c1" else configs.c2()
prm = asc.default_params(); prm.device = local
if name != "c1":
    prm.max_iterations = 4
n = asc.lib().asc_comm_id_size()
idt = torch.zeros(n, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = np.zeros(n, dtype=np.uint8)
    assert asc.lib().asc_comm_create_id(buf.ctypes.data) == 0
    idt.copy_(torch.from_numpy(buf))
dist.broadcast(idt, 0)
sh = shard_problem(pb, rank, world)
opt = asc.BatchOptimizer(sh, prm, comm=(rank, world, idt.cpu().numpy().tobytes()))
e0 = opt.graph_error()
torch.cuda.synchronize(); dist.barrier()
t = time.time()
q, l, x = opt.optimize(0)
torch.cuda.synchronize()
dt = time.time() - t
log = [(g.iteration, g.inner_steps, g.pcg_iterations, g.error) for g in opt.log]
if rank == 0:
    for g in opt.log:
        print("  [%d GPUs] it %2d pcg %4d lin %.2f ms solve %.2f ms (%.1f us/pcg-it)" % (world, g.iteration, g.pcg_iterations, g.linearize_ms, g.solve_ms, 1e3 * g.solve_ms / max(g.pcg_iterations, 1)), flush=True)
print("rank %d: shard L=%d M=%d  e0 %.9g  %d iterations in %.3fs final %.9g" % (rank, sh.n_landmarks, sh.n_obs, e0, len(log), dt, opt.error()), flush=True)
qs = [torch.zeros_like(torch.from_numpy(q)).cuda() for _ in range(world)]
dist.all_gather(qs, torch.from_numpy(q).cuda())
if rank == 0:
    for r in range(1, world):
        assert torch.equal(qs[0], qs[r]), "replicated poses differ between ranks"
    p1 = asc.default_params(); p1.device = local; p1.max_iterations = prm.max_iterations
    single = asc.BatchOptimizer(pb, p1)
    e1 = single.graph_error()
    t = time.time(); q1, l1, x1 = single.optimize(0); torch.cuda.synchronize(); dt1 = time.time() - t
    for g in single.log:
        print("  [1 GPU] it %2d pcg %4d lin %.2f ms solve %.2f ms (%.1f us/pcg-it)" % (g.iteration, g.pcg_iterations, g.linearize_ms, g.solve_ms, 1e3 * g.solve_ms / max(g.pcg_iterations, 1)))
    print("single GPU: e0 %.9g  %d iterations in %.3fs final %.9g" % (e1, single.iterations(), dt1, single.error()))
    assert abs(e0 - e1) <= 1e-12 * e1
    assert single.iterations() == len(log)
    worst = 0.0
    for a, g in zip(log, single.log):
        assert a[1] == g.inner_steps
        worst = max(worst, abs(a[3] - g.error) / g.error)
    lo, hi = sh.meta["lm_range"]
    print("max rel per-iteration error diff %.2e  |dq| %.2e  |dx| %.2e  |dl| %.2e" % (worst, np.abs(q - q1).max(), np.abs(x - x1).max(), np.abs(l - l1[lo:hi]).max()))
    assert worst < 1e-6 and np.abs(q - q1).max() < 1e-5 and np.abs(x - x1).max() < 1e-5
    print("MGPU_CHECK_OK world=%d speedup %.2fx" % (world, dt1 / dt))
dist.barrier()
opt.close()
dist.destroy_process_group()
