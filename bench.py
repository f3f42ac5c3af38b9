#!/usr/bin/env python
"""bench.py -- LM iterations/s of the batch optimisation path on the synthetic Mico problem (BASELINE.json).

One "step" = one full BatchLM optimize() of the workload from its initial estimate to GTSAM's convergence test
(ArmSlamCalib::OptimizeStep, BatchLM branch, src/ArmSlamCalib.cpp:458-467).  `value` = LM iterations / second with
the graph and the initial values already resident in HBM; `e2e` = the same through the C ABI from host buffers
(graph packing + H2D + optimize + D2H every step).  `--impl reference` times the CPU oracle port (the reference's
GTSAM/DART stack cannot be built here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KERNEL_NAMES = {1: "linearize_pose (K1+K2b)", 2: "landmark_blocks (K2a)", 3: "schur_pass_a (K3)", 4: "schur_pass_b (K3)",
                5: "schur_pass_c (K3)", 6: "error_proj (K1e)", 7: "proj_kk (K1b)", 8: "coarse_gemv_f32 (preconditioner)",
                9: "pcg_cluster_update (preconditioner)"}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(kid, N, P, L, M, nz=0, sc=0):
    """Algorithmic bytes one launch must move; fp64 = 8 B, int32 = 4 B.  Streams are counted
    per observation, gathered tables (landmarks, camera rows, u) once at their size."""
    planes = 16 * (N + 3)                       # J_q (2N) + J_l (6) doubles per observation
    if kid == 1:   # uv + landmark index in, planes out; landmark table, FK rows, q/enc in; pose blocks out
        return M * (16 + 4 + planes) + L * 32 + P * 8 * (12 + 6 * N + 2 * N + N * N + 6 * N + N)
    if kid == 2:   # uv + pose index in; camera rows, landmarks, priors in; C, g, W_K out
        return M * (16 + 4) + P * 96 + L * (32 + 24) + L * 8 * (6 + 3 + 18)
    if kid == 3:   # J_l planes + landmark index + slot index in, v (padded 32 B) out; landmark table, FK rows, p, z in, p out
        return M * (48 + 4 + 4 + 32) + L * 32 + P * 8 * (6 * N + 3 * N)
    if kid == 4:   # v segments (one 32-byte slot per observation) + segment pointers in; W_K, Cinv in, u out
        return M * 32 + L * (4 + 8 * (18 + 6) + 32)
    if kid == 5:   # J_l planes + landmark index in; u table + landmark table in; FK rows, B_pp, B_pK, x in, y out
        return M * (48 + 4) + L * 64 + P * 8 * (6 * N + N * N + 6 * N + 2 * N)
    if kid == 8:   # fp32 inverse of the coarse operator in; restriction in, correction out
        return nz * nz * 4 + nz * 16
    if kid == 9:   # cluster blocks in; x, p, r, y in; x, r, z out; restriction out
        return sc * 8 + P * N * 8 * 7
    if kid in (6, 7):   # uv + 2 indices in; camera rows + landmark table in
        return M * (16 + 8) + P * 96 + L * 32
    return 0


LM_ITERATIONS_PER_STEP = 3   # the bounded sample both arms run: first 3 LM iterations of the BatchLM optimize


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  The sampler is started before
    the warm-up steps (nvidia-smi needs ~1 s to deliver its first row) and only rows that arrive inside the timed
    window are kept; when the window is too short to catch two rows, the warm-up rows (same workload) are used and the
    JSON says so."""

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_mark = index, [], None, None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first_row(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        t_end = time.perf_counter()
        time.sleep(0.06)   # let the row that covers the end of the window arrive
        self.proc.terminate()
        rows = [r for t, r in self.rows if self.t_mark is None or self.t_mark <= t <= t_end + 0.06]
        window = "timed region"
        if len(rows) < 2:
            rows, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than two sampling periods)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def make_workload(n_gpus):
    """The workload BASELINE.json's metric is quoted on (configs[1] = C2) at every N: multi-GPU runs shard its landmarks
    (strong scaling).  The preconditioner's dense coarse level stops at 16 384 unknowns, so the
    larger configs C4/C5 are not used as bench lines yet."""
    from arm_slam_calib_b200 import configs
    name = "C2: synthetic Mico 6-DOF, 10k poses / ~90k landmarks / ~2M projection observations, BatchLM"
    if n_gpus > 1:
        name += ", landmarks sharded over %d GPUs" % n_gpus
    return configs.c2(), name


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build()
    import arm_slam_calib_b200 as asc
    from oracle import Oracle
    pb, wname = make_workload(args.gpus)
    prm = asc.default_params()
    prm.max_iterations = LM_ITERATIONS_PER_STEP
    orc = Oracle(pb, prm)
    cores = Oracle.max_threads()
    sample = ("first %d LM iterations of the workload's BatchLM optimize from the initial estimate (per iteration: linearisation of all factors "
              "with per-factor FK, cluster-Jacobi PCG solve to 1e-12, error evaluation); oracle port with OpenMP on all host cores" % LM_ITERATIONS_PER_STEP)

    def step():
        t = time.perf_counter()
        r = orc.optimize(pb.q0, pb.l0, pb.x0, mode=0, solver=1)
        return time.perf_counter() - t, r
    warm = min(args.warmup, 1)   # a CPU step is ~25 s on the box's 16 cores: one untimed pass is all the warm-up it needs
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    its = 0
    for _ in range(args.steps):
        _, r = step()
        its += r["iterations"]
    dt = time.perf_counter() - t0
    val = its / dt
    line = {"impl": "reference", "metric": "LM iterations/s", "value": val, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "warmup_run": warm, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "poses": pb.n_poses, "landmarks": pb.n_landmarks, "observations": pb.n_obs, "dof": pb.n_dof},
            "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample,
                             "pcg_iterations": [int(gg.pcg_iterations) for gg in r["log"]]},
            "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-run", action="store_true",
                    help="warm-up + timed region only (no full optimize, e2e, per-kernel or CPU legs): the command the ncu launch list is taken from")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import arm_slam_calib_b200 as asc
    from arm_slam_calib_b200.capi import shard_problem
    asc.lib()
    pb_full, wname = make_workload(args.gpus)
    prm = asc.default_params()
    prm.device = local_rank
    comm = None
    pb = pb_full
    if world > 1:
        n = asc.lib().asc_comm_id_size()
        idt = torch.zeros(n, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = np.zeros(n, dtype=np.uint8)
            rc = asc.lib().asc_comm_create_id(buf.ctypes.data)
            assert rc == 0, asc.lib().asc_last_error()
            idt.copy_(torch.from_numpy(buf))
        dist.broadcast(idt, 0)
        comm = (rank, world, idt.cpu().numpy().tobytes())
        pb = shard_problem(pb_full, rank, world)
    N, P, L, M = pb.n_dof, pb.n_poses, pb.n_landmarks, pb.n_obs

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    prm.max_iterations = LM_ITERATIONS_PER_STEP
    opt = asc.BatchOptimizer(pb, prm, comm=comm)
    opt.save_values()

    def step():
        opt.restore_values()
        opt.optimize(asc.ASC_BATCH_LM)
        return opt.iterations()

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        step()
    clocks.wait_first_row()
    barrier()
    clocks.mark()
    launches0 = opt.launch_count()
    DOMINANT = 5                # schur_pass_c: largest share of the step among this library's kernels (profiles/README.md)
    opt.profile_begin(DOMINANT, 128)   # its first 128 launches are event-bracketed; the rest of the region runs as CUDA graphs
    opt.timer_start()
    iters = 0
    for _ in range(args.steps):
        iters += step()
    ms = opt.timer_stop()
    barrier()
    clk = clocks.stop()
    n_prof, mean_ms, min_ms = opt.profile_read()
    launches = opt.launch_count() - launches0
    ms = max_over_ranks(ms)
    value = iters / (ms * 1e-3)
    pcg_per_iter = [int(g.pcg_iterations) for g in opt.log]
    final_error = opt.error()
    if args.profile_run:
        if rank == 0:
            print(json.dumps({"metric": "LM iterations/s", "value": value, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms / args.steps, "profile_run": True, "gpu_launches": int(launches)}), flush=True)
        opt.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- full optimize to GTSAM's convergence test, once, for context (not the timed step)
    prm.max_iterations = 100
    opt.set_params(prm)
    opt.restore_values()
    opt.timer_start()
    opt.optimize(asc.ASC_BATCH_LM)
    fms = max_over_ranks(opt.timer_stop())
    full = {"lm_iterations": opt.iterations(), "seconds": fms * 1e-3, "iterations_per_s": opt.iterations() / (fms * 1e-3),
            "final_error": opt.error(), "pcg_iterations": [int(g.pcg_iterations) for g in opt.log]}
    prm.max_iterations = LM_ITERATIONS_PER_STEP
    opt.set_params(prm)

    # ---- e2e: host buffers -> pack + H2D -> optimize -> D2H, through the C ABI
    h2d = pb.pose_idx.nbytes * 6 + pb.uv.nbytes * 2 + pb.enc.nbytes + pb.lm_prior.nbytes + 96 + pb.q0.nbytes + pb.l0.nbytes + 96
    d2h = pb.q0.nbytes + pb.l0.nbytes + 96

    def e2e_step():
        opt.set_problem(pb)
        opt.set_values(pb.q0, pb.l0, pb.x0)
        q, l, x = opt.optimize(asc.ASC_BATCH_LM)
        return opt.iterations()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_iters = 0
    e_steps = max(1, min(args.steps, 2))
    for _ in range(e_steps):
        e_iters += e2e_step()
    torch.cuda.synchronize()
    e_dt = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": e_iters / e_dt, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e_steps,
           "ms_per_step": 1e3 * e_dt / e_steps, "timer": "host perf_counter between device synchronisations (host packing is part of the path)"}

    # ---- per-kernel achieved bandwidth (CUDA events on the launching stream, live calls of the normal entry points)
    peak, peak_src = hbm_peak()
    kernels = {}
    opt.set_values(pb.q0, pb.l0, pb.x0)
    n_cl, nz, sc = opt.solver_info()
    for kid in [k for k in (1, 7, 2, 6, 3, 4, 5, 8, 9) if k != DOMINANT]:
        opt.profile_begin(kid, 64)
        if kid in (1, 2, 7):
            for _ in range(8):
                opt.linearize()
        elif kid == 6:
            for _ in range(8):
                opt.graph_error()
        else:
            opt.linearize()            # fresh linearisation: the solve below starts cold (no warm start, coarse operator rebuilt)
            opt.profile_begin(kid, 64)
            opt.solve_damped(1e-5)
        n_s, mean_k, min_k = opt.profile_read()
        if n_s:
            b = algorithmic_bytes(kid, N, P, L, M, nz, sc)
            kernels[KERNEL_NAMES[kid]] = {"samples": n_s, "mean_us": 1e3 * mean_k, "algorithmic_MB": b / 1e6, "achieved_GBs": b / (mean_k * 1e-3) / 1e9,
                                          "frac": b / (mean_k * 1e-3) / 1e9 / peak}
    bytes_a = algorithmic_bytes(DOMINANT, N, P, L, M)
    achieved = bytes_a / (mean_ms * 1e-3) / 1e9 if n_prof else None
    kernels[KERNEL_NAMES[DOMINANT]] = {"samples": n_prof, "mean_us": 1e3 * mean_ms, "algorithmic_MB": bytes_a / 1e6, "achieved_GBs": achieved, "frac": achieved / peak if achieved else None,
                                      "where": "inside the timed region"}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            traffic = json.load(open(tpath)).get("schur_pass_c_kernel")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "schur_pass_c_kernel<%d> (implicit Schur product, third pass; largest share of the step among this library's kernels)" % N, "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_a, "samples": n_prof, "frac_of_nominal_8000": achieved / 8000.0 if achieved else None}
    lin = kernels.get(KERNEL_NAMES[1])
    line = {"metric": "LM iterations/s", "value": value, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "poses": pb_full.n_poses, "landmarks": pb_full.n_landmarks, "observations": pb_full.n_obs, "dof": N,
                       "mode": "BatchLM", "l2": "inputs larger than L2 (Jacobian planes %.0f MB per rank)" % (16 * (N + 3) * M / 1e6),
                       "step": "first %d LM iterations of one BatchLM optimize() from the initial estimate" % LM_ITERATIONS_PER_STEP,
                       "lm_iterations_per_step": iters // max(args.steps, 1), "pcg_iterations_per_lm_iteration": pcg_per_iter,
                       "error_after_step": final_error},
            "factor_linearizations_per_s": (M * world / (lin["mean_us"] * 1e-6)) if lin else None,
            "full_optimize": full, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "kernels": kernels}
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        from oracle import Oracle
        p1 = asc.default_params()
        p1.max_iterations = LM_ITERATIONS_PER_STEP
        orc = Oracle(pb_full, p1)
        t0 = time.perf_counter()
        r = orc.optimize(pb_full.q0, pb_full.l0, pb_full.x0, mode=0, solver=1)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": r["iterations"] / dt, "unit": "iterations/s", "cores": Oracle.max_threads(), "kind": "port",
                                "sample": "the same step as the GPU arm: first %d LM iterations of the BatchLM optimize (%d factors, per-factor FK, cluster-Jacobi "
                                          "PCG to 1e-12: %s iterations); the reference's GTSAM/DART stack is not buildable here, so this is the "
                                          "oracle port (OpenMP)" % (LM_ITERATIONS_PER_STEP, pb_full.n_obs, [int(gg.pcg_iterations) for gg in r["log"]]),
                                "seconds": dt}
    if rank == 0:
        print(json.dumps(line), flush=True)
    opt.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
