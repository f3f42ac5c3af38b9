#!/usr/bin/env python
"""bench.py -- LM iterations/s of the batch optimisation path on the synthetic Mico problem (BASELINE.json).

One "step" = the first three LM iterations of the workload's BatchLM optimize() from its initial estimate (ArmSlamCalib::OptimizeStep,
BatchLM branch, src/ArmSlamCalib.cpp:458-467): the bounded sample that BOTH arms run at EVERY GPU count, so that the driver's ratios
(GPU / CPU, N GPUs / 1 GPU) compare the same work.  (A full optimize of the 8-GPU weak-scaling workload takes ~50 s per step on the
GPUs and a quarter of an hour per step on the CPU, against the driver's 870 s per run.)  `value` = LM iterations / second with the
graph and the initial values already resident in HBM; `e2e` = the same step through the C ABI from host buffers (graph packing + H2D +
optimize + D2H every step).  The FULL optimize to GTSAM's convergence test is run once after the timed region and reported as
`full_optimize` (LM iterations, iterations/s, PCG iterations, per-stage device times).

Workloads (--workload; BASELINE.json configs): c2 (default at --gpus 1), c3 (the same arrays, BatchDogleg), c4 / c4_12 (random
7 / 12-DOF chains, 50k poses), c5 (200k poses / 40M observations), c2xN.  With --gpus N > 1 the default is WEAK scaling: the C2
generator at N times the poses / landmarks / observations (c2xN), landmarks sharded over the ranks, and `value` counts one LM
iteration of the N-fold problem as N iterations of a C2-sized problem (the units all ranks processed per second).

`--impl reference` times the CPU oracle port (the reference's GTSAM/DART stack cannot be built here) on the host
cores: the same landmark elimination, the same two-level preconditioner and PCG as the GPU path, OpenMP over all cores, on a
bounded sample (the first three LM iterations) of the same workload.  It loads only libasc_host.so (generator) and the oracle.
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
                5: "schur_pass_c (K3)", 6: "error_proj (K1e)", 7: "proj_kk (K1b)", 8: "ml_vcycle (coarse level of the preconditioner)",
                9: "pcg_cluster_update (preconditioner)"}
SAMPLE_ITERATIONS = 3   # the bounded sample both arms can run: first 3 LM iterations of the BatchLM optimize


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(kid, N, P, L, M, sc=0, ml_bytes=0):
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
    if kid == 8:   # every level's blocks once per Chebyshev product (2 m per level) -- L2-resident, latency-bound: reported, not a roofline
        return ml_bytes
    if kid == 9:   # cluster blocks in; x, p, r, y in; x, r, z out; restriction out
        return sc * 8 + P * N * 8 * 7
    if kid in (6, 7):   # uv + 2 indices in; camera rows + landmark table in
        return M * (16 + 8) + P * 96 + L * 32
    return 0


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


def make_workload(name, n_gpus):
    """-> (problem, params, mode, description, c2_equivalents).  Needs only the host library (generator)."""
    import arm_slam_calib_b200 as asc
    from arm_slam_calib_b200 import configs
    if name is None:
        name = "c2" if n_gpus == 1 else "c2x%d" % n_gpus
    prm = asc.default_params()
    mode, eq = asc.ASC_BATCH_LM, 1.0
    if name == "c2":
        pb, desc = configs.c2(), "C2: synthetic Mico 6-DOF, 10k poses / ~90k landmarks / ~2M projection observations, BatchLM"
    elif name == "c3":
        pb, desc, mode = configs.c2(), "C3: the C2 arrays (synthetic Mico, 10k poses / ~2M observations), BatchDogleg", asc.ASC_BATCH_DOGLEG
    elif name in ("c4", "c4_7", "c4_12"):
        n_dof = 12 if name.endswith("12") else 7
        pb = configs.c4(n_dof)
        prm = configs.params_for("c4", asc.default_params)
        desc = "C4: sim_genrobot random %d-DOF serial chain, 50k poses / ~500k landmarks / ~10M observations, BatchLM" % n_dof
    elif name == "c5":
        pb, desc, eq = configs.c5(), "C5: synthetic Mico 6-DOF, 200k poses / ~2M landmarks / ~40M observations, BatchLM", 20.0
    elif name.startswith("c2x"):
        k = float(name[3:])
        pb, eq = configs.c2(scale=k), k
        desc = "C2 x %g (weak scaling): synthetic Mico 6-DOF, %dk poses, BatchLM" % (k, round(10 * k))
    else:
        raise SystemExit("unknown workload %r" % name)
    if n_gpus > 1:
        desc += ", landmarks sharded over %d GPUs" % n_gpus
    return pb, prm, mode, desc, eq, name


def config_dict(pb, desc, name, mode, eq):
    """identical in both arms (the driver compares the dicts)"""
    return {"workload": desc, "name": name, "poses": int(pb.n_poses), "landmarks": int(pb.n_landmarks), "observations": int(pb.n_obs), "dof": int(pb.n_dof),
            "mode": "BatchLM" if mode == 0 else "BatchDogleg", "c2_equivalents": eq,
            "l2": "inputs larger than L2 (Jacobian planes %.0f MB)" % (16 * (pb.n_dof + 3) * pb.n_obs / 1e6),
            "step": "the first %d LM iterations of one optimize() from the initial estimate (both arms; the reference arm drops to the first iteration "
                    "on workloads larger than C2 to stay inside the driver's time limit)" % SAMPLE_ITERATIONS}


def cpu_sample(pb, prm, mode, threads=None, iterations=SAMPLE_ITERATIONS):
    """The oracle port on the host cores: first `iterations` iterations of the optimize.  -> (iterations/s, seconds, log, threads)"""
    from oracle import Oracle
    import copy
    p1 = copy.copy(prm)
    p1.max_iterations = iterations
    if threads:
        Oracle.set_threads(threads)
    orc = Oracle(pb, p1)
    t0 = time.perf_counter()
    r = orc.optimize(pb.q0, pb.l0, pb.x0, mode=mode, solver=1)
    dt = time.perf_counter() - t0
    used = threads or Oracle.max_threads()
    if threads:
        Oracle.set_threads(Oracle.max_threads())
    return r["iterations"] / dt, dt, r["log"], used


def cpu_linearize_rate(pb, prm):
    """factor linearisations/s of the oracle's linearise (per-factor FK like the reference, all cores; includes copying the blocks out)"""
    from oracle import Oracle
    orc = Oracle(pb, prm)
    orc.error(pb.q0, pb.l0, pb.x0)          # touch the problem once
    t0 = time.perf_counter()
    orc.linearize(pb.q0, pb.l0, pb.x0)
    return pb.n_obs / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from arm_slam_calib_b200.build import build_host
    build_host()
    from oracle.capi import build_oracle
    build_oracle()
    from oracle import Oracle
    pb, prm, mode, desc, eq, name = make_workload(args.workload, args.gpus)
    ncpu = len(os.sched_getaffinity(0))
    if "WORLD_SIZE" in os.environ and Oracle.max_threads() < ncpu:
        # torchrun exports OMP_NUM_THREADS=1 to its workers; rank 0 is the only one working here, so it takes the host's cores back
        Oracle.set_threads(ncpu)
    cores = Oracle.max_threads()
    sample = ("first %d iterations of the workload's optimize from the initial estimate (per iteration: linearisation of all factors with per-factor FK, "
              "landmark elimination, PCG to 1e-12 with the same cluster-Jacobi + multilevel coarse preconditioner as the GPU path, error evaluation); "
              "oracle port, OpenMP on all host cores" % SAMPLE_ITERATIONS)
    n_it = SAMPLE_ITERATIONS if eq <= 1.5 else 1   # larger workloads: the first LM iteration only (a 3-iteration step of c2x8 is ~2 CPU-minutes)
    sample = sample.replace("first %d iterations" % SAMPLE_ITERATIONS, "first %d iteration(s)" % n_it)
    warm = min(args.warmup, 1) if eq <= 1.5 else 0   # a CPU step is tens of seconds: one untimed pass is all the warm-up it needs
    for _ in range(warm):
        cpu_sample(pb, prm, mode, iterations=n_it)
    budget_s = 420.0                               # the driver allows 870 s per run: stop adding timed steps when the next one would not fit
    t0 = time.perf_counter()
    its, log, steps_run, last = 0, None, 0, 0.0
    for _ in range(args.steps):
        if steps_run >= 1 and (time.perf_counter() - t0) + last > budget_s:
            break
        rate, last, log, _ = cpu_sample(pb, prm, mode, iterations=n_it)
        its += len(log)
        steps_run += 1
    dt = time.perf_counter() - t0
    val = eq * its / dt
    line = {"impl": "reference", "metric": "LM iterations/s", "value": val, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "warmup_run": warm, "steps_run": steps_run, "ms_per_step": 1e3 * dt / steps_run, "higher_is_better": True,
            "scaling": "weak" if args.gpus > 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(pb, desc, name, mode, eq),
            "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample,
                             "pcg_iterations": [int(g.pcg_iterations) for g in log]},
            "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="c2 (default at 1 GPU), c3, c4, c4_12, c5, c2x<k>; default at N GPUs: c2xN (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-run", action="store_true",
                    help="warm-up + timed region only (no e2e, per-kernel or CPU legs): the command the ncu launch list is taken from")
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
    pb_full, prm, mode, desc, eq, wname = make_workload(args.workload, args.gpus)
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

    import copy
    p3 = copy.copy(prm)
    p3.max_iterations = SAMPLE_ITERATIONS          # the step: first 3 LM iterations
    opt = asc.BatchOptimizer(pb, p3, comm=comm)
    opt.save_values()

    def step():
        opt.restore_values()
        opt.optimize(mode)
        return opt.iterations()

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        step()
    clocks.wait_first_row()
    barrier()
    clocks.mark()
    launches0 = opt.launch_count()
    opt.timer_start()
    iters = 0
    for _ in range(args.steps):
        iters += step()
    ms = opt.timer_stop()
    barrier()
    clk = clocks.stop()
    launches = opt.launch_count() - launches0
    ms = max_over_ranks(ms)
    value = eq * iters / (ms * 1e-3)
    log = list(opt.log)
    sample = {"iterations_per_s": value, "ms_per_step": ms / args.steps, "lm_iterations_per_step": SAMPLE_ITERATIONS,
              "pcg_iterations": [int(g.pcg_iterations) for g in log], "errors": [g.error for g in log], "inner_steps": [int(g.inner_steps) for g in log]}
    if args.profile_run:
        if rank == 0:
            print(json.dumps({"metric": "LM iterations/s", "value": value, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms / args.steps, "profile_run": True, "gpu_launches": int(launches)}), flush=True)
        opt.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- the FULL optimize to GTSAM's convergence test, once, device-resident (context for the sample above)
    opt.set_params(prm)
    opt.restore_values()
    barrier()
    opt.timer_start()
    opt.optimize(mode)
    f_ms = max_over_ranks(opt.timer_stop())
    flog = list(opt.log)
    full = {"lm_iterations": len(flog), "seconds": f_ms * 1e-3, "iterations_per_s": eq * len(flog) / (f_ms * 1e-3), "final_error": opt.error(),
            "pcg_iterations": [int(g.pcg_iterations) for g in flog], "inner_steps": [int(g.inner_steps) for g in flog],
            "stage_ms": {"linearize": sum(g.linearize_ms for g in flog), "prepare": sum(g.prepare_ms for g in flog), "coarse_refresh": sum(g.coarse_ms for g in flog),
                         "pcg": sum(g.pcg_ms for g in flog), "solve_total": sum(g.solve_ms for g in flog)}}
    opt.set_params(p3)

    # ---- e2e: host buffers -> pack + H2D -> the same step -> D2H, through the C ABI
    h2d = pb.pose_idx.nbytes * 6 + pb.uv.nbytes * 2 + pb.enc.nbytes + pb.lm_prior.nbytes + 96 + pb.q0.nbytes + pb.l0.nbytes + 96
    d2h = pb.q0.nbytes + pb.l0.nbytes + 96

    def e2e_step():
        opt.set_problem(pb)
        opt.set_values(pb.q0, pb.l0, pb.x0)
        q, l, x = opt.optimize(mode)
        return opt.iterations()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_iters = 0
    e_steps = max(1, min(args.steps, 3))
    for _ in range(e_steps):
        e_iters += e2e_step()
    torch.cuda.synchronize()
    e_dt = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": eq * e_iters / e_dt, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e_steps,
           "ms_per_step": 1e3 * e_dt / e_steps, "timer": "host perf_counter between device synchronisations (host packing is part of the path)"}

    # ---- per-kernel achieved bandwidth (CUDA events on the launching stream, live calls of the normal entry points)
    peak, peak_src = hbm_peak()
    kernels = {}
    opt.set_values(pb.q0, pb.l0, pb.x0)
    n_cl, nz, sc = opt.solver_info()
    cinfo = opt.coarse_info()
    ml_bytes = sum(2 * cinfo["degree"] * blocks * N * N * 8 for _, blocks in cinfo["levels"][:-1])
    for kid in (1, 7, 2, 6, 3, 4, 5, 9):
        opt.profile_begin(kid, 128)
        if kid in (1, 2, 7):
            for _ in range(8):
                opt.linearize()
        elif kid == 6:
            for _ in range(8):
                opt.graph_error()
        else:
            opt.linearize()            # fresh linearisation: the solve below starts cold (no warm start)
            opt.profile_begin(kid, 128)
            opt.solve_damped(1e-5)
        n_s, mean_k, min_k = opt.profile_read()
        if n_s:
            b = algorithmic_bytes(kid, N, P, L, M, sc, ml_bytes)
            kernels[KERNEL_NAMES[kid]] = {"samples": n_s, "mean_us": 1e3 * mean_k, "algorithmic_MB": b / 1e6, "achieved_GBs": b / (mean_k * 1e-3) / 1e9,
                                          "frac": b / (mean_k * 1e-3) / 1e9 / peak}
    # the whole linearisation (FK + K1 + K1b + K2a + assembly) for the factor-linearisations/s metric
    lin_ms = []
    for _ in range(8):
        st = opt.linearize()
        lin_ms.append(st.linearize_ms + st.assemble_ms)
    lin_total_ms = sorted(lin_ms)[len(lin_ms) // 2]
    DOMINANT = 5                # schur_pass_c: largest share of the step among this library's HBM-bound kernels
    dom = kernels.get(KERNEL_NAMES[DOMINANT], {})
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            traffic = json.load(open(tpath)).get("schur_pass_c_kernel")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "schur_pass_c_kernel<%d> (implicit Schur product, third pass; largest share of the step among this library's HBM-bound kernels)" % N,
                "achieved": dom.get("achieved_GBs"), "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": dom.get("frac"), "traffic": traffic,
                "algorithmic_bytes_per_launch": algorithmic_bytes(DOMINANT, N, P, L, M), "samples": dom.get("samples"),
                "frac_of_nominal_8000": dom.get("achieved_GBs") / 8000.0 if dom.get("achieved_GBs") else None,
                "timing": "CUDA events on the launching stream around live launches inside a damped solve of the workload (un-graphed for the measurement)"}
    line = {"metric": "LM iterations/s", "value": value, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if args.gpus > 1 else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(pb_full, desc, wname, mode, eq),
            "factor_linearizations_per_s": M * world / (lin_total_ms * 1e-3),
            "factor_linearizations_note": "all observations / device time of the WHOLE linearisation (FK + K1 + K1b + K2a + assembly), %.3f ms" % lin_total_ms,
            "full_optimize": full, "sample": sample, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "kernels": kernels,
            "coarse_hierarchy": cinfo, "clusters": n_cl}
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        from oracle import Oracle
        cores = Oracle.max_threads()
        rate_all, dt_all, clog, _ = cpu_sample(pb_full, prm, mode)
        rate_1, dt_1, clog1, _ = cpu_sample(pb_full, prm, mode, threads=1, iterations=1)
        gpu_err, cpu_err = sample["errors"], [g.error for g in clog]
        line["cpu_baseline"] = {"value": eq * rate_all, "unit": "iterations/s", "cores": cores, "kind": "port",
                                "sample": "first %d iterations of the same optimize (%d factors, per-factor FK, landmark elimination, PCG to 1e-12 with the same "
                                          "cluster-Jacobi + multilevel preconditioner as the GPU path: %s iterations); the reference's GTSAM/DART stack is not "
                                          "buildable here, so this is the oracle port (OpenMP, all cores)" % (SAMPLE_ITERATIONS, pb_full.n_obs, [int(g.pcg_iterations) for g in clog]),
                                "seconds": dt_all,
                                "one_thread": {"value": eq * rate_1, "unit": "iterations/s", "cores": 1, "seconds": dt_1,
                                               "sample": "first LM iteration only (the reference serialises factor evaluation behind a mutex, RobotProjectionFactor.h:203)"},
                                "factor_linearizations_per_s": cpu_linearize_rate(pb_full, prm),
                                "per_iteration_error_gpu": gpu_err, "per_iteration_error_cpu": cpu_err,
                                "max_rel_error_diff": max(abs(a - b) / b for a, b in zip(gpu_err, cpu_err)) if len(gpu_err) == len(cpu_err) else None}
    if rank == 0:
        print(json.dumps(line), flush=True)
    opt.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
