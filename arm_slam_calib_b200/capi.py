"""ctypes binding of include/asc.h plus a small host-side mirror of the reference's optimiser surface.

`BatchOptimizer` plays the role of gtsam::LevenbergMarquardtOptimizer / gtsam::DoglegOptimizer as
ArmSlamCalib::OptimizeStep uses them (src/ArmSlamCalib.cpp:438-467): construct over (graph, values),
call optimize(), read error().  The graph is the flattened `Problem`.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _ctypes_defs as D

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libasc_b200.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_vp = C.c_void_p

# name -> (restype, argtypes); also the list tests/test_abi.py checks against include/asc.h
SIGNATURES = {
    "asc_default_params": (None, [C.POINTER(D.AscParams)]),
    "asc_last_error": (C.c_char_p, []),
    "asc_version": (C.c_char_p, []),
    "asc_create": (C.c_int, [C.POINTER(D.AscParams), C.POINTER(_vp)]),
    "asc_destroy": (C.c_int, [_vp]),
    "asc_set_params": (C.c_int, [_vp, C.POINTER(D.AscParams)]),
    "asc_set_chain": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    "asc_set_camera": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "asc_set_problem": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "asc_set_values": (C.c_int, [_vp, _vp, _vp, _vp]),
    "asc_get_values": (C.c_int, [_vp, _vp, _vp, _vp]),
    "asc_append": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "asc_error": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "asc_linearize": (C.c_int, [_vp, C.POINTER(D.AscLinearStats)]),
    "asc_optimize": (C.c_int, [_vp, C.c_int32, C.POINTER(D.AscIterLog), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "asc_indeterminate_key": (C.c_int, [_vp, C.POINTER(C.c_char), C.POINTER(C.c_int64)]),
    "asc_get_projection_blocks": (C.c_int, [_vp, _vp, _vp, _vp]),
    "asc_get_landmark_blocks": (C.c_int, [_vp, _vp, _vp, _vp]),
    "asc_get_pose_blocks": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "asc_set_drift": (C.c_int, [_vp, _vp, _vp, _vp]),
    "asc_set_velocity": (C.c_int, [_vp, _vp, _vp, _vp]),
    "asc_solve_damped": (C.c_int, [_vp, C.c_double, _vp, _vp, _vp, C.POINTER(C.c_int32)]),
    "asc_extrinsic_marginal": (C.c_int, [_vp, _vp, C.POINTER(C.c_int32)]),
    "asc_save_values": (C.c_int, [_vp]),
    "asc_restore_values": (C.c_int, [_vp]),
    "asc_launch_count": (C.c_int64, [_vp]),
    "asc_solver_info": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "asc_coarse_info": (C.c_int, [_vp, C.POINTER(C.c_int32), _vp, _vp, C.c_int32, C.POINTER(C.c_int32)]),
    "asc_debug_ml_trace": (C.c_int, [_vp, _vp, C.c_int32]),
    "asc_timer_start": (C.c_int, [_vp]),
    "asc_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "asc_profile_begin": (C.c_int, [_vp, C.c_int32, C.c_int32]),
    "asc_profile_read": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "asc_comm_id_size": (C.c_int, []),
    "asc_comm_create_id": (C.c_int, [_vp]),
    "asc_comm_init": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp]),
    "asc_shard_by_landmark": (C.c_int, [C.c_int32, C.c_int64, _vp, C.c_int32, _vp]),
    "asc_sim_default_config": (None, [C.POINTER(D.AscSimConfig)]),
    "asc_chain_mico": (C.c_int, [_vp, _vp, _vp]),
    "asc_chain_generate": (C.c_int, [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, _vp, _vp, _vp]),
    "asc_sim_create": (C.c_int, [C.POINTER(D.AscSimConfig), C.c_int32, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double,
                                 C.c_double, C.POINTER(_vp)]),
    "asc_sim_destroy": (None, [_vp]),
    "asc_sim_sizes": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "asc_sim_export": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "asc_chain_fk": (C.c_int, [C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    "asc_srand": (None, [C.c_uint]),
}


def lib():
    """Loads libasc_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("libasc_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


_HOST_LIB_PATH = os.path.join(_HERE, "libasc_host.so")
_hostlib = None
HOST_SYMBOLS = ("asc_default_params", "asc_last_error", "asc_version", "asc_sim_default_config", "asc_chain_mico", "asc_chain_generate",
                "asc_sim_create", "asc_sim_destroy", "asc_sim_sizes", "asc_sim_export", "asc_chain_fk", "asc_srand")


def hostlib():
    """libasc_host.so: the generator and the default parameters without any CUDA code (never maps the product library)."""
    global _hostlib
    if _hostlib is None:
        if not os.path.exists(_HOST_LIB_PATH):
            raise RuntimeError("libasc_host.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(_HOST_LIB_PATH)
        for name in HOST_SYMBOLS:
            res, args = SIGNATURES[name]
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _hostlib = L
    return _hostlib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


class AscError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("asc status %d: %s" % (code, msg))
        self.code = code


class IndeterminantLinearSystemException(AscError):
    """Same name and role as gtsam::IndeterminantLinearSystemException (ArmSlamCalib.cpp:451-455)."""

    def __init__(self, code, msg, kind, index):
        super().__init__(code, msg)
        self.kind, self.index = kind, index

    def nearbyVariable(self):
        return (self.kind, self.index)


def default_params():
    p = D.AscParams()
    hostlib().asc_default_params(C.byref(p))
    return p


def optimization_mode_from_string(name, previous=D.ASC_BATCH_DOGLEG):
    """ArmSlamCalib::Params::InitializeNodehandleParams, src/ArmSlamCalib.cpp:408-422: the code reads
    "BatchDogLeg" while the README says BatchDogleg; unknown strings silently keep the previous mode.
    ISAM is out of scope for this library and is reported as such."""
    if name == "BatchLM":
        return D.ASC_BATCH_LM
    if name in ("BatchDogLeg", "BatchDogleg"):
        return D.ASC_BATCH_DOGLEG
    if name == "ISAM":
        raise NotImplementedError("optimization_mode=ISAM is outside the batch path this library accelerates")
    return previous


@dataclass
class Problem:
    """The factor graph of ArmSlamCalib (graph + chain + calibration), flattened (include/asc.h)."""
    n_dof: int
    t_p2j: np.ndarray
    axis: np.ndarray
    t_c2j: np.ndarray
    fx: float
    fy: float
    cx: float
    cy: float
    pose_idx: np.ndarray
    lm_idx: np.ndarray
    uv: np.ndarray
    enc: np.ndarray
    lm_prior: np.ndarray
    x_prior: np.ndarray
    continuous: np.ndarray = None
    has_enc: np.ndarray = None
    t_base: np.ndarray = None
    # DriftFactor(q_{p-1}, q_p) links (DriftFactor.h:44-61; use_drift_noise in the drivers): has_drift[p] != 0, measurement
    # drift_meas[p] = encoders[p] - encoders[p-1] (ArmSlamCalib.cpp:1668), per-joint sigma (drift_noise_level, default 0.05)
    has_drift: np.ndarray = None
    drift_meas: np.ndarray = None
    sigma_drift: np.ndarray = None
    # VelocityFactor(q_{p-1}, q_p, velocities[p-1]) links (VelocityFactor.h:40-62; use_velocity_noise): residual (q_p - q_{p-1}) - vel_meas[p],
    # per-joint sigma (velocity_noise_level, default 0.01)
    has_vel: np.ndarray = None
    vel_meas: np.ndarray = None
    sigma_vel: np.ndarray = None
    # initial estimate (Values)
    q0: np.ndarray = None
    l0: np.ndarray = None
    x0: np.ndarray = None
    meta: dict = field(default_factory=dict)

    @property
    def n_poses(self):
        return self.enc.shape[0]

    @property
    def n_landmarks(self):
        return self.lm_prior.shape[0]

    @property
    def n_obs(self):
        return self.pose_idx.shape[0]


def shard_by_landmark(n_landmarks, lm_idx, world_size):
    out = np.zeros(world_size + 1, dtype=np.int32)
    lm_idx = np.ascontiguousarray(lm_idx, dtype=np.int32)
    rc = lib().asc_shard_by_landmark(n_landmarks, lm_idx.shape[0], _ptr(lm_idx), world_size, _ptr(out))
    if rc:
        raise AscError(rc, lib().asc_last_error().decode())
    return out


def shard_problem(pb, rank, world_size):
    """Rank's shard: its landmark range and all their observations; poses/encoders replicated."""
    bounds = shard_by_landmark(pb.n_landmarks, pb.lm_idx, world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    sel = (pb.lm_idx >= lo) & (pb.lm_idx < hi)
    import dataclasses
    return dataclasses.replace(
        pb, pose_idx=np.ascontiguousarray(pb.pose_idx[sel]), lm_idx=np.ascontiguousarray(pb.lm_idx[sel] - lo),
        uv=np.ascontiguousarray(pb.uv[sel]), lm_prior=np.ascontiguousarray(pb.lm_prior[lo:hi]),
        l0=None if pb.l0 is None else np.ascontiguousarray(pb.l0[lo:hi]), meta=dict(pb.meta, lm_range=(lo, hi), obs_sel=sel))


class BatchOptimizer:
    """Device-resident optimiser over one Problem.  Mirrors the GTSAM optimiser objects the reference
    constructs in OptimizeStep: BatchOptimizer(problem, params).optimize(mode) / .error()."""

    def __init__(self, problem, params=None, comm=None):
        self.L = lib()
        self.pb = problem
        self.params = params if params is not None else default_params()
        self.h = _vp()
        self._check(self.L.asc_create(C.byref(self.params), C.byref(self.h)))
        pb = problem
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self._keep = dict(t_p2j=f64(pb.t_p2j), axis=f64(pb.axis), t_c2j=f64(pb.t_c2j), t_base=f64(pb.t_base),
                          cont=None if pb.continuous is None else np.ascontiguousarray(pb.continuous, dtype=np.uint8))
        k = self._keep
        self._check(self.L.asc_set_chain(self.h, pb.n_dof, _ptr(k["t_p2j"]), _ptr(k["axis"]), _ptr(k["t_c2j"]), _ptr(k["cont"]), _ptr(k["t_base"])))
        self._check(self.L.asc_set_camera(self.h, pb.fx, pb.fy, pb.cx, pb.cy))
        if comm is not None:
            rank, world, id_bytes = comm
            buf = np.frombuffer(bytes(id_bytes), dtype=np.uint8).copy()
            self._check(self.L.asc_comm_init(self.h, rank, world, _ptr(buf)))
        self.set_problem(pb)
        if pb.q0 is not None:
            self.set_values(pb.q0, pb.l0, pb.x0)
        self.log = []
        self._error = None

    def _check(self, rc):
        if rc == D.ASC_OK:
            return
        msg = self.L.asc_last_error().decode()
        if rc == D.ASC_ERR_INDETERMINATE:
            kind, idx = C.c_char(), C.c_int64()
            self.L.asc_indeterminate_key(self.h, C.byref(kind), C.byref(idx))
            raise IndeterminantLinearSystemException(rc, msg, kind.value.decode() if kind.value else "?", idx.value)
        raise AscError(rc, msg)

    def set_params(self, params):
        self.params = params
        self._check(self.L.asc_set_params(self.h, C.byref(params)))

    def set_problem(self, pb):
        pi = np.ascontiguousarray(pb.pose_idx, dtype=np.int32)
        li = np.ascontiguousarray(pb.lm_idx, dtype=np.int32)
        uv = np.ascontiguousarray(pb.uv, dtype=np.float64)
        enc = np.ascontiguousarray(pb.enc, dtype=np.float64)
        he = None if pb.has_enc is None else np.ascontiguousarray(pb.has_enc, dtype=np.uint8)
        lp = np.ascontiguousarray(pb.lm_prior, dtype=np.float64)
        xp = np.ascontiguousarray(pb.x_prior, dtype=np.float64)
        self._check(self.L.asc_set_problem(self.h, enc.shape[0], lp.shape[0], pi.shape[0], _ptr(pi), _ptr(li), _ptr(uv), _ptr(enc), _ptr(he), _ptr(lp), _ptr(xp)))
        self.P, self.Lm, self.M, self.N = enc.shape[0], lp.shape[0], pi.shape[0], pb.n_dof
        if getattr(pb, "has_drift", None) is not None:
            self.set_drift(pb.has_drift, pb.drift_meas, pb.sigma_drift)
        if getattr(pb, "has_vel", None) is not None:
            self.set_velocity(pb.has_vel, pb.vel_meas, pb.sigma_vel)

    def _set_links(self, fn, has, meas, sigma):
        if has is None:
            self._check(fn(self.h, None, None, None))
            return
        hd = np.ascontiguousarray(has, dtype=np.uint8)
        dm = np.ascontiguousarray(meas, dtype=np.float64)
        sg = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float64), (self.N,)))
        self._check(fn(self.h, _ptr(hd), _ptr(dm), _ptr(sg)))

    def set_drift(self, has_drift, drift_meas, sigma_drift):
        """DriftFactor(q_{p-1}, q_p) links (asc_set_drift); has_drift None removes them."""
        self._set_links(self.L.asc_set_drift, has_drift, drift_meas, sigma_drift)

    def set_velocity(self, has_vel, vel_meas, sigma_vel):
        """VelocityFactor(q_{p-1}, q_p, qDot) links (asc_set_velocity); has_vel None removes them."""
        self._set_links(self.L.asc_set_velocity, has_vel, vel_meas, sigma_vel)

    def set_values(self, q, l, x):
        q = np.ascontiguousarray(q, dtype=np.float64)
        l = np.ascontiguousarray(l, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._check(self.L.asc_set_values(self.h, _ptr(q), _ptr(l), _ptr(x)))

    def append(self, pose_idx, lm_idx, uv, enc_new, lm_prior_new, q_new, l_new, has_enc_new=None):
        """asc_append: grow the packed graph by new keyframes / landmarks / observations, keeping the current estimate (incremental mode,
        sim_genrobot.cpp:97-104).  New observations index the grown pose / landmark sets."""
        pi = np.ascontiguousarray(pose_idx, dtype=np.int32); li = np.ascontiguousarray(lm_idx, dtype=np.int32)
        uvs = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        enc_new = np.ascontiguousarray(enc_new, dtype=np.float64).reshape(-1, self.N)
        q_new = np.ascontiguousarray(q_new, dtype=np.float64).reshape(-1, self.N)
        lp = np.ascontiguousarray(lm_prior_new, dtype=np.float64).reshape(-1, 3)
        ln = np.ascontiguousarray(l_new, dtype=np.float64).reshape(-1, 3)
        he = None if has_enc_new is None else np.ascontiguousarray(has_enc_new, dtype=np.uint8)
        self._check(self.L.asc_append(self.h, enc_new.shape[0], lp.shape[0], pi.shape[0], _ptr(pi), _ptr(li), _ptr(uvs), _ptr(enc_new), _ptr(he),
                                      _ptr(lp), _ptr(q_new), _ptr(ln)))
        self.P += enc_new.shape[0]; self.Lm += lp.shape[0]; self.M += pi.shape[0]
        self._error = None

    def values(self):
        q = np.empty((self.P, self.N)); l = np.empty((self.Lm, 3)); x = np.empty(12)
        self._check(self.L.asc_get_values(self.h, _ptr(q), _ptr(l), _ptr(x)))
        return q, l, x

    def graph_error(self):
        e = C.c_double()
        self._check(self.L.asc_error(self.h, C.byref(e)))
        return e.value

    def linearize(self):
        st = D.AscLinearStats()
        self._check(self.L.asc_linearize(self.h, C.byref(st)))
        return st

    def projection_blocks(self):
        b = np.empty((self.M, 2)); jq = np.empty((self.M, 2, self.N)); jl = np.empty((self.M, 2, 3))
        self._check(self.L.asc_get_projection_blocks(self.h, _ptr(b), _ptr(jq), _ptr(jl)))
        return b, jq, jl

    def landmark_blocks(self):
        c = np.empty((self.Lm, 3, 3)); gl = np.empty((self.Lm, 3)); wk = np.empty((self.Lm, 6, 3))
        self._check(self.L.asc_get_landmark_blocks(self.h, _ptr(c), _ptr(gl), _ptr(wk)))
        return c, gl, wk

    def pose_blocks(self):
        bpp = np.empty((self.P, self.N, self.N)); bpk = np.empty((self.P, self.N, 6)); gp = np.empty((self.P, self.N))
        bkk = np.empty((6, 6)); gk = np.empty(6)
        self._check(self.L.asc_get_pose_blocks(self.h, _ptr(bpp), _ptr(bpk), _ptr(gp), _ptr(bkk), _ptr(gk)))
        return bpp, bpk, gp, bkk, gk

    def solve_damped(self, lam):
        dq = np.empty((self.P, self.N)); dl = np.empty((self.Lm, 3)); dx = np.empty(6)
        its = C.c_int32()
        self._check(self.L.asc_solve_damped(self.h, float(lam), _ptr(dq), _ptr(dl), _ptr(dx), C.byref(its)))
        return dq, dl, dx, its.value

    def solver_info(self):
        """(clusters, coarse dimension, doubles in the cluster blocks) of the packed problem."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int64()
        self._check(self.L.asc_solver_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def coarse_info(self):
        """Hierarchy of the multilevel coarse solve: dict(levels=[(nodes, nonzero N x N blocks), ...], degree=Chebyshev degree); the last
        level is the dense top level.  Empty list: no coarse level."""
        n, deg = C.c_int32(), C.c_int32()
        nodes = np.zeros(16, np.int32); blocks = np.zeros(16, np.int64)
        self._check(self.L.asc_coarse_info(self.h, C.byref(n), _ptr(nodes), _ptr(blocks), 16, C.byref(deg)))
        return dict(levels=[(int(nodes[i]), int(blocks[i])) for i in range(min(n.value, 16))], degree=deg.value)

    def extrinsic_marginal(self):
        """K0 block of (A^T A)^-1 at the current values (src/ArmSlamCalib.cpp:490-493); returns (cov 6x6, PCG iterations)."""
        cov = np.empty((6, 6))
        its = C.c_int32()
        self._check(self.L.asc_extrinsic_marginal(self.h, _ptr(cov), C.byref(its)))
        return cov, its.value

    def optimize(self, mode=D.ASC_BATCH_LM):
        cap = max(1, int(self.params.max_iterations))
        log = (D.AscIterLog * cap)()
        n = C.c_int32()
        err = C.c_double()
        rc = self.L.asc_optimize(self.h, mode, log, cap, C.byref(n), C.byref(err))
        self.log = [log[i] for i in range(min(n.value, cap))]
        self._error = err.value
        self._check(rc)
        return self.values()

    def error(self):
        return self._error if self._error is not None else self.graph_error()

    def iterations(self):
        return len(self.log)

    def save_values(self):
        self._check(self.L.asc_save_values(self.h))

    def restore_values(self):
        self._check(self.L.asc_restore_values(self.h))

    def launch_count(self):
        return int(self.L.asc_launch_count(self.h))

    def timer_start(self):
        self._check(self.L.asc_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.L.asc_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_begin(self, kernel_id, max_samples=256):
        self._check(self.L.asc_profile_begin(self.h, kernel_id, max_samples))

    def profile_read(self):
        n, mean, mn = C.c_int32(), C.c_double(), C.c_double()
        self._check(self.L.asc_profile_read(self.h, C.byref(n), C.byref(mean), C.byref(mn)))
        return n.value, mean.value, mn.value

    def close(self):
        if self.h:
            self.L.asc_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
