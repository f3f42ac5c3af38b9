"""ctypes binding of oracle/libasc_oracle.so (the C restatement of the reference path)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libasc_oracle.so")
_lib = None
_vp = C.c_void_p


def build_oracle(force=False):
    src = os.path.join(_HERE, "asc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(_LIB_PATH):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True, env=env)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build_oracle()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = _vp
        L.orc_create.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                 C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_destroy.argtypes = [_vp]
        L.orc_error.restype = C.c_double
        L.orc_error.argtypes = [_vp, _vp, _vp, _vp]
        L.orc_linearize.restype = C.c_double
        L.orc_linearize.argtypes = [_vp] + [_vp] * 16
        L.orc_set_links.restype = None
        L.orc_set_links.argtypes = [_vp, C.c_int, _vp, _vp, _vp]
        L.orc_extrinsic_marginal.restype = C.c_int
        L.orc_extrinsic_marginal.argtypes = [_vp, _vp, _vp, _vp, _vp]
        L.orc_solve_damped.restype = C.c_int
        L.orc_solve_damped.argtypes = [_vp, _vp, _vp, _vp, C.c_double, C.c_int, _vp, _vp, _vp, C.POINTER(C.c_int)]
        L.orc_optimize.restype = C.c_int
        L.orc_optimize.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_project.restype = C.c_int
        L.orc_project.argtypes = [_vp] * 8
        L.orc_fk.argtypes = [_vp] * 5
        L.orc_pose3_retract.argtypes = [C.c_int, _vp, _vp, _vp]
        L.orc_pose3_local.argtypes = [C.c_int, _vp, _vp, _vp]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_fail_key.argtypes = [_vp, C.POINTER(C.c_char), C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """Oracle over one flattened problem (same fields as arm_slam_calib_b200.Problem) and one asc_params."""

    def __init__(self, pb, params):
        self.L = _load()
        self.pb, self.params = pb, params
        self.N, self.P, self.Lm, self.M = pb.n_dof, pb.enc.shape[0], pb.lm_prior.shape[0], pb.pose_idx.shape[0]
        k = dict(t_p2j=_f64(pb.t_p2j), axis=_f64(pb.axis), t_c2j=_f64(pb.t_c2j), t_base=_f64(pb.t_base),
                 cont=None if pb.continuous is None else np.ascontiguousarray(pb.continuous, dtype=np.uint8),
                 pi=np.ascontiguousarray(pb.pose_idx, dtype=np.int32), li=np.ascontiguousarray(pb.lm_idx, dtype=np.int32),
                 uv=_f64(pb.uv), enc=_f64(pb.enc), he=None if pb.has_enc is None else np.ascontiguousarray(pb.has_enc, dtype=np.uint8),
                 lp=_f64(pb.lm_prior), xp=_f64(pb.x_prior))
        self.h = self.L.orc_create(C.byref(params), self.N, _p(k["t_p2j"]), _p(k["axis"]), _p(k["t_c2j"]), _p(k["cont"]), _p(k["t_base"]),
                                   pb.fx, pb.fy, pb.cx, pb.cy, self.P, self.Lm, self.M, _p(k["pi"]), _p(k["li"]), _p(k["uv"]),
                                   _p(k["enc"]), _p(k["he"]), _p(k["lp"]), _p(k["xp"]))
        if getattr(pb, "has_drift", None) is not None:
            self.set_links(0, pb.has_drift, pb.drift_meas, pb.sigma_drift)
        if getattr(pb, "has_vel", None) is not None:
            self.set_links(1, pb.has_vel, pb.vel_meas, pb.sigma_vel)

    def set_links(self, kind, has, meas, sigma):
        """Link factors on (q_{p-1}, q_p) where has[p] != 0: kind 0 = DriftFactor (DriftFactor.h:44-61, meas = q1Encoders - q0Encoders),
        kind 1 = VelocityFactor (VelocityFactor.h:40-62, meas = qDot)."""
        if has is None:
            self.L.orc_set_links(self.h, kind, None, None, None)
            return
        hd = np.ascontiguousarray(has, dtype=np.uint8); m = _f64(meas); sg = _f64(np.broadcast_to(np.asarray(sigma, dtype=np.float64), (self.N,)))
        self.L.orc_set_links(self.h, kind, _p(hd), _p(m), _p(sg))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def set_threads(n):
        _load().orc_set_threads(n)

    @staticmethod
    def max_threads():
        return _load().orc_max_threads()

    def fk(self, q):
        q = _f64(q); t = np.empty(12); a = np.empty((self.N, 3)); o = np.empty((self.N, 3))
        self.L.orc_fk(self.h, _p(q), _p(t), _p(a), _p(o))
        return t, a, o

    def project(self, q, l, x, jac=True):
        q, l, x = _f64(q), _f64(l), _f64(x)
        uv = np.empty(2)
        if not jac:
            ch = self.L.orc_project(self.h, _p(q), _p(l), _p(x), _p(uv), None, None, None)
            return uv, ch
        J1 = np.empty((2, self.N)); J2 = np.empty((2, 3)); J3 = np.empty((2, 6))
        ch = self.L.orc_project(self.h, _p(q), _p(l), _p(x), _p(uv), _p(J1), _p(J2), _p(J3))
        return uv, J1, J2, J3, ch

    def error(self, q, l, x):
        q, l, x = _f64(q), _f64(l), _f64(x)
        return self.L.orc_error(self.h, _p(q), _p(l), _p(x))

    def linearize(self, q, l, x):
        q, l, x = _f64(q), _f64(l), _f64(x)
        M, N, P, Lm = self.M, self.N, self.P, self.Lm
        o = dict(b=np.empty((M, 2)), jq=np.empty((M, 2, N)), jl=np.empty((M, 2, 3)), jk=np.empty((M, 2, 6)), c=np.empty((Lm, 3, 3)),
                 gl=np.empty((Lm, 3)), wk=np.empty((Lm, 6, 3)), bpp=np.empty((P, N, N)), bpk=np.empty((P, N, 6)), gp=np.empty((P, N)),
                 bkk=np.empty((6, 6)), gk=np.empty(6))
        ch = C.c_int64()
        e = self.L.orc_linearize(self.h, _p(q), _p(l), _p(x), *[_p(o[k]) for k in ("b", "jq", "jl", "jk", "c", "gl", "wk", "bpp", "bpk", "gp", "bkk", "gk")],
                                 C.cast(C.byref(ch), _vp))
        o["error"], o["cheirality"] = e, ch.value
        return o

    def solve_damped(self, q, l, x, lam, solver=0):
        q, l, x = _f64(q), _f64(l), _f64(x)
        dq = np.empty((self.P, self.N)); dl = np.empty((self.Lm, 3)); dx = np.empty(6)
        its = C.c_int()
        rc = self.L.orc_solve_damped(self.h, _p(q), _p(l), _p(x), float(lam), solver, _p(dq), _p(dl), _p(dx), C.byref(its))
        return rc, dq, dl, dx, its.value

    def extrinsic_marginal(self, q, l, x):
        """K0 block of (A^T A)^-1 at the given values (6x6, tangent order [w, v]); rc as in solve_damped."""
        q, l, x = _f64(q), _f64(l), _f64(x)
        cov = np.empty((6, 6))
        rc = self.L.orc_extrinsic_marginal(self.h, _p(q), _p(l), _p(x), _p(cov))
        return rc, cov

    def optimize(self, q, l, x, mode=0, solver=0):
        from arm_slam_calib_b200._ctypes_defs import AscIterLog
        q, l, x = _f64(q).copy(), _f64(l).copy(), _f64(x).copy()
        cap = max(1, int(self.params.max_iterations))
        log = (AscIterLog * cap)()
        n = C.c_int(); err = C.c_double()
        rc = self.L.orc_optimize(self.h, mode, solver, _p(q), _p(l), _p(x), C.cast(log, _vp), cap, C.byref(n), C.byref(err))
        return dict(rc=rc, q=q, l=l, x=x, iterations=n.value, error=err.value, log=[log[i] for i in range(min(n.value, cap))])

    def fail_key(self):
        kind, idx = C.c_char(), C.c_int64()
        self.L.orc_fail_key(self.h, C.byref(kind), C.byref(idx))
        return kind.value.decode() if kind.value else "", idx.value

    @staticmethod
    def pose3_retract(chart, x, xi):
        x, xi = _f64(x), _f64(xi); out = np.empty(12)
        _load().orc_pose3_retract(chart, _p(x), _p(xi), _p(out))
        return out

    @staticmethod
    def pose3_local(chart, x0, x1):
        x0, x1 = _f64(x0), _f64(x1); out = np.empty(6)
        _load().orc_pose3_local(chart, _p(x0), _p(x1), _p(out))
        return out
