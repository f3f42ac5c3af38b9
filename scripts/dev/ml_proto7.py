"""Prototype 7: ADDITIVE multilevel with Chebyshev-polynomial level solvers (all levels concurrent: m + 2 dependent phases instead of
~4 m per V-cycle), bordered K.  usage: ml_proto7.py <scale> <lambda> <top_dim>"""
import sys, time, os
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ml_proto import build_S, clusters_of, aggregate, pcg


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
    lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-5
    top_dim = int(sys.argv[3]) if len(sys.argv) > 3 else 600
    from arm_slam_calib_b200.sim import make_sim_problem, pose12
    n = int(round(900 * np.sqrt(scale)))
    pb = make_sim_problem(trajectory_size=int(round(10000 * scale)), nx=n, ny=n, size_x=1.3, size_y=1.3, z_range=(-0.9, 1.6),
                          max_range=0.27 * scale ** (-1.0 / 3.0), segment_length=10, seed=0, extrinsic_guess=pose12(t=[0.02, 0.02, 0.02]))
    Sqq, SqK, SKK, rhs, N, P = build_S(pb, lam)
    PN = P * N
    def apply_S(x):
        return np.concatenate([Sqq @ x[:PN] + SqK @ x[PN:], SqK.T @ x[:PN] + SKK @ x[PN:]])
    cl_ptr = clusters_of(pb, N, P)
    nC = len(cl_ptr) - 1
    blocks = [Sqq[cl_ptr[c] * N:cl_ptr[c + 1] * N, cl_ptr[c] * N:cl_ptr[c + 1] * N].toarray() for c in range(nC)]
    Minv = [np.linalg.inv(b) for b in blocks]
    MK = np.linalg.inv(SKK)
    def M_cl(r):
        z = np.empty_like(r)
        for c in range(nC):
            a, b = cl_ptr[c] * N, cl_ptr[c + 1] * N
            z[a:b] = Minv[c] @ r[a:b]
        z[PN:] = MK @ r[PN:]
        return z
    def run(name, M):
        t0 = time.time(); x, it = pcg(apply_S, M, rhs, maxit=3000); print("%-56s: %4d iterations (%.1f s)" % (name, it, time.time() - t0)); sys.stdout.flush()
    pose_cl = np.repeat(np.arange(nC), np.diff(cl_ptr))
    r_ = np.arange(PN); c_ = pose_cl[r_ // N] * N + r_ % N
    Z1 = sp.csr_matrix((np.ones(PN), (r_, c_)), shape=(PN, nC * N))
    E1 = (Z1.T @ Sqq @ Z1).tocsr(); E1K = Z1.T @ SqK
    n1 = nC * N

    def point_block_inv(E, nn):
        d = sp.bsr_matrix(E, blocksize=(N, N))
        diag = np.zeros((nn, N, N))
        for i in range(nn):
            for k in range(d.indptr[i], d.indptr[i + 1]):
                if d.indices[k] == i:
                    diag[i] = d.data[k]
        return sp.bsr_matrix((np.linalg.inv(diag), np.arange(nn), np.arange(nn + 1)), shape=(nn * N, nn * N)).tocsr()

    for gs in (8, 4):
        levels = []
        E = E1; nn = nC
        Pacc = sp.identity(n1, format="csr")      # accumulated prolongator level -> level 0
        while nn * N > top_dim:
            bs = sp.bsr_matrix(E, blocksize=(N, N))
            w = sp.csr_matrix((np.sqrt((bs.data ** 2).sum(axis=(1, 2))), bs.indices, bs.indptr), shape=(nn, nn))
            w.setdiag(0); w.eliminate_zeros()
            agg, na = aggregate(w, gs)
            r2 = np.arange(nn * N); c2 = agg[r2 // N] * N + r2 % N
            Pm = sp.csr_matrix((np.ones(nn * N), (r2, c2)), shape=(nn * N, na * N))
            Dinv = point_block_inv(E, nn)
            v = np.random.default_rng(0).standard_normal(E.shape[0])
            for _ in range(16):
                v = Dinv @ (E @ v); lmax = np.linalg.norm(v); v /= lmax
            levels.append((E, Dinv, Pacc, lmax))
            Pacc = (Pacc @ Pm).tocsr()
            E = (Pm.T @ E @ Pm).tocsr(); nn = na
        cf = sla.cho_factor(E.toarray())
        Ptop = Pacc
        print("gs %d: %d smoothed levels, top dim %d" % (gs, len(levels), E.shape[0]))

        def cheb_solve(E, Dinv, r, m, lo, hi):
            theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
            sigma = theta / delta; rho = 1.0 / sigma
            d = (Dinv @ r) / theta
            z = d.copy()
            for _ in range(m - 1):
                res = r - E @ z
                rho_n = 1.0 / (2 * sigma - rho)
                d = rho_n * rho * d + (2 * rho_n / delta) * (Dinv @ res)
                z = z + d; rho = rho_n
            return z

        for m, frac in ((2, 4), (3, 8), (4, 10), (6, 20), (8, 30)):
            def V(r):
                z = Ptop @ sla.cho_solve(cf, Ptop.T @ r)
                for E_, Dinv, Pa, lmax in levels:
                    hi = 1.2 * lmax
                    z = z + Pa @ cheb_solve(E_, Dinv, Pa.T @ r, m, hi / frac, hi)
                return z
            Y = np.stack([V(E1K[:, i]) for i in range(6)], axis=1)
            Sk = SKK - E1K.T @ Y
            Sk = 0.5 * (Sk + Sk.T)
            ev = np.linalg.eigvalsh(Sk).min()
            if ev <= 0:
                print("   additive m=%d: S_k not positive definite (min eig %.3g): V is not <= E^-1 -> need damping" % (m, ev))
                # damp V so that V <= E^-1: scale by 1 / lambda_max(V E)
                v = np.random.default_rng(1).standard_normal(n1)
                for _ in range(30):
                    v = V(E1 @ v); lm = np.linalg.norm(v); v /= lm
                sc = 1.0 / (1.05 * lm)
                Y = sc * Y; Sk = SKK - E1K.T @ Y; Sk = 0.5 * (Sk + Sk.T)
                print("   lambda_max(V E) = %.3f -> scale %.3f, min eig S_k %.3g" % (lm, sc, np.linalg.eigvalsh(Sk).min()))
            else:
                sc = 1.0
            Ski = np.linalg.inv(Sk)
            def M(r):
                rc = Z1.T @ r[:PN]
                t = sc * V(rc)
                zK = Ski @ (r[PN:] - E1K.T @ t)
                z = t - Y @ zK
                return M_cl(r) + np.concatenate([Z1 @ z, zK])
            run("additive: gs=%d levels=%d cheb m=%d lo=hi/%d" % (gs, len(levels), m, frac), M)


if __name__ == "__main__":
    main()
