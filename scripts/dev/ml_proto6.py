"""Prototype 6: V-cycle on the cluster part E_cc only, extrinsic border K eliminated by bordering (Y = V E_cK, S_k = S_KK - E_Kc Y).
usage: ml_proto6.py <scale> <lambda> <top_dim>"""
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
        t0 = time.time(); x, it = pcg(apply_S, M, rhs, maxit=3000); print("%-46s: %4d iterations (%.1f s)" % (name, it, time.time() - t0)); sys.stdout.flush()
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

    # pair-count-like static weights: use block norms here
    for gs in (8, 16, 32, 64):
        levels = []
        E = E1; nn = nC
        while nn * N > top_dim:
            bs = sp.bsr_matrix(E, blocksize=(N, N))
            w = sp.csr_matrix((np.sqrt((bs.data ** 2).sum(axis=(1, 2))), bs.indices, bs.indptr), shape=(nn, nn))
            w.setdiag(0); w.eliminate_zeros()
            agg, na = aggregate(w, gs)
            r2 = np.arange(nn * N); c2 = agg[r2 // N] * N + r2 % N
            Pm = sp.csr_matrix((np.ones(nn * N), (r2, c2)), shape=(nn * N, na * N))
            Dinv = point_block_inv(E, nn)
            v = np.random.default_rng(0).standard_normal(E.shape[0])
            hist = []
            for _ in range(30):
                v = Dinv @ (E @ v); lmax = np.linalg.norm(v); v /= lmax; hist.append(lmax)
            print("  level %d nodes lmax after 8/12/16/20/30 its: %.3f %.3f %.3f %.3f %.3f" % (nn, hist[7], hist[11], hist[15], hist[19], hist[29]))
            levels.append((E, Dinv, Pm, hist[15]))
            E = (Pm.T @ E @ Pm).tocsr(); nn = na
        cf = sla.cho_factor(E.toarray())
        print("  top dim %d" % E.shape[0])

        def cheb(E, Dinv, r, z0, m, lo, hi):
            theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
            z = np.zeros_like(r) if z0 is None else z0
            res = r if z0 is None else r - E @ z
            sigma = theta / delta; rho = 1.0 / sigma
            d = (Dinv @ res) / theta
            z = z + d
            for _ in range(m - 1):
                res = r - E @ z
                rho_n = 1.0 / (2 * sigma - rho)
                d = rho_n * rho * d + (2 * rho_n / delta) * (Dinv @ res)
                z = z + d; rho = rho_n
            return z

        FR = {1: 4, 2: 4, 3: 8, 4: 10, 5: 15, 6: 20, 8: 30}
        for mm in ((4, 4),):
            def V(li, r):
                if li == len(levels):
                    return sla.cho_solve(cf, r)
                E, Dinv, Pm, lmax = levels[li]
                m = mm[0] if li == 0 else mm[1]
                frac = FR[m]
                hi = 1.2 * lmax; lo = hi / frac
                z = cheb(E, Dinv, r, None, m, lo, hi)
                z = z + Pm @ V(li + 1, Pm.T @ (r - E @ z))
                return cheb(E, Dinv, r, z, m, lo, hi)
            Y = np.stack([V(0, E1K[:, i]) for i in range(6)], axis=1)
            Sk = SKK - E1K.T @ Y
            Ski = np.linalg.inv(Sk)
            print("   Sk eig min %.3g (SKK eig min %.3g)" % (np.linalg.eigvalsh(0.5 * (Sk + Sk.T)).min(), np.linalg.eigvalsh(SKK).min()))
            def M(r):
                rc = Z1.T @ r[:PN]
                t = V(0, rc)
                zK = Ski @ (r[PN:] - E1K.T @ t)
                z = t - Y @ zK
                return M_cl(r) + np.concatenate([Z1 @ z, zK])
            run("bordered K: gs=%d levels=%d cheb m=%s" % (gs, len(levels), mm), M)


if __name__ == "__main__":
    main()
