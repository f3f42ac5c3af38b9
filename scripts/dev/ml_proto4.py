"""Prototype 4: Chebyshev polynomial smoothers in the coarse V-cycle (plain piecewise-constant aggregation).  usage: ml_proto4.py <scale>"""
import sys, time, os
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ml_proto import build_S, clusters_of, aggregate, pcg


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
    lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-5
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
    # the K border is eliminated exactly on the coarse level: E = [E1 E1K; E1K^T SKK]; work with the full bordered matrix
    Efull = sp.bmat([[E1, sp.csr_matrix(E1K)], [sp.csr_matrix(E1K.T), sp.csr_matrix(SKK)]]).tocsr()
    bs = sp.bsr_matrix(E1, blocksize=(N, N))
    wts = sp.csr_matrix((np.sqrt((bs.data ** 2).sum(axis=(1, 2))), bs.indices, bs.indptr), shape=(nC, nC))
    wts.setdiag(0); wts.eliminate_zeros()

    def block_jacobi(E, idx, nfull):
        rr, cc, vv = [], [], []
        for ix in idx:
            Di = np.linalg.inv(E[ix][:, ix].toarray())
            R, Cc = np.meshgrid(ix, ix, indexing="ij")
            rr.append(R.ravel()); cc.append(Cc.ravel()); vv.append(Di.ravel())
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(nfull, nfull))

    for gs in (1, 8):
        agg, na = aggregate(wts, gs) if gs > 1 else (np.arange(nC), nC)
        members = [np.where(agg == a)[0] for a in range(na)]
        idx = [np.concatenate([np.arange(m * N, (m + 1) * N) for m in mem]) for mem in members] + [np.arange(n1, n1 + 6)]
        Dinv = block_jacobi(Efull, idx, n1 + 6)
        # lambda_max of Dinv E by power iteration
        v = np.random.default_rng(0).standard_normal(n1 + 6)
        for _ in range(40):
            v = Dinv @ (Efull @ v); lmax = np.linalg.norm(v); v /= lmax
        print("gs %d: lambda_max(Dinv E) = %.3f" % (gs, lmax))
        # coarse level: aggregates of 8 (independent of the smoother's blocks)
        agg2, na2 = aggregate(wts, 8)
        r2 = np.arange(n1); c2 = agg2[r2 // N] * N + r2 % N
        Pm = sp.csr_matrix((np.ones(n1), (r2, c2)), shape=(n1, na2 * N))
        Pf = sp.bmat([[Pm, None], [None, sp.identity(6)]]).tocsr()
        E2 = (Pf.T @ Efull @ Pf).toarray()
        cf = sla.cho_factor(E2)

        def cheb(r, z0, m, lo, hi):
            """m steps of Chebyshev on Dinv E for E z = r starting from z0 (z0 may be None = 0)"""
            theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
            z = np.zeros_like(r) if z0 is None else z0
            res = r if z0 is None else r - Efull @ z
            sigma = theta / delta; rho = 1.0 / sigma
            d = (Dinv @ res) / theta
            z = z + d
            for _ in range(m - 1):
                res = r - Efull @ z
                rho_n = 1.0 / (2 * sigma - rho)
                d = rho_n * rho * d + (2 * rho_n / delta) * (Dinv @ res)
                z = z + d; rho = rho_n
            return z

        for m, frac in ((1, 4), (2, 4), (3, 8), (4, 10), (6, 20)):
            hi = 1.1 * lmax; lo = hi / frac
            def V(r):
                z = cheb(r, None, m, lo, hi)
                z = z + Pf @ sla.cho_solve(cf, Pf.T @ (r - Efull @ z))
                z = cheb(r, z, m, lo, hi)
                return z
            def M(r):
                zc = V(np.concatenate([Z1.T @ r[:PN], r[PN:]]))
                return M_cl(r) + np.concatenate([Z1 @ zc[:n1], zc[n1:]])
            run("smoother blocks gs=%d cheb m=%d lo=hi/%d" % (gs, m, frac), M)


if __name__ == "__main__":
    main()
