"""Prototype 8: V-cycle quality at a LATER LM iteration (stiffer system: weights near 1, lambda 1e-9).  usage: ml_proto8.py <scale> <lm_iters>"""
import sys, time, os
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ml_proto
from ml_proto import clusters_of, aggregate, pcg
import arm_slam_calib_b200 as asc
from oracle import Oracle


def main():
    scale = float(sys.argv[1]); k_lm = int(sys.argv[2]); lam = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-9
    from arm_slam_calib_b200.sim import make_sim_problem, pose12
    n = int(round(900 * np.sqrt(scale)))
    pb = make_sim_problem(trajectory_size=int(round(10000 * scale)), nx=n, ny=n, size_x=1.3, size_y=1.3, z_range=(-0.9, 1.6),
                          max_range=0.27 * scale ** (-1.0 / 3.0), segment_length=10, seed=0, extrinsic_guess=pose12(t=[0.02, 0.02, 0.02]))
    prm = asc.default_params(); prm.max_iterations = k_lm
    r = Oracle(pb, prm).optimize(pb.q0, pb.l0, pb.x0, mode=0, solver=1)
    print("after %d LM iterations: error %.6g, pcg its %s" % (k_lm, r["error"], [g.pcg_iterations for g in r["log"]]))
    import dataclasses
    pb2 = dataclasses.replace(pb, q0=r["q"], l0=r["l"], x0=r["x"])
    Sqq, SqK, SKK, rhs, N, P = ml_proto.build_S(pb2, lam)
    PN = P * N
    def apply_S(x):
        return np.concatenate([Sqq @ x[:PN] + SqK @ x[PN:], SqK.T @ x[:PN] + SKK @ x[PN:]])
    cl_ptr = clusters_of(pb, N, P); nC = len(cl_ptr) - 1
    blocks = [Sqq[cl_ptr[c] * N:cl_ptr[c + 1] * N, cl_ptr[c] * N:cl_ptr[c + 1] * N].toarray() for c in range(nC)]
    Minv = [np.linalg.inv(b) for b in blocks]; MK = np.linalg.inv(SKK)
    def M_cl(r):
        z = np.empty_like(r)
        for c in range(nC):
            a, b = cl_ptr[c] * N, cl_ptr[c + 1] * N
            z[a:b] = Minv[c] @ r[a:b]
        z[PN:] = MK @ r[PN:]
        return z
    def run(name, M):
        t0 = time.time(); x, it = pcg(apply_S, M, rhs, maxit=3000); print("%-40s: %4d iterations (%.1f s)" % (name, it, time.time() - t0)); sys.stdout.flush()
    pose_cl = np.repeat(np.arange(nC), np.diff(cl_ptr))
    r_ = np.arange(PN); c_ = pose_cl[r_ // N] * N + r_ % N
    Z1 = sp.csr_matrix((np.ones(PN), (r_, c_)), shape=(PN, nC * N))
    E1 = (Z1.T @ Sqq @ Z1).tocsr(); E1K = Z1.T @ SqK; n1 = nC * N
    run("cluster-Jacobi only", M_cl)
    Efull = np.block([[E1.toarray(), E1K], [E1K.T, SKK]]); cf = sla.cho_factor(Efull)
    def M_exact(r):
        zc = sla.cho_solve(cf, np.concatenate([Z1.T @ r[:PN], r[PN:]]))
        return M_cl(r) + np.concatenate([Z1 @ zc[:n1], zc[n1:]])
    run("exact coarse", M_exact)
    d = sp.bsr_matrix(E1, blocksize=(N, N)); diag = np.zeros((nC, N, N))
    for i in range(nC):
        for k in range(d.indptr[i], d.indptr[i + 1]):
            if d.indices[k] == i: diag[i] = d.data[k]
    Dinv = sp.bsr_matrix((np.linalg.inv(diag), np.arange(nC), np.arange(nC + 1)), shape=(n1, n1)).tocsr()
    w = sp.csr_matrix((np.sqrt((d.data ** 2).sum(axis=(1, 2))), d.indices, d.indptr), shape=(nC, nC)); w.setdiag(0); w.eliminate_zeros()
    agg, na = aggregate(w, 16)
    r2 = np.arange(n1); c2 = agg[r2 // N] * N + r2 % N
    Pm = sp.csr_matrix((np.ones(n1), (r2, c2)), shape=(n1, na * N))
    Ec = (Pm.T @ E1 @ Pm).toarray(); cfc = sla.cho_factor(Ec)
    v = np.random.default_rng(0).standard_normal(n1)
    for _ in range(16): v = Dinv @ (E1 @ v); lmax = np.linalg.norm(v); v /= lmax
    print("lmax %.3f, aggregates %d" % (lmax, na))
    FR = {1: 4, 2: 4, 3: 8, 4: 10, 5: 15, 6: 20, 8: 30, 10: 40, 12: 50}
    def cheb(r, z0, m, lo, hi):
        Dinv = cheb.D
        theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo); sigma = theta / delta; rho = 1.0 / sigma
        z = np.zeros_like(r) if z0 is None else z0
        res = r if z0 is None else r - E1 @ z
        dd = (Dinv @ res) / theta; z = z + dd
        for _ in range(m - 1):
            res = r - E1 @ z; rn = 1.0 / (2 * sigma - rho)
            dd = rn * rho * dd + (2 * rn / delta) * (Dinv @ res); z = z + dd; rho = rn
        return z
    # block smoother: D = diagonal blocks of E over aggregates of 8 clusters
    def block_dinv(gs_s):
        aggs, nas = aggregate(w, gs_s)
        rr, cc, vv = [], [], []
        for a in range(nas):
            mem = np.where(aggs == a)[0]
            ix = np.concatenate([np.arange(m_ * N, m_ * N + N) for m_ in mem])
            Di = np.linalg.inv(E1[ix][:, ix].toarray())
            R, Cc = np.meshgrid(ix, ix, indexing="ij"); rr.append(R.ravel()); cc.append(Cc.ravel()); vv.append(Di.ravel())
        Db = sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(n1, n1))
        v = np.random.default_rng(0).standard_normal(n1)
        for _ in range(16): v = Db @ (E1 @ v); lm = np.linalg.norm(v); v /= lm
        return Db, lm, nas
    variants = [("point", Dinv, lmax)]
    for gs_s in (4, 8, 16):
        Db, lmb, nas = block_dinv(gs_s)
        print("block smoother gs=%d: %d blocks, lmax %.3f" % (gs_s, nas, lmb))
        variants.append(("block%d" % gs_s, Db, lmb))
    for vname, Dv, lmv in variants:
      cheb.D = Dv
      for m in (2, 3, 4, 6):
        hi = 1.2 * lmv; lo = hi / FR[m]
        def V(r):
            z = cheb(r, None, m, lo, hi)
            z = z + Pm @ sla.cho_solve(cfc, Pm.T @ (r - E1 @ z))
            return cheb(r, z, m, lo, hi)
        Y = np.stack([V(E1K[:, i]) for i in range(6)], axis=1); Sk = SKK - E1K.T @ Y; Ski = np.linalg.inv(0.5 * (Sk + Sk.T))
        def M(r):
            t = V(Z1.T @ r[:PN]); zK = Ski @ (r[PN:] - E1K.T @ t)
            return M_cl(r) + np.concatenate([Z1 @ (t - Y @ zK), zK])
        run("V-cycle %s cheb m=%d" % (vname, m), M)


if __name__ == "__main__":
    main()
