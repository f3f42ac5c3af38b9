"""Prototype 3: level-2 aggregation in world-twist coordinates (rigid-body candidates, SA-AMG style).  usage: ml_proto3.py <scale>"""
import sys, time, os
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ml_proto import build_S, clusters_of, aggregate, pcg
import arm_slam_calib_b200 as asc
from oracle import Oracle


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
        t0 = time.time(); x, it = pcg(apply_S, M, rhs, maxit=3000); print("%-40s: %4d iterations (%.1f s)" % (name, it, time.time() - t0)); sys.stdout.flush()
    pose_cl = np.repeat(np.arange(nC), np.diff(cl_ptr))
    r_ = np.arange(PN); c_ = pose_cl[r_ // N] * N + r_ % N
    Z1 = sp.csr_matrix((np.ones(PN), (r_, c_)), shape=(PN, nC * N))
    E1 = (Z1.T @ Sqq @ Z1).tocsr(); E1K = Z1.T @ SqK
    n1 = nC * N
    # per-cluster twist Jacobian at the cluster's middle keyframe: columns [a_i ; o_i x a_i]
    orc = Oracle(pb, asc.default_params())
    Jc = []
    for c in range(nC):
        p = (cl_ptr[c] + cl_ptr[c + 1]) // 2
        t, a, o = orc.fk(pb.q0[p])
        Jc.append(np.concatenate([a.T, np.cross(o, a).T], axis=0))     # 6 x N
    bs = sp.bsr_matrix(E1, blocksize=(N, N))
    wts = sp.csr_matrix((np.sqrt((bs.data ** 2).sum(axis=(1, 2))), bs.indices, bs.indptr), shape=(nC, nC))
    wts.setdiag(0); wts.eliminate_zeros()

    def vcycle_precond(levels, Etop_cf, ntop, om, nu):
        def V(li, r, rK):
            if li == len(levels):
                s = sla.cho_solve(Etop_cf, np.concatenate([r, rK]))
                return s[:ntop], s[ntop:]
            E, EK, Dinv, Pm = levels[li]
            z = om * (Dinv @ r)
            for _ in range(nu - 1):
                z = z + om * (Dinv @ (r - E @ z))
            zc, zK = V(li + 1, Pm.T @ (r - E @ z), rK - EK.T @ z)
            z = z + Pm @ zc
            for _ in range(nu):
                z = z + om * (Dinv @ (r - E @ z - EK @ zK))
            return z, zK
        def M(r):
            zc, zK = V(0, Z1.T @ r[:PN], r[PN:])
            return M_cl(r) + np.concatenate([Z1 @ zc, zK])
        return M

    def block_jacobi(E, idx):
        rr, cc, vv = [], [], []
        for ix in idx:
            Di = np.linalg.inv(E[ix][:, ix].toarray())
            R, Cc = np.meshgrid(ix, ix, indexing="ij")
            rr.append(R.ravel()); cc.append(Cc.ravel()); vv.append(Di.ravel())
        return sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=E.shape)

    for gs_list in ((8,), (16,), (32,), (4, 4), (8, 8)):
        levels = []
        E, EK, nn, dof = E1, E1K, nC, N
        w = wts
        first = True
        for gs in gs_list:
            agg, na = aggregate(w, gs)
            members = [np.where(agg == a)[0] for a in range(na)]
            idx = [np.concatenate([np.arange(m * dof, (m + 1) * dof) for m in mem]) for mem in members]
            Dinv = block_jacobi(E, idx)
            rr, cc, vv = [], [], []
            for m in range(nn):
                Pc = np.linalg.pinv(Jc[m]) if first else np.eye(6)      # dof x 6
                R, Cc = np.meshgrid(np.arange(m * dof, (m + 1) * dof), np.arange(agg[m] * 6, agg[m] * 6 + 6), indexing="ij")
                rr.append(R.ravel()); cc.append(Cc.ravel()); vv.append(Pc.ravel())
            Pm = sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(nn * dof, na * 6))
            levels.append((E, EK, Dinv, Pm))
            E = (Pm.T @ E @ Pm).tocsr(); EK = Pm.T @ EK; nn = na; dof = 6; first = False
            b2 = sp.bsr_matrix(E, blocksize=(6, 6))
            w = sp.csr_matrix((np.sqrt((b2.data ** 2).sum(axis=(1, 2))), b2.indices, b2.indptr), shape=(na, na))
            w.setdiag(0); w.eliminate_zeros()
            print("   level -> %d aggregates, blocks/row %.1f" % (na, b2.data.shape[0] / na))
        Etop = np.block([[E.toarray(), EK], [EK.T, SKK]])
        cf = sla.cho_factor(Etop)
        for om, nu in ((1.0, 1), (0.8, 1), (0.8, 2)):
            run("twist-agg gs=%s top %d w=%.1f nu=%d" % (gs_list, nn * 6 + 6, om, nu), vcycle_precond(levels, cf, nn * 6, om, nu))


if __name__ == "__main__":
    main()
