"""Numerical prototype (CPU, scipy) of the coarse-level variants of the PCG preconditioner.
Builds the reduced system S explicitly for a scaled-down C2 and counts PCG iterations to 1e-12 for:
  cl        cluster-Jacobi only
  2lv       cluster-Jacobi + exact coarse solve (round-1 GPU preconditioner)
  ml-add    additive multilevel (block-Jacobi on aggregates of every coarse level + dense top level)
  ml-v      symmetric V-cycle on the coarse levels (damped block-Jacobi smoother), added to cluster-Jacobi
usage: python scripts/dev/ml_proto.py <scale> [lambda]"""
import sys, time, os
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import arm_slam_calib_b200 as asc
from arm_slam_calib_b200 import configs
from oracle import Oracle


def build_S(pb, lam):
    prm = asc.default_params()
    orc = Oracle(pb, prm)
    lin = orc.linearize(pb.q0, pb.l0, pb.x0)
    N, P, L, M = orc.N, orc.P, orc.Lm, orc.M
    pi, li = pb.pose_idx.astype(np.int64), pb.lm_idx.astype(np.int64)
    jq, jl = lin["jq"], lin["jl"]
    # W_m = jq^T jl  (N x 3) per observation
    Wm = np.einsum("mra,mrc->mac", jq, jl)                       # [M, N, 3]
    rows = (pi[:, None, None] * N + np.arange(N)[None, :, None]) + np.zeros((1, 1, 3), np.int64)
    cols = (li[:, None, None] * 3 + np.arange(3)[None, None, :]) + np.zeros((1, N, 1), np.int64)
    Hql = sp.csr_matrix((Wm.ravel(), (rows.ravel(), cols.ravel())), shape=(P * N, 3 * L))
    Cd = lin["c"] + lam * np.eye(3)[None]
    Cinv = np.linalg.inv(Cd)
    Cinv_bd = sp.bsr_matrix((Cinv, np.arange(L), np.arange(L + 1)), shape=(3 * L, 3 * L)).tocsr()
    Bpp = sp.bsr_matrix((lin["bpp"] + lam * np.eye(N)[None], np.arange(P), np.arange(P + 1)), shape=(P * N, P * N)).tocsr()
    t0 = time.time()
    T = Hql @ Cinv_bd
    Sqq = (Bpp - T @ Hql.T).tocsr()
    WK = lin["wk"].transpose(1, 0, 2).reshape(6, 3 * L)              # [6, 3L]
    SqK = lin["bpk"].reshape(P * N, 6) - T @ WK.T
    SKK = lin["bkk"] + lam * np.eye(6) - WK @ (Cinv_bd @ WK.T)
    cg = (Cinv_bd @ lin["gl"].reshape(-1))
    rhs_q = lin["gp"].reshape(-1) - Hql @ cg
    rhs_K = lin["gk"] - WK @ cg
    print("S built: nnz %d (%.1f s)" % (Sqq.nnz, time.time() - t0))
    return Sqq, SqK, SKK, np.concatenate([rhs_q, rhs_K]), N, P


def clusters_of(pb, N, P, gmax=10, covis=0.25):
    order = np.argsort(pb.pose_idx, kind="stable")
    pp = pb.pose_idx[order]; lm = pb.lm_idx[order]
    ptr = np.searchsorted(pp, np.arange(P + 1))
    cl_ptr = []; size = 0; prev = set()
    for p in range(P):
        cur = set(lm[ptr[p]:ptr[p + 1]].tolist())
        shared = len(cur & prev)
        mn = min(len(cur), len(prev)) if p > 0 else 0
        joined = p > 0 and size < gmax and mn > 0 and shared >= covis * mn
        if not joined:
            cl_ptr.append(p); size = 0
        size += 1
        prev = cur
    cl_ptr.append(P)
    return np.array(cl_ptr)


def aggregate(Ecc_blocknorm, gsize):
    """greedy heavy-edge aggregation of the nodes of a symmetric weight matrix into groups of <= gsize"""
    n = Ecc_blocknorm.shape[0]
    A = Ecc_blocknorm.tocsr()
    agg = -np.ones(n, np.int64); na = 0
    deg = np.asarray(A.sum(axis=1)).ravel()
    for seed in np.argsort(-deg):
        if agg[seed] >= 0:
            continue
        members = [seed]; agg[seed] = na
        # grow by strongest connection to the current aggregate
        w = {}
        def add_nb(i):
            for k in range(A.indptr[i], A.indptr[i + 1]):
                j = A.indices[k]
                if agg[j] < 0:
                    w[j] = w.get(j, 0.0) + A.data[k]
        add_nb(seed)
        while len(members) < gsize and w:
            j = max(w, key=w.get)
            del w[j]
            if agg[j] >= 0:
                continue
            agg[j] = na; members.append(j); add_nb(j)
        na += 1
    return agg, na


class Level:
    pass


def pcg(apply_S, apply_M, b, tol=1e-12, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = apply_M(r); p = z.copy(); rz = r @ z; rz0 = rz
    for it in range(1, maxit + 1):
        y = apply_S(p); a = rz / (p @ y)
        x += a * p; r -= a * y; z = apply_M(r); rzn = r @ z
        if not (rzn > tol * tol * rz0):
            return x, it
        p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
    lam = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-5
    gs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    from arm_slam_calib_b200.sim import make_sim_problem, pose12
    n = int(round(900 * np.sqrt(scale)))
    pb = make_sim_problem(trajectory_size=int(round(10000 * scale)), nx=n, ny=n, size_x=1.3, size_y=1.3, z_range=(-0.9, 1.6),
                          max_range=0.27 * scale ** (-1.0 / 3.0), segment_length=10, seed=0, extrinsic_guess=pose12(t=[0.02, 0.02, 0.02]))
    print("P %d L %d M %d" % (pb.enc.shape[0], pb.lm_prior.shape[0], pb.pose_idx.shape[0]))
    Sqq, SqK, SKK, rhs, N, P = build_S(pb, lam)
    PN = P * N
    def apply_S(x):
        return np.concatenate([Sqq @ x[:PN] + SqK @ x[PN:], SqK.T @ x[:PN] + SKK @ x[PN:]])
    cl_ptr = clusters_of(pb, N, P)
    nC = len(cl_ptr) - 1
    print("clusters %d" % nC)
    # cluster-Jacobi
    Minv = []
    for c in range(nC):
        a, b = cl_ptr[c] * N, cl_ptr[c + 1] * N
        Minv.append(np.linalg.inv(Sqq[a:b, a:b].toarray()))
    MK = np.linalg.inv(SKK)
    def M_cl(r):
        z = np.empty_like(r)
        for c in range(nC):
            a, b = cl_ptr[c] * N, cl_ptr[c + 1] * N
            z[a:b] = Minv[c] @ r[a:b]
        z[PN:] = MK @ r[PN:]
        return z
    # coarse space Z1: uniform N-vector per cluster + K
    pose_cl = np.repeat(np.arange(nC), np.diff(cl_ptr))
    Zr = np.arange(PN); Zc = pose_cl[Zr // N] * N + Zr % N
    Z1 = sp.csr_matrix((np.ones(PN), (Zr, Zc)), shape=(PN, nC * N))
    Ecc = (Z1.T @ Sqq @ Z1).tocsr()
    EcK = Z1.T @ SqK
    n1 = nC * N
    E1 = sp.bmat([[Ecc, sp.csr_matrix(EcK)], [sp.csr_matrix(EcK.T), sp.csr_matrix(SKK)]]).tocsr()
    bn = sp.bsr_matrix(Ecc, blocksize=(N, N))
    print("E1 dim %d, block density %.4f (%d blocks of %d)" % (n1 + 6, bn.data.shape[0] / nC ** 2, bn.data.shape[0], nC * nC))
    def restrict1(r):
        return np.concatenate([Z1.T @ r[:PN], r[PN:]])
    def prolong1(zc):
        return np.concatenate([Z1 @ zc[:n1], zc[n1:]])
    t0 = time.time(); x, it = pcg(apply_S, M_cl, rhs); print("cl       : %4d iterations (%.1f s)" % (it, time.time() - t0))
    E1d = E1.toarray() if n1 < 8000 else None
    if E1d is not None:
        cf = sla.cho_factor(E1d)
        def M_2lv(r):
            return M_cl(r) + prolong1(sla.cho_solve(cf, restrict1(r)))
        t0 = time.time(); x, it = pcg(apply_S, M_2lv, rhs); print("2lv exact: %4d iterations (%.1f s)" % (it, time.time() - t0))
    import scipy.sparse.linalg as spla
    def build_levels(gs, top_dim):
        levels = []
        Ecur = Ecc; EKcur = EcK; ncur = nC
        while True:
            lv = Level(); lv.E = Ecur; lv.EK = EKcur; lv.n = ncur
            levels.append(lv)
            if ncur * N <= top_dim:
                break
            bs = sp.bsr_matrix(Ecur, blocksize=(N, N))
            wts = sp.csr_matrix((np.sqrt((bs.data ** 2).sum(axis=(1, 2))), bs.indices, bs.indptr), shape=(ncur, ncur))
            wts.setdiag(0); wts.eliminate_zeros()
            agg, na = aggregate(wts, gs)
            members = [np.where(agg == a)[0] for a in range(na)]
            idx = [np.concatenate([np.arange(m * N, m * N + N) for m in mem]) for mem in members]
            blocks = []; rr = []; cc = []
            for ix in idx:
                Di = np.linalg.inv(Ecur[ix][:, ix].toarray())
                R, Cc = np.meshgrid(ix, ix, indexing="ij")
                rr.append(R.ravel()); cc.append(Cc.ravel()); blocks.append(Di.ravel())
            lv.Dinv = sp.csr_matrix((np.concatenate(blocks), (np.concatenate(rr), np.concatenate(cc))), shape=(ncur * N, ncur * N))
            r_ = np.arange(ncur * N); c_ = agg[r_ // N] * N + r_ % N
            lv.Z = sp.csr_matrix((np.ones(ncur * N), (r_, c_)), shape=(ncur * N, na * N))
            Ecur = (lv.Z.T @ Ecur @ lv.Z).tocsr(); EKcur = lv.Z.T @ EKcur; ncur = na
            print("  level: %d nodes -> %d aggregates (max block %d, E nnz-blocks/row %.1f)" % (lv.n, na, max(len(i) for i in idx), sp.bsr_matrix(lv.E, blocksize=(N, N)).data.shape[0] / lv.n))
        top = levels[-1]
        Etop = np.block([[top.E.toarray(), top.EK], [top.EK.T, SKK]])
        levels_ctop = sla.cho_factor(Etop)
        print("  top level dim %d" % (top.n * N + 6))
        return levels, levels_ctop

    def run(name, M):
        t0 = time.time(); x, it = pcg(apply_S, M, rhs, maxit=3000); print("%-28s: %4d iterations (%.1f s)" % (name, it, time.time() - t0)); sys.stdout.flush()

    for gs, top_dim in ((8, 1600), (4, 3200), (16, 1600)):
        print("gs %d top_dim %d" % (gs, top_dim))
        levels, ctop = build_levels(gs, top_dim)
        ntop = levels[-1].n * N

        def V_add(li, r, rK):
            lv = levels[li]
            if li == len(levels) - 1:
                s = sla.cho_solve(ctop, np.concatenate([r, rK]))
                return s[:ntop], s[ntop:]
            zc, zK = V_add(li + 1, lv.Z.T @ r, rK)
            return lv.Dinv @ r + lv.Z @ zc, zK

        def V_mul(li, r, rK, om, nu):
            lv = levels[li]
            if li == len(levels) - 1:
                s = sla.cho_solve(ctop, np.concatenate([r, rK]))
                return s[:ntop], s[ntop:]
            z = om * (lv.Dinv @ r)
            for _ in range(nu - 1):
                z = z + om * (lv.Dinv @ (r - lv.E @ z))
            r2 = r - lv.E @ z
            rK2 = rK - lv.EK.T @ z
            zc, zK = V_mul(li + 1, lv.Z.T @ r2, rK2, om, nu)
            z = z + lv.Z @ zc
            for _ in range(nu):
                z = z + om * (lv.Dinv @ (r - lv.E @ z - lv.EK @ zK))
            return z, zK

        def make_M(V):
            def M(r):
                z = M_cl(r)
                rc = restrict1(r)
                zc, zK = V(rc[:n1], rc[n1:])
                return z + prolong1(np.concatenate([zc, zK]))
            return M
        run("ml-add", make_M(lambda r, rK: V_add(0, r, rK)))
        for om, nu in ((1.0, 1), (0.8, 1), (0.8, 2), (0.8, 3)):
            run("ml-v w=%.1f nu=%d" % (om, nu), make_M(lambda r, rK: V_mul(0, r, rK, om, nu)))
        # Chebyshev / fixed-step Richardson wrapped around the V-cycle: z_{k+1} = z_k + V(r - E1 z_k), k steps (symmetric if V symmetric)
        def wrapped(k, om, nu):
            lv0 = levels[0]
            def V(r, rK):
                z, zK = V_mul(0, r, rK, om, nu)
                for _ in range(k - 1):
                    rr = r - lv0.E @ z - lv0.EK @ zK
                    rrK = rK - lv0.EK.T @ z - SKK @ zK
                    dz, dzK = V_mul(0, rr, rrK, om, nu)
                    z = z + dz; zK = zK + dzK
                return z, zK
            return V
        for k in (2, 3):
            run("ml-v w=0.8 nu=1 x%d cycles" % k, make_M(wrapped(k, 0.8, 1)))


if __name__ == "__main__":
    main()
