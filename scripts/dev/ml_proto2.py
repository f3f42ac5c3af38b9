"""Prototype 2: spectral (lowest-eigenvector) coarse bases instead of piecewise-constant ones.  usage: ml_proto2.py <scale>"""
import sys, time, os
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ml_proto import build_S, clusters_of, aggregate, pcg


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
    lam = 1e-5
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
        t0 = time.time(); x, it = pcg(apply_S, M, rhs, maxit=3000); print("%-34s: %4d iterations (%.1f s)" % (name, it, time.time() - t0)); sys.stdout.flush()
    eig = [np.linalg.eigh(b) for b in blocks]

    def two_level(Z, name):
        nz = Z.shape[1]
        Ecc = (Z.T @ Sqq @ Z).toarray(); EcK = Z.T @ SqK
        E = np.block([[Ecc, EcK], [EcK.T, SKK]])
        cf = sla.cho_factor(E)
        def M(r):
            rc = np.concatenate([Z.T @ r[:PN], r[PN:]])
            zc = sla.cho_solve(cf, rc)
            return M_cl(r) + np.concatenate([Z @ zc[:nz], zc[nz:]])
        run("%s (coarse dim %d)" % (name, nz + 6), M)

    # uniform basis (round 1)
    pose_cl = np.repeat(np.arange(nC), np.diff(cl_ptr))
    r_ = np.arange(PN); c_ = pose_cl[r_ // N] * N + r_ % N
    two_level(sp.csr_matrix((np.ones(PN), (r_, c_)), shape=(PN, nC * N)), "uniform")
    for k in (2, 3, 4, 6, 9, 12):
        rr, cc, vv = [], [], []
        col = 0
        for c in range(nC):
            a = cl_ptr[c] * N; nb = blocks[c].shape[0]; kk = min(k, nb)
            V = eig[c][1][:, :kk]
            R, Cc = np.meshgrid(np.arange(a, a + nb), np.arange(col, col + kk), indexing="ij")
            rr.append(R.ravel()); cc.append(Cc.ravel()); vv.append(V.ravel()); col += kk
        Z = sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(PN, col))
        two_level(Z, "spectral k=%d" % k)
    # uniform + spectral complement
    for k in (3, 6):
        rr, cc, vv = [], [], []
        col = 0
        for c in range(nC):
            a = cl_ptr[c] * N; nb = blocks[c].shape[0]
            U = np.tile(np.eye(N), (nb // N, 1)) / np.sqrt(nb // N)
            V = eig[c][1][:, :min(k, nb)]
            V = V - U @ (U.T @ V)
            B = np.linalg.qr(np.concatenate([U, V], axis=1))[0]
            kk = B.shape[1]
            R, Cc = np.meshgrid(np.arange(a, a + nb), np.arange(col, col + kk), indexing="ij")
            rr.append(R.ravel()); cc.append(Cc.ravel()); vv.append(B.ravel()); col += kk
        Z = sp.csr_matrix((np.concatenate(vv), (np.concatenate(rr), np.concatenate(cc))), shape=(PN, col))
        two_level(Z, "uniform + spectral k=%d" % k)


if __name__ == "__main__":
    main()
