// Multilevel coarse solve of the two-level PCG preconditioner (replaces the dense cuSOLVER/cuBLAS coarse level of round 1).
//
// Coarse space of level 1 (unchanged): one uniform joint-offset vector per cluster (N columns) + the extrinsic (6).
//   E = Z^T S Z = [ E_cc  E_cK ; E_Kc  S_KK ]
// E_cc is BLOCK-SPARSE: block (a, b) is non-zero only when clusters a and b share a landmark (C2: 2.7 % of the
// blocks, ~27 per block row, and the per-row count does not grow with the problem).  It is stored as BSR with N x N fp64
// blocks and never densified.  E_cc^-1 is approximated by one symmetric V-cycle V:
//   level l:   Chebyshev(m) smoother on D_l^-1 E_l (D_l = the N x N diagonal blocks), coarse space = one N-vector per
//              AGGREGATE of ~8 strongly connected nodes (greedy heavy-edge aggregation on the shared-landmark counts, built on
//              the host), E_{l+1} = P^T E_l P = block sums (no triple product: P is piecewise constant);
//   top level: dense (<= pcg_ml_top_dim unknowns), explicit inverse by an in-place Gauss-Jordan sweep in a thread-block cluster.
// The dense extrinsic border is eliminated exactly around V by bordering:
//   Y = V E_cK (six V-cycles per refresh),  S_k = S_KK - E_Kc Y,
//   B [r; r_K]:  t = V r,  z_K = S_k^-1 (r_K - E_Kc t),  z = t - Y z_K          ( = inverse of [V^-1 E_cK; E_Kc S_KK], SPD because V <= E_cc^-1 )
// Everything is a fixed linear SPD operator, so PCG stays exact; only the iteration count depends on it (measured on C2, first
// solve: exact dense coarse solve 90 iterations, this V-cycle with m = 4: ~110; scripts/dev/ml_proto6.py is the numpy prototype).
// Reductions use fixed trees / fixed member orders: results are run-to-run identical and identical on every rank.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace asck {

// Level description handed to the kernels (device pointers; handle-lifetime, so the captured PCG graph can keep them)
struct MlLevelDev {
    int n;                    // nodes (block rows)
    const int* row_ptr;       // [n + 1]
    const int* col;           // [nnzb]
    const double* Eb;         // [nnzb][N*N] row-major blocks
    const double* Dinv;       // [n][N*N] inverses of the diagonal blocks
    const double* lmax;       // [1] power-iteration estimate of lambda_max(D^-1 E)
};

constexpr double kMlSafety = 1.2;   // upper end of the Chebyshev interval = kMlSafety * estimate (the power iteration converges from below)

// Row i of E times v: ONE BLOCK PER LANE (a block row holds ~27 blocks), so the dependent loads are row_ptr -> col -> v once per
// row instead of once per block (the phases of the V-cycle are latency-bound: < 10 MB of L2-resident data, ~20 dependent phases
// per PCG iteration).  The N partial results per lane are summed over the warp with a fixed butterfly; lane a < N returns entry a.
// v_eff = v (+ vc[agg] when vc != null: prolongation on the fly).  Vectors are read through L2 (__ldcg).
template <int N>
__device__ __forceinline__ double bsr_row_dot(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ Eb,
                                              const double* v, const double* vc, const int* __restrict__ agg, int i, int lane) {
    const int k0 = __ldg(row_ptr + i), k1 = __ldg(row_ptr + i + 1);
    double acc[N];
#pragma unroll
    for (int a = 0; a < N; ++a) acc[a] = 0.0;
    for (int k = k0 + lane; k < k1; k += 32) {
        const int c = __ldg(col + k);
        const double* blk = Eb + (size_t)k * N * N;
        double x[N];
#pragma unroll
        for (int b = 0; b < N; ++b) x[b] = __ldcg(v + (size_t)c * N + b);
        if (vc) {
            const double* cc = vc + (size_t)__ldg(agg + c) * N;
#pragma unroll
            for (int b = 0; b < N; ++b) x[b] += __ldcg(cc + b);
        }
#pragma unroll
        for (int a = 0; a < N; ++a) {
            double t = 0.0;
#pragma unroll
            for (int b = 0; b < N; ++b) t = fma(__ldg(blk + a * N + b), x[b], t);
            acc[a] += t;
        }
    }
    double mine = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a) {
        const double t = warp_sum(acc[a]);
        if (lane == a) mine = t;
    }
    return mine;
}

// y_a = sum_b D[a][b] x_b over the first N lanes
template <int N>
__device__ __forceinline__ double small_matvec(const double* __restrict__ D, double x, int lane) {
    double y = 0.0;
#pragma unroll
    for (int b = 0; b < N; ++b) {
        const double xb = __shfl_sync(kFull, x, b);
        if (lane < N) y = fma(D[lane * N + b], xb, y);
    }
    return y;
}

// One Chebyshev step of the smoother on level `lv` for node i (one warp):
//   res = r - E (z_in [+ P zc]);  t = D^-1 res;  d <- c1 d + c2 t;  z_out = z_in [+ P zc] + d
// step 0 of a sweep: d is not read (c1 = 0, c2 = 1/theta); zero_in: z_in == 0 (pre-smoothing), the product is skipped.
// Vectors are read through L2 (__ldcg): inside the persistent V-cycle kernel other CTAs rewrote them a phase ago.
template <int N>
__device__ __forceinline__ void ml_cheb_node(const MlLevelDev& lv, double frac, int step, int zero_in, const double* r, const double* z_in,
                                             double* z_out, double* d, const double* zc, const int* agg, int i, int lane) {
    const double hi = kMlSafety * *lv.lmax, lo = hi / frac;
    const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
    double c1 = 0.0, c2 = 1.0 / theta;
    {
        double rho = 1.0 / sigma;
        for (int k = 1; k <= step; ++k) { const double rn = 1.0 / (2.0 * sigma - rho); c1 = rn * rho; c2 = 2.0 * rn / delta; rho = rn; }
    }
    double res = lane < N ? __ldcg(r + (size_t)i * N + lane) : 0.0;
    double zin = 0.0;
    if (!zero_in) {
        const double p = bsr_row_dot<N>(lv.row_ptr, lv.col, lv.Eb, z_in, zc, agg, i, lane);
        if (lane < N) {
            res -= p;
            zin = __ldcg(z_in + (size_t)i * N + lane);
            if (zc) zin += __ldcg(zc + (size_t)agg[i] * N + lane);
        }
    }
    const double t = small_matvec<N>(lv.Dinv + (size_t)i * N * N, res, lane);
    if (lane < N) {
        const double dn = (step > 0 ? c1 * __ldcg(d + (size_t)i * N + lane) : 0.0) + c2 * t;
        d[(size_t)i * N + lane] = dn;
        z_out[(size_t)i * N + lane] = zin + dn;
    }
}

// r_c[A] = sum over the members i of aggregate A of (r - E z)[i].  The rows of one aggregate are summed anyway, so the warp walks
// the flattened list of ALL blocks of the aggregate's rows (rptr / ridx, built on the host), one block per lane and round:
// the loads of different rounds are independent, and one butterfly at the end finishes the sum.
template <int N>
__device__ __forceinline__ void ml_restrict_node(const MlLevelDev& lv, const int* __restrict__ mptr, const int* __restrict__ midx,
                                                 const int* __restrict__ rptr, const int* __restrict__ ridx, const double* r, const double* z,
                                                 double* rc, int A, int lane) {
    double acc[N];
#pragma unroll
    for (int a = 0; a < N; ++a) acc[a] = 0.0;
    const int q0 = __ldg(rptr + A), q1 = __ldg(rptr + A + 1);
    for (int q = q0 + lane; q < q1; q += 32) {
        const int k = __ldg(ridx + q);
        const int c = __ldg(lv.col + k);
        const double* blk = lv.Eb + (size_t)k * N * N;
        double x[N];
#pragma unroll
        for (int b = 0; b < N; ++b) x[b] = __ldcg(z + (size_t)c * N + b);
#pragma unroll
        for (int a = 0; a < N; ++a) {
            double t = 0.0;
#pragma unroll
            for (int b = 0; b < N; ++b) t = fma(__ldg(blk + a * N + b), x[b], t);
            acc[a] -= t;
        }
    }
    const int m0 = __ldg(mptr + A), m1 = __ldg(mptr + A + 1);
    for (int q = m0 + lane; q < m1; q += 32) {
        const int i = __ldg(midx + q);
#pragma unroll
        for (int a = 0; a < N; ++a) acc[a] += __ldcg(r + (size_t)i * N + a);
    }
#pragma unroll
    for (int a = 0; a < N; ++a) {
        const double t = warp_sum(acc[a]);
        if (lane == a) rc[(size_t)A * N + a] = t;
    }
}

// row `row` of y = A x, dense row-major n x n (the explicit inverse of the top level); one warp
__device__ __forceinline__ void ml_gemv_row(const double* __restrict__ A, const double* x, double* y, int n, int row, int lane) {
    const double* a = A + (size_t)row * n;
    double acc0 = 0.0, acc1 = 0.0;
    int i = lane;
    for (; i + 32 < n; i += 64) { acc0 = fma(a[i], __ldcg(x + i), acc0); acc1 = fma(a[i + 32], __ldcg(x + i + 32), acc1); }
    if (i < n) acc0 = fma(a[i], __ldcg(x + i), acc0);
    const double t = warp_sum(acc0 + acc1);
    if (lane == 0) y[row] = t;
}

template <int N>
__global__ void __launch_bounds__(128)
ml_cheb_step_kernel(MlLevelDev lv, double frac, int step, int zero_in, const double* r, const double* z_in, double* z_out, double* d,
                    const double* zc, const int* agg, const int* __restrict__ done) {
    if (done && *done) return;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= lv.n) return;
    ml_cheb_node<N>(lv, frac, step, zero_in, r, z_in, z_out, d, zc, agg, i, lane);
}

template <int N>
__global__ void __launch_bounds__(128)
ml_residual_restrict_kernel(MlLevelDev lv, int n_coarse, const int* __restrict__ mptr, const int* __restrict__ midx, const int* __restrict__ rptr,
                            const int* __restrict__ ridx, const double* r, const double* z, double* rc, const int* __restrict__ done) {
    if (done && *done) return;
    const int A = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (A >= n_coarse) return;
    ml_restrict_node<N>(lv, mptr, midx, rptr, ridx, r, z, rc, A, lane);
}

__global__ void __launch_bounds__(128) ml_dense_gemv_kernel(const double* __restrict__ A, const double* x, double* y, int n, const int* __restrict__ done) {
    if (done && *done) return;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    ml_gemv_row(A, x, y, n, row, lane);
}

// ---- the whole V-cycle as ONE persistent kernel: the phases above separated by a grid barrier instead of kernel boundaries.
// On C2 the V-cycle is ~20 dependent phases over < 10 MB of L2-resident data: as separate launches it costs ~170 us per PCG
// iteration (launch gaps), as one kernel ~40 us.  kCycleCtas CTAs must be co-resident (they are: <= one per SM, 256 threads, no
// shared memory); a barrier that does not complete within ~2 s reports through `fail` instead of hanging.
constexpr int kCycleMaxLevels = 8;
constexpr int kCycleThreads = 256;
struct MlCycleArgs {
    int nl, m, ntop, first_cluster_level, dbg;
    double frac;
    MlLevelDev lv[kCycleMaxLevels];
    const int* agg[kCycleMaxLevels];
    const int* mptr[kCycleMaxLevels];
    const int* midx[kCycleMaxLevels];
    const int* rptr[kCycleMaxLevels];
    const int* ridx[kCycleMaxLevels];
    int n_next[kCycleMaxLevels];
    const double* r[kCycleMaxLevels + 1];   // r[0] = the caller's right-hand side; r[l] = restricted residual of level l
    double* rw[kCycleMaxLevels + 1];        // the same buffers, writable (rw[0] unused)
    double* za[kCycleMaxLevels];
    double* zb[kCycleMaxLevels];
    double* d[kCycleMaxLevels];
    const double* Atop;
    double* ztop;
    unsigned* bar;                          // [2]: arrival counter, exit counter (both return to zero)
    int* fail;
    const int* done;
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void ml_grid_barrier(unsigned* bar, unsigned& target, int* fail, int variant = 0) {
    __syncthreads();   // the CTA's writes of this phase happen-before thread 0's release below (bar.sync + cumulativity)
    if (threadIdx.x == 0) {
        target += gridDim.x;
        if (variant == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(bar) < target) {
                if (clock64() - t0 > 4000000000ll) { *fail = 1; break; }
            }
        } else {   // one fence on each side, plain polling in between (an acquire load invalidates L1 on EVERY poll: CCTL.IVALL)
            __threadfence();
            atomicAdd(bar, 1u);
            const long long t0 = clock64();
            while (*((volatile unsigned*)bar) < target) {
                if (clock64() - t0 > 4000000000ll) { *fail = 1; break; }
            }
            __threadfence();
        }
    }
    __syncthreads();
}

// Phases of the FINE levels span the whole grid and are separated by the grid barrier (~8 us per phase on a B200: three dependent L2
// round trips plus the barrier).  Levels with at most kClusterLevelNodes nodes (and everything below them, top level included) are
// handled by thread-block cluster 0 alone, with the hardware cluster barrier between phases (~2.5 us per phase); the other
// clusters wait at the next grid barrier.  a.first_cluster_level = first such level (nl + 1: none).
constexpr int kCycleCluster = 8;
constexpr int kClusterLevelNodes = 640;

template <int N>
__global__ void __launch_bounds__(kCycleThreads) ml_vcycle_kernel(const __grid_constant__ MlCycleArgs a) {
    if (a.done && *a.done) return;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, wpc = kCycleThreads / 32;
    const int gw = blockIdx.x * wpc + (threadIdx.x >> 5), ngw = gridDim.x * wpc;
    const bool c0 = blockIdx.x < kCycleCluster;                                   // member of cluster 0
    const int cw = (blockIdx.x % kCycleCluster) * wpc + (threadIdx.x >> 5), ncw = kCycleCluster * wpc;
    const int fc = a.first_cluster_level;
    unsigned target = 0;
    const double* cur[kCycleMaxLevels + 1];
    // which output buffer every level ends its sweeps in is static: every CTA tracks it, whether or not it does the work
    for (int l = 0; l < a.nl; ++l) {   // down
        const bool cl = l >= fc;
        const double* r = a.r[l];
        double* bufs[2] = {a.za[l], a.zb[l]};
        const int w0 = cl ? cw : gw, nw = cl ? ncw : ngw;
        for (int k = 0; k < a.m; ++k) {
            if (k > 0 || l > 0) {
                if (!cl || l == fc && k == 0) ml_grid_barrier(a.bar, target, a.fail, a.dbg >> 1);   // (entering cluster mode: the restriction above was a grid phase)
                else if (c0) cluster.sync();
            }
            if ((!cl || c0) && !(a.dbg & 1))
                for (int i = w0; i < a.lv[l].n; i += nw)
                    ml_cheb_node<N>(a.lv[l], a.frac, k, k == 0 ? 1 : 0, r, bufs[(k + 1) & 1], bufs[k & 1], a.d[l], nullptr, nullptr, i, lane);
        }
        cur[l] = bufs[(a.m - 1) & 1];
        if (!cl) ml_grid_barrier(a.bar, target, a.fail, a.dbg >> 1);
        else if (c0) cluster.sync();
        if ((!cl || c0) && !(a.dbg & 1))
            for (int A = w0; A < a.n_next[l]; A += nw) ml_restrict_node<N>(a.lv[l], a.mptr[l], a.midx[l], a.rptr[l], a.ridx[l], r, cur[l], a.rw[l + 1], A, lane);
    }
    {   // top level
        const bool cl = a.nl >= fc;
        if (a.nl > 0) {
            if (!cl || a.nl == fc) ml_grid_barrier(a.bar, target, a.fail, a.dbg >> 1);
            else if (c0) cluster.sync();
        }
        const int w0 = cl ? cw : gw, nw = cl ? ncw : ngw;
        if ((!cl || c0) && !(a.dbg & 1))
            for (int row = w0; row < a.ntop; row += nw) ml_gemv_row(a.Atop, a.r[a.nl], a.ztop, a.ntop, row, lane);
        cur[a.nl] = a.ztop;
    }
    for (int l = a.nl - 1; l >= 0; --l) {   // up
        const bool cl = l >= fc;
        const double* r = a.r[l];
        const double* zi = cur[l];
        const int w0 = cl ? cw : gw, nw = cl ? ncw : ngw;
        for (int k = 0; k < a.m; ++k) {
            if (!cl) ml_grid_barrier(a.bar, target, a.fail, a.dbg >> 1);   // also the hand-over from cluster 0's coarse section
            else if (c0) cluster.sync();
            double* zo = zi == a.za[l] ? a.zb[l] : a.za[l];
            if ((!cl || c0) && !(a.dbg & 1))
                for (int i = w0; i < a.lv[l].n; i += nw)
                    ml_cheb_node<N>(a.lv[l], a.frac, k, 0, r, zi, zo, a.d[l], k == 0 ? cur[l + 1] : nullptr, k == 0 ? a.agg[l] : nullptr, i, lane);
            zi = zo;
        }
        cur[l] = zi;
    }
    // the last CTA to leave resets the counters for the next launch (everybody has passed every barrier by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.bar + 1, 1u) == gridDim.x - 1) { a.bar[0] = 0; a.bar[1] = 0; __threadfence(); }
    }
}

// ---- refresh kernels (once per damped solve)
// level-1 blocks from the incidence products (see coarse_blocks_kernel in kernels.cuh for the pair lists):
//   E[ca][cb] = - sum over the destination's (incidence a, incidence b) pairs of T_a V_b^T, written at its BSR position and,
//   transposed, at the mirrored one.  One warp per destination block; destinations without pairs on this rank write zeros.
template <int N>
__global__ void __launch_bounds__(128)
ml_blocks_l1_kernel(int nDest, const int* __restrict__ dest_ptr, const int* __restrict__ pos_lo, const int* __restrict__ pos_up,
                    const int* __restrict__ pair_a, const int* __restrict__ pair_b, const double* __restrict__ V, const double* __restrict__ T,
                    double* __restrict__ Eb) {
    constexpr int NE = (N * N + 31) / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x * 4 + warp;
    if (d >= nDest) return;
    double acc[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) acc[i] = 0.0;
    for (int pr = dest_ptr[d]; pr < dest_ptr[d + 1]; ++pr) {
        const double* Ta = T + (size_t)__ldg(pair_a + pr) * 3 * N;
        const double* Vb = V + (size_t)__ldg(pair_b + pr) * 3 * N;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            const int e = lane + 32 * i;
            if (e < N * N) {
                const int a = e / N, b = e % N;
                acc[i] += Ta[3 * a] * Vb[3 * b] + Ta[3 * a + 1] * Vb[3 * b + 1] + Ta[3 * a + 2] * Vb[3 * b + 2];
            }
        }
    }
    const int plo = pos_lo[d], pup = pos_up[d];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        const int e = lane + 32 * i;
        if (e < N * N) {
            Eb[(size_t)plo * N * N + e] = -acc[i];
            if (pup != plo) Eb[(size_t)pup * N * N + (e % N) * N + e / N] = -acc[i];
        }
    }
}

// diagonal blocks += sum_{p in cluster} (B_pp + lambda I)   (rank 0 only in a multi-GPU run); thread per (c, a, b)
__global__ void ml_diag_l1_kernel(int nC, int N, double lambda, const int* __restrict__ cl_ptr, const double* __restrict__ Bpp,
                                  const int* __restrict__ diag_pos, double* __restrict__ Eb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nC * N * N) return;
    const int c = i / (N * N), a = (i / N) % N, b = i % N;
    double s = 0.0;
    for (int p = cl_ptr[c]; p < cl_ptr[c + 1]; ++p) s += Bpp[(size_t)p * N * N + a * N + b];
    if (a == b) s += lambda * (cl_ptr[c + 1] - cl_ptr[c]);
    Eb[(size_t)diag_pos[c] * N * N + a * N + b] += s;
}

// extrinsic border of level 1: EK[(c N + a) 6 + r] = sum_{p in c} B_pK[a][r] - sum over the cluster's incidences of (T W_Kj^T)[a][r]
__global__ void ml_border_l1_kernel(int nC, int N, int add_diag, const int* __restrict__ cl_ptr, const int* __restrict__ cl_inc_ptr,
                                    const int* __restrict__ cl_inc, const int* __restrict__ inc_lm, const double* __restrict__ T,
                                    const double* __restrict__ WK, const double* __restrict__ BpK, double* __restrict__ EK) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nC * N * 6) return;
    const int c = i / (N * 6), a = (i / 6) % N, r = i % 6;
    double s = 0.0;
    if (add_diag)
        for (int p = cl_ptr[c]; p < cl_ptr[c + 1]; ++p) s += BpK[(size_t)p * N * 6 + a * 6 + r];
    for (int k = cl_inc_ptr[c]; k < cl_inc_ptr[c + 1]; ++k) {
        const int inc = cl_inc[k];
        const double* t = T + (size_t)inc * 3 * N + 3 * a;
        const double* w = WK + 18 * (size_t)inc_lm[inc] + 3 * r;
        s -= t[0] * w[0] + t[1] * w[1] + t[2] * w[2];
    }
    EK[(size_t)(c * N + a) * 6 + r] = s;
}

// Galerkin coarsening with a piecewise-constant prolongator = block sums: coarse block k = sum of its child blocks (list order)
__global__ void ml_coarsen_kernel(int nnzb_c, int NN, const int* __restrict__ cptr, const int* __restrict__ cidx, const double* __restrict__ Ef,
                                  double* __restrict__ Ec) {
    const int k = blockIdx.x, e = threadIdx.x;
    if (k >= nnzb_c || e >= NN) return;
    double s = 0.0;
    for (int q = cptr[k]; q < cptr[k + 1]; ++q) s += Ef[(size_t)cidx[q] * NN + e];
    Ec[(size_t)k * NN + e] = s;
}

// Dinv[i] = inverse of the N x N diagonal block (SPD): Gauss-Jordan without pivoting, one thread per node
template <int N>
__global__ void __launch_bounds__(64) ml_dinv_kernel(int n, const int* __restrict__ diag_pos, const double* __restrict__ Eb, double* __restrict__ Dinv,
                                                     int* __restrict__ fail) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double A[N][N];
    const double* src = Eb + (size_t)diag_pos[i] * N * N;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) A[a][b] = src[a * N + b];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double piv = A[k][k];
        ok = ok && (piv > 0.0);
        const double ip = 1.0 / piv;
        double colk[N];
#pragma unroll
        for (int a = 0; a < N; ++a) colk[a] = A[a][k];
#pragma unroll
        for (int b = 0; b < N; ++b) A[k][b] = b == k ? ip : A[k][b] * ip;
#pragma unroll
        for (int a = 0; a < N; ++a)
            if (a != k)
#pragma unroll
                for (int b = 0; b < N; ++b) A[a][b] = (b == k ? 0.0 : A[a][b]) - colk[a] * A[k][b];
    }
    double* dst = Dinv + (size_t)i * N * N;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) dst[a * N + b] = 0.5 * (A[a][b] + A[b][a]);   // exactly symmetric
    if (!ok) atomicExch(fail, 1);
}

// power iteration on D^-1 E (no normalisation inside: 16 steps grow the vector by ~2^16): w = D^-1 E v; with `part`, also the
// CTA partial of |w|^2 (for the estimate |w_k| / |w_{k-1}| of the last two steps)
template <int N>
__global__ void __launch_bounds__(128)
ml_power_step_kernel(int n, const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ Eb, const double* __restrict__ Dinv,
                     const double* __restrict__ v, double* __restrict__ w, double* __restrict__ part) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + warp;
    __shared__ double s[4];
    double n2 = 0.0;
    if (i < n) {
        const double p = bsr_row_dot<N>(row_ptr, col, Eb, v, nullptr, nullptr, i, lane);
        const double t = small_matvec<N>(Dinv + (size_t)i * N * N, lane < N ? p : 0.0, lane);
        if (lane < N) { w[(size_t)i * N + lane] = t; n2 = t * t; }
    }
    n2 = warp_sum(n2);
    if (lane == 0) s[warp] = n2;
    __syncthreads();
    if (part && threadIdx.x == 0) part[blockIdx.x] = (s[0] + s[1]) + (s[2] + s[3]);
}

// deterministic start vector of the power iteration: a fixed hash of the index in [0.5, 1.5)
__global__ void ml_power_init_kernel(int n, double* __restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned x = (unsigned)i * 2654435761u + 12345u;
    x ^= x >> 16; x *= 2246822519u; x ^= x >> 13;
    v[i] = 0.5 + (double)(x & 0xffffff) / 16777216.0;
}

// lmax = sqrt(sum partB / sum partA); one block
__global__ void __launch_bounds__(256) ml_power_finish_kernel(const double* __restrict__ partA, const double* __restrict__ partB, int nPart,
                                                               double* __restrict__ lmax) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nPart; i += 256) { a += partA[i]; b += partB[i]; }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *lmax = (sa[0] > 0.0 && sb[0] > 0.0) ? sqrt(sb[0] / sa[0]) : 1.0;
}

// dense copy of the last level: A[n N][n N] row-major from its BSR blocks
__global__ void ml_top_assemble_kernel(int n, int N, const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ Eb,
                                       double* __restrict__ A) {
    const int i = blockIdx.x, e = threadIdx.x;
    if (i >= n || e >= N * N) return;
    const int ld = n * N, a = e / N, b = e % N;
    for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) A[(size_t)(i * N + a) * ld + col[k] * N + b] = Eb[(size_t)k * N * N + e];
}

// In-place inverse of the dense SPD top-level matrix (n <= ~1024) by a Gauss-Jordan sweep without pivoting; one thread-block
// CLUSTER of kTopCtas CTAs shares the work of every pivot step (rows are dealt round-robin to the cluster's threads' CTAs) and
// synchronises with the hardware cluster barrier twice per step.  The matrix stays in L2 (n = 600: 2.9 MB).
constexpr int kTopCtas = 8;
constexpr int kTopThreads = 512;
__global__ void __launch_bounds__(kTopThreads) ml_top_invert_kernel(double* A, int n, int* __restrict__ fail) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double colk[];   // [n] pivot column, per CTA
    const int cta = (int)cluster.block_rank(), ncta = (int)cluster.num_blocks();
    const int tid = threadIdx.x;
    bool bad = false;
    // rows are written by one CTA and read by the others in later steps: every access goes to L2 (.cg), never through L1
    for (int k = 0; k < n; ++k) {
        const double piv = __ldcg(A + (size_t)k * n + k);
        bad = bad || !(piv > 0.0);
        const double ip = 1.0 / piv;
        for (int i = tid; i < n; i += kTopThreads) colk[i] = __ldcg(A + (size_t)i * n + k);
        cluster.sync();   // everybody holds column k and the pivot before row k changes
        if (cta == 0)
            for (int c = tid; c < n; c += kTopThreads) __stcg(A + (size_t)k * n + c, c == k ? ip : __ldcg(A + (size_t)k * n + c) * ip);
        cluster.sync();
        // rows r != k, dealt to the cluster's warps round-robin; a warp walks one row (coalesced)
        const int warp = tid >> 5, lane = tid & 31, nw = kTopThreads / 32;
        for (int r = cta * nw + warp; r < n; r += ncta * nw) {
            if (r == k) continue;
            const double f = colk[r];
            double* row = A + (size_t)r * n;
            const double* rk = A + (size_t)k * n;
            for (int c = lane; c < n; c += 32) __stcg(row + c, (c == k ? 0.0 : __ldcg(row + c)) - f * __ldcg(rk + c));
        }
        cluster.sync();
    }
    if (bad && tid == 0 && cta == 0) atomicExch(fail, 1);
}

// The Gauss-Jordan sweep leaves an inverse that is symmetric only up to cond * eps; CG needs an exactly symmetric preconditioner
// (an asymmetry of 1e-7 stalls it at a 1e-7 relative residual): A <- (A + A^T) / 2, one thread per upper-triangle pair
__global__ void __launch_bounds__(256) ml_symmetrize_kernel(double* __restrict__ A, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int r = (int)(idx / n), c = (int)(idx % n);
    if (c <= r) return;
    const double v = 0.5 * (A[(size_t)r * n + c] + A[(size_t)c * n + r]);
    A[(size_t)r * n + c] = v; A[(size_t)c * n + r] = v;
}

// column c of the extrinsic border as a right-hand side: r[i] = EK[i * 6 + c]
__global__ void ml_border_column_kernel(int n, int c, const double* __restrict__ EK, double* __restrict__ r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = EK[(size_t)i * 6 + c];
}
// Y[:, c] = t
__global__ void ml_store_column_kernel(int n, int c, const double* __restrict__ t, double* __restrict__ Y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Y[(size_t)i * 6 + c] = t[i];
}

// S_k = S_KK - E_Kc Y (6 x 6, symmetrised) and its inverse; one block of 256 threads
__global__ void __launch_bounds__(256) ml_border_schur_kernel(int n, const double* __restrict__ EK, const double* __restrict__ Y, const double* __restrict__ SKK,
                                                               double* __restrict__ Ski, int* __restrict__ fail) {
    __shared__ double s[8][36];
    double acc[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) acc[e] = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        double ek[6], y[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { ek[c] = EK[(size_t)i * 6 + c]; y[c] = Y[(size_t)i * 6 + c]; }
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = 0; b < 6; ++b) acc[6 * a + b] = fma(ek[a], y[b], acc[6 * a + b]);
    }
#pragma unroll
    for (int e = 0; e < 36; ++e) {
        const double t = warp_sum(acc[e]);
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][e] = t;
    }
    __syncthreads();
    if (threadIdx.x < 36) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
        s[0][threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double A[6][6];
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) A[a][b] = SKK[6 * a + b] - 0.5 * (s[0][6 * a + b] + s[0][6 * b + a]);
        bool ok = true;
        for (int k = 0; k < 6; ++k) {
            const double piv = A[k][k];
            ok = ok && (piv > 0.0);
            const double ip = 1.0 / piv;
            double colk[6];
            for (int a = 0; a < 6; ++a) colk[a] = A[a][k];
            for (int b = 0; b < 6; ++b) A[k][b] = b == k ? ip : A[k][b] * ip;
            for (int a = 0; a < 6; ++a)
                if (a != k)
                    for (int b = 0; b < 6; ++b) A[a][b] = (b == k ? 0.0 : A[a][b]) - colk[a] * A[k][b];
        }
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) Ski[6 * a + b] = 0.5 * (A[a][b] + A[b][a]);
        if (!ok) atomicExch(fail, 1);
    }
}

}  // namespace asck
