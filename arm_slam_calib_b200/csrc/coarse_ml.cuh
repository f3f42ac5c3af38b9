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
//   top level: dense (<= pcg_ml_top_dim <= 1024 unknowns), explicit inverse by an in-place Gauss-Jordan sweep (one persistent kernel).
// The dense extrinsic border is eliminated exactly around V by bordering:
//   Y = V E_cK (six V-cycles per refresh),  S_k = S_KK - E_Kc Y,
//   B [r; r_K]:  t = V r,  z_K = S_k^-1 (r_K - E_Kc t),  z = t - Y z_K          ( = inverse of [V^-1 E_cK; E_Kc S_KK], SPD because V <= E_cc^-1 )
// Everything is a fixed linear SPD operator, so PCG stays exact; only the iteration count depends on it (measured on C2, first
// solve: exact dense coarse solve 90 iterations, this V-cycle with m = 4: ~110; scripts/dev/ml_proto6.py is the numpy prototype).
// Reductions use fixed trees / fixed member orders: results are run-to-run identical and identical on every rank.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace asck {

// Level description handed to the kernels (device pointers; handle-lifetime, so the captured PCG graph can keep them)
struct MlLevelDev {
    int n;                    // nodes (block rows)
    const int* row_ptr;       // [n + 1]
    const int* col;           // [nnzb]
    const double* Eb;         // blocks, stored TRANSPOSED inside each block row: element e of block k lives at bpos[k].x + e * bpos[k].y
                              // (.x = row_ptr[row] * N*N + position in the row, .y = blocks in the row), so that the lanes of a warp,
                              // one block each, read consecutive addresses (the phases of the V-cycle are bound by L1 sector requests)
    const int2* bpos;         // [nnzb]
    const double* Dinv;       // [n][N*N] inverses of the diagonal blocks
    const double* lmax;       // [1] power-iteration estimate of lambda_max(D^-1 E)
};

constexpr double kMlSafety = 1.2;   // upper end of the Chebyshev interval = kMlSafety * estimate (the power iteration converges from below)

// Row i of E times v: ONE BLOCK PER LANE (a block row holds ~27 blocks), so the dependent loads are row_ptr -> col -> v once per
// row instead of once per block (the phases of the V-cycle are latency-bound: < 10 MB of L2-resident data, ~20 dependent phases
// per PCG iteration).  The N partial results per lane are summed over the warp with a fixed butterfly; lane a < N returns entry a.
// v_eff = v (+ vc[agg] when vc != null: prolongation on the fly).  Vectors are read through L2 (__ldcg).
template <int N>
__device__ __forceinline__ void ml_load_x(const double* v, const double* vc, const int* __restrict__ agg, int c, double (&x)[N]) {
#pragma unroll
    for (int b = 0; b < N; ++b) x[b] = 0.0;
    if (c < 0) return;
#pragma unroll
    for (int b = 0; b < N; ++b) x[b] = __ldcg(v + (size_t)c * N + b);
    if (vc) {
        const double* cc = vc + (size_t)__ldg(agg + c) * N;
#pragma unroll
        for (int b = 0; b < N; ++b) x[b] += __ldcg(cc + b);
    }
}

// All loads of a round are ISSUED BEFORE the first use (the compiler otherwise interleaves eight-byte loads and DFMAs with only a
// handful of loads in flight: measured ~10 us per phase, seven L2 latencies back to back), and the column index / vector entries of
// the next round are prefetched.  Blocks with N*N > 64 entries are processed in row chunks to bound the register count.
template <int N>
__device__ __forceinline__ double bsr_row_dot(const int* __restrict__ row_ptr, const int* __restrict__ col, const double* __restrict__ Eb,
                                              const double* v, const double* vc, const int* __restrict__ agg, int i, int lane) {
    constexpr int RC = N * N <= 64 ? N : (48 / N > 0 ? 48 / N : 1);   // block rows per chunk
    const int k0 = __ldg(row_ptr + i), k1 = __ldg(row_ptr + i + 1), nb = k1 - k0;
    const double* rowbase = Eb + (size_t)k0 * N * N;
    double acc[N];
#pragma unroll
    for (int a = 0; a < N; ++a) acc[a] = 0.0;
    // a lane takes blocks lane, lane + 32, ...; the columns and vector entries of TWO rounds are fetched together (rows hold ~27
    // blocks on average and at most ~60: one dependent row_ptr -> col -> vector chain per row instead of one per round)
    for (int kl = lane; kl < nb; kl += 64) {
        const int kn = kl + 32;
        const int ca = __ldg(col + k0 + kl), cb = kn < nb ? __ldg(col + k0 + kn) : -1;
        double xa[N], xb[N];
        ml_load_x<N>(v, vc, agg, ca, xa);
        ml_load_x<N>(v, vc, agg, cb, xb);
#pragma unroll
        for (int a0 = 0; a0 < N; a0 += RC) {
            double e[RC * N], f[RC * N];
#pragma unroll
            for (int q = 0; q < RC * N; ++q) e[q] = (a0 * N + q < N * N) ? __ldg(rowbase + (size_t)(a0 * N + q) * nb + kl) : 0.0;
#pragma unroll
            for (int q = 0; q < RC * N; ++q) f[q] = (cb >= 0 && a0 * N + q < N * N) ? __ldg(rowbase + (size_t)(a0 * N + q) * nb + kn) : 0.0;
#pragma unroll
            for (int a = 0; a < RC; ++a) {
                if (a0 + a < N) {
                    double t = 0.0, u = 0.0;
#pragma unroll
                    for (int b = 0; b < N; ++b) { t = fma(e[a * N + b], xa[b], t); u = fma(f[a * N + b], xb[b], u); }
                    acc[a0 + a] += t + u;
                }
            }
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
// gather (first pre-smoothing step of a coarse level): r[i] is the sum of the finer level's residual over the members of aggregate i
// (restriction with a piecewise-constant prolongator); it is formed here and stored for the later steps.
// Vectors are read through L2 (__ldcg): inside the persistent V-cycle kernel other CTAs rewrote them a phase ago.
template <int N>
__device__ __forceinline__ void ml_cheb_node(const MlLevelDev& lv, double frac, int step, int zero_in, const double* r, const double* z_in,
                                             double* z_out, double* d, const double* zc, const int* agg, const double* fine_res,
                                             const int* __restrict__ mptr, const int* __restrict__ midx, double* r_store, int i, int lane) {
    // loads that do not depend on the product first
    double drow[N];
#pragma unroll
    for (int b = 0; b < N; ++b) drow[b] = lane < N ? __ldg(lv.Dinv + (size_t)i * N * N + lane * N + b) : 0.0;
    const double dold = (step > 0 && lane < N) ? __ldcg(d + (size_t)i * N + lane) : 0.0;
    double res = 0.0;
    if (fine_res) {
        const int m0 = __ldg(mptr + i), m1 = __ldg(mptr + i + 1);
        if (lane < N)
            for (int q = m0; q < m1; ++q) res += __ldcg(fine_res + (size_t)__ldg(midx + q) * N + lane);
        if (lane < N) r_store[(size_t)i * N + lane] = res;
    } else if (lane < N) res = __ldcg(r + (size_t)i * N + lane);
    const double c1 = __ldg(lv.lmax + 2 + 2 * step), c2 = __ldg(lv.lmax + 3 + 2 * step);   // Chebyshev coefficients of this step (ml_power_finish_kernel)
    double zin = 0.0;
    if (!zero_in) {
        if (lane < N) {
            zin = __ldcg(z_in + (size_t)i * N + lane);
            if (zc) zin += __ldcg(zc + (size_t)__ldg(agg + i) * N + lane);
        }
        const double p = bsr_row_dot<N>(lv.row_ptr, lv.col, lv.Eb, z_in, zc, agg, i, lane);
        res -= p;
    }
    double t = 0.0;
#pragma unroll
    for (int b = 0; b < N; ++b) t = fma(drow[b], __shfl_sync(kFull, res, b), t);
    if (lane < N) {
        const double dn = c1 * dold + c2 * t;
        d[(size_t)i * N + lane] = dn;
        z_out[(size_t)i * N + lane] = zin + dn;
    }
}

// residual of row i: out = r - E z
template <int N>
__device__ __forceinline__ void ml_residual_node(const MlLevelDev& lv, const double* r, const double* z, double* out, int i, int lane) {
    const double ri = lane < N ? __ldcg(r + (size_t)i * N + lane) : 0.0;
    const double p = bsr_row_dot<N>(lv.row_ptr, lv.col, lv.Eb, z, nullptr, nullptr, i, lane);
    if (lane < N) out[(size_t)i * N + lane] = ri - p;
}

// rc[A] = sum of the finer level's residual over the members of aggregate A (restriction onto the dense top level)
template <int N>
__device__ __forceinline__ void ml_sum_members_node(const int* __restrict__ mptr, const int* __restrict__ midx, const double* fine_res, double* rc, int A, int lane) {
    if (lane >= N) return;
    double acc = 0.0;
    for (int q = __ldg(mptr + A); q < __ldg(mptr + A + 1); ++q) acc += __ldcg(fine_res + (size_t)__ldg(midx + q) * N + lane);
    rc[(size_t)A * N + lane] = acc;
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

// ---- the whole V-cycle as ONE persistent kernel: its phases are separated by a grid barrier instead of kernel boundaries.
// On C2 the V-cycle is ~20 dependent phases over < 10 MB of L2-resident data.  kCycleCtas CTAs must be co-resident (they are:
// <= one per SM, 256 threads, no shared memory); a barrier that does not complete within ~2 s reports through `fail` instead of
// hanging.  Measured with the barrier trace (scripts/vc_trace.py): the barrier itself costs ~1.1 us; a phase is bound by the
// dependent L2 accesses of one block row (row_ptr -> col -> vector), which is why every load of a round is issued up front.
constexpr int kCycleMaxLevels = 8;
constexpr int kCycleThreads = 256;
struct MlCycleArgs {
    int nl, m, ntop;
    double frac;
    MlLevelDev lv[kCycleMaxLevels];
    const int* agg[kCycleMaxLevels];
    const int* mptr[kCycleMaxLevels];
    const int* midx[kCycleMaxLevels];
    int n_next[kCycleMaxLevels];
    const double* r0;                       // the caller's right-hand side (level 0)
    double* r[kCycleMaxLevels + 1];         // r[l], l >= 1: restricted residual of level l (r[nl]: right-hand side of the top level)
    double* rs[kCycleMaxLevels];            // residual r - E z of level l after pre-smoothing
    double* za[kCycleMaxLevels];
    double* zb[kCycleMaxLevels];
    double* d[kCycleMaxLevels];
    const double* Atop;
    double* ztop;
    unsigned* bar;                          // [2]: arrival counter, exit counter (both return to zero)
    unsigned long long* trace;              // debugging: per CTA and barrier, globaltimer at arrival and at release (null: off)
    int* fail;
    const int* done;
};

__device__ __forceinline__ unsigned long long ml_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One fence on each side and plain polling in between (an acquire load would invalidate L1 on EVERY poll: CCTL.IVALL).
__device__ __forceinline__ void ml_grid_barrier(unsigned* bar, unsigned& target, int* fail, unsigned long long* trace) {
    __syncthreads();   // the CTA's writes of this phase happen-before thread 0's fence below (bar.sync + cumulativity)
    if (threadIdx.x == 0) {
        if (trace) trace[(size_t)(target / gridDim.x) * gridDim.x * 2 + blockIdx.x * 2] = ml_globaltimer();
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        const long long t0 = clock64();
        while (*((volatile unsigned*)bar) < target) {
            if (clock64() - t0 > 4000000000ll) { *fail = 1; break; }
        }
        __threadfence();
        if (trace) trace[(size_t)(target / gridDim.x - 1) * gridDim.x * 2 + blockIdx.x * 2 + 1] = ml_globaltimer();
    }
    __syncthreads();
}

template <int N>
__global__ void __launch_bounds__(kCycleThreads) ml_vcycle_kernel(const __grid_constant__ MlCycleArgs a) {
    if (a.done && *a.done) return;
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (kCycleThreads / 32) + (threadIdx.x >> 5), nw = gridDim.x * (kCycleThreads / 32);
    unsigned target = 0;
    const double* cur[kCycleMaxLevels + 1];
    for (int l = 0; l < a.nl; ++l) {   // down: pre-smooth from zero, residual
        const double* r = l == 0 ? a.r0 : a.r[l];
        double* bufs[2] = {a.za[l], a.zb[l]};
        for (int k = 0; k < a.m; ++k) {
            if (k > 0 || l > 0) ml_grid_barrier(a.bar, target, a.fail, a.trace);
            const bool gather = k == 0 && l > 0;   // restriction of the finer residual, formed on the fly and stored in r[l]
            for (int i = gw; i < a.lv[l].n; i += nw)
                ml_cheb_node<N>(a.lv[l], a.frac, k, k == 0 ? 1 : 0, r, bufs[(k + 1) & 1], bufs[k & 1], a.d[l], nullptr, nullptr,
                                gather ? a.rs[l - 1] : nullptr, gather ? a.mptr[l - 1] : nullptr, gather ? a.midx[l - 1] : nullptr, a.r[l], i, lane);
        }
        cur[l] = bufs[(a.m - 1) & 1];
        ml_grid_barrier(a.bar, target, a.fail, a.trace);
        for (int i = gw; i < a.lv[l].n; i += nw) ml_residual_node<N>(a.lv[l], r, cur[l], a.rs[l], i, lane);
    }
    if (a.nl > 0) {   // right-hand side of the top level
        ml_grid_barrier(a.bar, target, a.fail, a.trace);
        for (int A = gw; A < a.n_next[a.nl - 1]; A += nw) ml_sum_members_node<N>(a.mptr[a.nl - 1], a.midx[a.nl - 1], a.rs[a.nl - 1], a.r[a.nl], A, lane);
        ml_grid_barrier(a.bar, target, a.fail, a.trace);
    }
    for (int row = gw; row < a.ntop; row += nw) ml_gemv_row(a.Atop, a.nl == 0 ? a.r0 : a.r[a.nl], a.ztop, a.ntop, row, lane);
    cur[a.nl] = a.ztop;
    for (int l = a.nl - 1; l >= 0; --l) {   // up: prolong (inside the first step), post-smooth
        const double* r = l == 0 ? a.r0 : a.r[l];
        const double* zi = cur[l];
        for (int k = 0; k < a.m; ++k) {
            ml_grid_barrier(a.bar, target, a.fail, a.trace);
            double* zo = zi == a.za[l] ? a.zb[l] : a.za[l];
            for (int i = gw; i < a.lv[l].n; i += nw)
                ml_cheb_node<N>(a.lv[l], a.frac, k, 0, r, zi, zo, a.d[l], k == 0 ? cur[l + 1] : nullptr, k == 0 ? a.agg[l] : nullptr, nullptr, nullptr, nullptr,
                                nullptr, i, lane);
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
                    const int2* __restrict__ bpos, double* __restrict__ Eb) {
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
    const int2 blo = bpos[plo], bup = bpos[pup];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        const int e = lane + 32 * i;
        if (e < N * N) {
            Eb[blo.x + (size_t)e * blo.y] = -acc[i];
            if (pup != plo) Eb[bup.x + (size_t)((e % N) * N + e / N) * bup.y] = -acc[i];
        }
    }
}

// diagonal blocks += sum_{p in cluster} (B_pp + lambda I)   (rank 0 only in a multi-GPU run); thread per (c, a, b)
__global__ void ml_diag_l1_kernel(int nC, int N, double lambda, const int* __restrict__ cl_ptr, const double* __restrict__ Bpp,
                                  const int* __restrict__ diag_pos, const int2* __restrict__ bpos, double* __restrict__ Eb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nC * N * N) return;
    const int c = i / (N * N), a = (i / N) % N, b = i % N;
    double s = 0.0;
    for (int p = cl_ptr[c]; p < cl_ptr[c + 1]; ++p) s += Bpp[(size_t)p * N * N + a * N + b];
    if (a == b) s += lambda * (cl_ptr[c + 1] - cl_ptr[c]);
    const int2 bp = bpos[diag_pos[c]];
    Eb[bp.x + (size_t)(a * N + b) * bp.y] += s;
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
__global__ void ml_coarsen_kernel(int nnzb_c, int NN, const int* __restrict__ cptr, const int* __restrict__ cidx, const int2* __restrict__ bpos_f,
                                  const double* __restrict__ Ef, const int2* __restrict__ bpos_c, double* __restrict__ Ec) {
    const int k = blockIdx.x, e = threadIdx.x;
    if (k >= nnzb_c || e >= NN) return;
    double s = 0.0;
    for (int q = cptr[k]; q < cptr[k + 1]; ++q) { const int2 bp = bpos_f[cidx[q]]; s += Ef[bp.x + (size_t)e * bp.y]; }
    const int2 bc = bpos_c[k];
    Ec[bc.x + (size_t)e * bc.y] = s;
}

// Dinv[i] = inverse of the N x N diagonal block (SPD): Gauss-Jordan without pivoting, one thread per node
template <int N>
__global__ void __launch_bounds__(64) ml_dinv_kernel(int n, const int* __restrict__ diag_pos, const int2* __restrict__ bpos, const double* __restrict__ Eb,
                                                     double* __restrict__ Dinv, int* __restrict__ fail) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double A[N][N];
    const int2 bp = bpos[diag_pos[i]];
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) A[a][b] = Eb[bp.x + (size_t)(a * N + b) * bp.y];
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
// also tabulates the Chebyshev coefficients of the smoother for the interval [hi / frac, hi], hi = kMlSafety * lmax:
//   lmax[2 + 2 k] = c1_k, lmax[3 + 2 k] = c2_k  (d <- c1_k d + c2_k D^-1 res; c1_0 = 0, c2_0 = 1 / theta)
__global__ void __launch_bounds__(256) ml_power_finish_kernel(const double* __restrict__ partA, const double* __restrict__ partB, int nPart,
                                                               double* __restrict__ lmax, double frac, int degree) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nPart; i += 256) { a += partA[i]; b += partB[i]; }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double lm = (sa[0] > 0.0 && sb[0] > 0.0) ? sqrt(sb[0] / sa[0]) : 1.0;
        lmax[0] = lm;
        const double hi = kMlSafety * lm, lo = hi / frac, theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
        double rho = 1.0 / sigma;
        lmax[2] = 0.0; lmax[3] = 1.0 / theta;
        for (int k = 1; k < degree; ++k) { const double rn = 1.0 / (2.0 * sigma - rho); lmax[2 + 2 * k] = rn * rho; lmax[3 + 2 * k] = 2.0 * rn / delta; rho = rn; }
    }
}

// dense copy of the last level: A[n N][n N] row-major from its BSR blocks
__global__ void ml_top_assemble_kernel(int n, int N, const int* __restrict__ row_ptr, const int* __restrict__ col, const int2* __restrict__ bpos,
                                       const double* __restrict__ Eb, double* __restrict__ A) {
    const int i = blockIdx.x, e = threadIdx.x;
    if (i >= n || e >= N * N) return;
    const int ld = n * N, a = e / N, b = e % N;
    for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) { const int2 bp = bpos[k]; A[(size_t)(i * N + a) * ld + col[k] * N + b] = Eb[bp.x + (size_t)e * bp.y]; }
}

// In-place inverse of the dense SPD top-level matrix (n <= kTopMaxDim) by a Gauss-Jordan sweep without pivoting, as ONE persistent
// kernel: every warp of the grid owns one matrix row for the whole sweep and keeps it in shared memory; the pivot row of step k is
// broadcast through a two-slot global buffer (written by its owner at the end of step k - 1), so a step costs one grid barrier
// (~1.1 us) plus ~1 us of work.  n = 876 (C2 with a one-level hierarchy): ~2 ms instead of 25 ms with one cluster and three
// cooperative-groups cluster barriers per step.
constexpr int kTopMaxDim = 1024;
constexpr int kTopThreads = 256;                       // 8 warps = 8 rows per CTA
constexpr int kTopCtas = kTopMaxDim / (kTopThreads / 32);
__global__ void __launch_bounds__(kTopThreads) ml_top_invert_kernel(double* A, int n, double* piv /*[2][n]*/, unsigned* bar, int* fail) {
    extern __shared__ double sm_top[];
    const int wpc = kTopThreads / 32, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* prow = sm_top;                               // [n] scaled pivot row of the step
    double* myrow = sm_top + n + (size_t)warp * n;       // [n] this warp's row
    const int r = blockIdx.x * wpc + warp;               // the row this warp owns (r >= n: idle, still takes part in the barriers)
    unsigned target = 0;
    bool bad = false;
    if (r < n)
        for (int c = lane; c < n; c += 32) myrow[c] = __ldcg(A + (size_t)r * n + c);
    if (r == 0)
        for (int c = lane; c < n; c += 32) __stcg(piv + c, myrow[c]);
    ml_grid_barrier(bar, target, fail, nullptr);
    for (int k = 0; k < n; ++k) {
        const double* pk = piv + (size_t)(k & 1) * n;
        const double pivot = __ldcg(pk + k);
        bad = bad || !(pivot > 0.0);
        const double ip = 1.0 / pivot;
        for (int c = threadIdx.x; c < n; c += kTopThreads) prow[c] = c == k ? ip : __ldcg(pk + c) * ip;
        __syncthreads();
        if (r < n) {
            if (r == k) {
                for (int c = lane; c < n; c += 32) myrow[c] = prow[c];
            } else {
                const double f = myrow[k];
                __syncwarp();
                for (int c = lane; c < n; c += 32) myrow[c] = (c == k ? 0.0 : myrow[c]) - f * prow[c];
            }
            if (r == k + 1) {   // next pivot row: publish it
                __syncwarp();
                double* pn = piv + (size_t)((k + 1) & 1) * n;
                for (int c = lane; c < n; c += 32) __stcg(pn + c, myrow[c]);
            }
        }
        ml_grid_barrier(bar, target, fail, nullptr);   // (also orders prow reuse inside the CTA)
    }
    if (r < n)
        for (int c = lane; c < n; c += 32) A[(size_t)r * n + c] = myrow[c];
    if (bad && threadIdx.x == 0 && blockIdx.x == 0) atomicExch(fail, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(bar + 1, 1u) == gridDim.x - 1) { bar[0] = 0; bar[1] = 0; __threadfence(); }
    }
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
