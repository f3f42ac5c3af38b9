// sm_100a kernels of the batch optimisation path.
//
// Data layout in HBM (all fp64 unless noted; k = position in the POSE-major observation order,
// kk = position in the LANDMARK-major order; both orders are built once in asc_set_problem):
//   cam   [P][12]      camera pose per keyframe  R_cam row-major (9) | t_cam (3)       (FK kernel)
//   chain [P][6N]      world joint axis a_i (3) | joint origin o_i (3) per joint        (FK kernel)
//   p_uv  [M] double2, p_lm [M] int32, p_pose [M] int32          pose-major static observation data
//   l_uv  [M] double2, l_pose [M] int32, l_perm [M] int32        landmark-major static data,
//                                                                l_perm[kk] = k of the same factor
//   J     [(N+3)][M] double2   whitened+reweighted Jacobian planes, pose-major: the flat vector
//                              [Jq row0 (N) | Jq row1 (N) | Jl row0 (3) | Jl row1 (3)] of observation k,
//                              component pair i at J[i*M + k] -> every load is a coalesced LDG.128
//   lm4   [L][4], u4 [L][4], v4 [M][4]   3-vectors padded to one 32-byte sector for gathers
//   C [L][6] (sym), gl [L][3], WK [L][18], Cinv [L][6]            landmark blocks
//   Bpp [P][N*N], BpK [P][N*6], gp [P][N], Minv [P][N*N]          pose blocks
// Reductions are two-level with a fixed tree (no atomics), so results are run-to-run identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace asck {

constexpr int kWarpsPerCta = 4;
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr int kPartCols = 32;   // width of the linearisation partial-sum rows (padded)
constexpr int kPcgCols = 8;     // width of the per-CTA partial rows written inside the PCG loop
constexpr int kReduceThreads = 1024;
constexpr unsigned kFull = 0xffffffffu;

struct Camera {
    double fx, fy, cx, cy;
    double inv_sigma_px, k2;      // 1/sigma_px, cauchy k^2
    double inv_sigma_lm;          // 1/sigma_landmark
};

template <int N>
struct ChainConst {               // host-precomposed serial chain
    double F0[12];                // T_base * T_p2j[0]
    double Lk[N][12];             // inverse(T_c2j[i]) * T_p2j[i+1]   (last: inverse(T_c2j[N-1]))
    double axis[N][3];
};

struct PriorConst {
    double x_prior[12];
    double inv_sigma_x[6];
    double inv_sigma_enc[20];
    double inv_sigma_link[2][20];   // DriftFactor / VelocityFactor sigmas (only read when the link table is passed)
    int continuous_mask;          // bit i set: joint i is continuous (RobotConfig.h:72-77)
    int use_dead_band;
    double dead_band;
    int chart;
};

// ------------------------------------------------------------------ peer-memory exchange (multi-GPU PCG)
// One process per GPU; every rank owns one exchange buffer (cudaMalloc + CUDA IPC, mapped by all peers over NVLink):
//   [ header 256 B: one sequence number per peer (words 0..7), the local sequence counter (word 32) | slot 0 | slot 1 ],
//   slot = [ y partial (PN) | pad to PN+8 | rowC (32) | rowB (32) ]
// Per PCG iteration a rank writes its shard's partial product into slot (seq & 1) of ITS OWN buffer (pass C), publishes
// seq to every peer's flag array (pcg_signal_kernel) and the consumers (pcg_mid_kernel, pcg_cluster_update_kernel) sum the
// slots of all ranks with peer loads in rank order -- the same order on every rank, so all ranks hold bitwise identical
// vectors and stay in lock-step.  No collective library call and no host involvement inside the PCG loop, which is
// what lets the multi-GPU loop run as one CUDA graph.  Two slots suffice: a rank can only overwrite slot s two
// iterations later, after every peer has published the iteration in between, i.e. after it finished reading slot s.
constexpr int kMaxPeers = 8;
struct PeerComm {
    double* x[kMaxPeers];            // exchange buffer of every rank (slot 0 base), peer-mapped
    unsigned* peer_flags[kMaxPeers]; // flag array of every rank
    unsigned* ctr;                   // local: last sequence number this rank published
    int* fail;                       // local: set when a wait timed out
    long long stride;                // doubles per slot
    long long row_off;               // offset of rowC inside a slot (= PN + 8)
    int rank, world;
};

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ double ld_peer_f64(const double* p) {   // system-scope load: never served from a stale L1 line
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// sum over ranks (fixed order) of element i of the current slot
__device__ __forceinline__ double peer_sum(const PeerComm& pc, unsigned seq, long long i) {
    const long long off = (long long)(seq & 1u) * pc.stride + i;
    double s = 0.0;
    for (int r = 0; r < pc.world; ++r) s += ld_peer_f64(pc.x[r] + off);
    return s;
}

// ------------------------------------------------------------------ tiny device helpers
__device__ __forceinline__ double2 ldg2(const double2* p) { return __ldg(p); }

// One 32-byte sector per instruction (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256).  The padded 3-vector tables (lm4, u4, v4,
// lml) are gathered one sector per lane; a 128-bit + 64-bit pair costs two L1 tag passes per sector, this costs one.
struct d4 { double x, y, z, w; };
__device__ __forceinline__ d4 ldg256(const double* p) {   // read-only data, 32-byte aligned
    d4 r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ void mul_xf(const double* A, const double* B, double* O) {   // 3x4 [R|t] row-major
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double s = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
            if (c == 3) s += A[4 * r + 3];
            O[4 * r + c] = s;
        }
    }
}

__device__ __forceinline__ void rodrigues(const double* ax, double th, double* R) {   // 3x3 row-major
    double s, c;
    sincos(th, &s, &c);
    const double v = 1.0 - c, x = ax[0], y = ax[1], z = ax[2];
    R[0] = c + x * x * v; R[1] = x * y * v - z * s; R[2] = x * z * v + y * s;
    R[3] = y * x * v + z * s; R[4] = c + y * y * v; R[5] = y * z * v - x * s;
    R[6] = z * x * v - y * s; R[7] = z * y * v + x * s; R[8] = c + z * z * v;
}

// [DEP] Rot3 / Pose3 charts; x12 is [R|t] row-major 3x4
__device__ inline void rot3_logmap(const double* A, double* w) {   // A 3x3 row-major
    const double tr = A[0] + A[4] + A[8];
    if (fabs(tr + 1.0) < 1e-10) {
        const double PI = 3.14159265358979323846;
        if (fabs(A[8] + 1.0) > 1e-10) { const double s = PI / sqrt(2.0 + 2.0 * A[8]); w[0] = s * A[2]; w[1] = s * A[5]; w[2] = s * (1.0 + A[8]); }
        else if (fabs(A[4] + 1.0) > 1e-10) { const double s = PI / sqrt(2.0 + 2.0 * A[4]); w[0] = s * A[1]; w[1] = s * (1.0 + A[4]); w[2] = s * A[7]; }
        else { const double s = PI / sqrt(2.0 + 2.0 * A[0]); w[0] = s * (1.0 + A[0]); w[1] = s * A[3]; w[2] = s * A[6]; }
        return;
    }
    double mag;
    const double tr3 = tr - 3.0;
    if (tr3 < -1e-7) { const double th = acos((tr - 1.0) / 2.0); mag = th / (2.0 * sin(th)); }
    else mag = 0.5 - tr3 * tr3 / 12.0;
    w[0] = mag * (A[7] - A[5]); w[1] = mag * (A[2] - A[6]); w[2] = mag * (A[3] - A[1]);
}

__device__ inline void pose3_local(int chart, const double* X0, const double* X1, double* xi) {
    double A[9], d[3], t[3];   // A = R0^T R1, t = R0^T (t1 - t0)
#pragma unroll
    for (int r = 0; r < 3; ++r) d[r] = X1[4 * r + 3] - X0[4 * r + 3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) A[3 * r + c] = X0[r] * X1[c] + X0[4 + r] * X1[4 + c] + X0[8 + r] * X1[8 + c];
        t[r] = X0[r] * d[0] + X0[4 + r] * d[1] + X0[8 + r] * d[2];
    }
    if (chart == 0) {   // inverse Cayley
        const double k = 2.0 / (1.0 + A[0] + A[4] + A[8]);
        xi[0] = k * (A[7] - A[5]); xi[1] = k * (A[2] - A[6]); xi[2] = k * (A[3] - A[1]);
    } else rot3_logmap(A, xi);
    if (chart == 2) {   // Pose3::Logmap translation part
        const double th = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
        if (th < 1e-10) { xi[3] = t[0]; xi[4] = t[1]; xi[5] = t[2]; }
        else {
            const double n[3] = {xi[0] / th, xi[1] / th, xi[2] / th};
            const double WT[3] = {n[1] * t[2] - n[2] * t[1], n[2] * t[0] - n[0] * t[2], n[0] * t[1] - n[1] * t[0]};
            const double WWT[3] = {n[1] * WT[2] - n[2] * WT[1], n[2] * WT[0] - n[0] * WT[2], n[0] * WT[1] - n[1] * WT[0]};
            const double Tan = tan(0.5 * th);
#pragma unroll
            for (int r = 0; r < 3; ++r) xi[3 + r] = t[r] - (0.5 * th) * WT[r] + (1 - th / (2. * Tan)) * WWT[r];
        }
    } else { xi[3] = t[0]; xi[4] = t[1]; xi[5] = t[2]; }
}

__device__ inline void pose3_retract(int chart, const double* X, const double* xi, double* O) {
    double E[12];   // increment [R|t]
    const double x = xi[0], y = xi[1], z = xi[2];
    if (chart == 0) {
        const double x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
        const double f = 1.0 / (4.0 + x2 + y2 + z2), f2 = 2.0 * f;
        E[0] = (4 + x2 - y2 - z2) * f; E[1] = (xy - 2 * z) * f2; E[2] = (xz + 2 * y) * f2;
        E[4] = (xy + 2 * z) * f2; E[5] = (4 - x2 + y2 - z2) * f; E[6] = (yz - 2 * x) * f2;
        E[8] = (xz - 2 * y) * f2; E[9] = (yz + 2 * x) * f2; E[10] = (4 - x2 - y2 + z2) * f;
    } else {
        const double th = sqrt(x * x + y * y + z * z);
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (!(th < 1e-10)) { const double ax[3] = {x / th, y / th, z / th}; rodrigues(ax, th, R); }
#pragma unroll
        for (int r = 0; r < 3; ++r) { E[4 * r] = R[3 * r]; E[4 * r + 1] = R[3 * r + 1]; E[4 * r + 2] = R[3 * r + 2]; }
    }
    if (chart == 2) {
        const double th = sqrt(x * x + y * y + z * z);
        const double* v = xi + 3;
        if (th < 1e-10) { E[3] = v[0]; E[7] = v[1]; E[11] = v[2]; }
        else {
            const double n[3] = {x / th, y / th, z / th};
            const double vn = n[0] * v[0] + n[1] * v[1] + n[2] * v[2];
            const double c[3] = {n[1] * v[2] - n[2] * v[1], n[2] * v[0] - n[0] * v[2], n[0] * v[1] - n[1] * v[0]};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double Rc = E[4 * r] * c[0] + E[4 * r + 1] * c[1] + E[4 * r + 2] * c[2];
                E[4 * r + 3] = (c[r] - Rc) / th + vn * n[r];
            }
        }
    } else { E[3] = xi[3]; E[7] = xi[4]; E[11] = xi[5]; }   // first order: t + R v
    mul_xf(X, E, O);
}

// ------------------------------------------------------------------ K0: forward kinematics table
// One thread per keyframe.  Replaces the three DART setPositions/FK passes the reference does for
// EVERY factor (RobotProjectionFactor.h:206,:235,:257): FK is evaluated once per pose.
template <int N, bool WITH_CHAIN>
__global__ void __launch_bounds__(128) fk_kernel(ChainConst<N> cc, const double* __restrict__ q, const double* __restrict__ X,
                                                 int P, double* __restrict__ cam, double* __restrict__ chain) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double F[12], T[12], R[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) F[i] = cc.F0[i];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double* ax = cc.axis[i];
        if (WITH_CHAIN) {
            double* o = chain + (size_t)p * 6 * N + 6 * i;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                o[r] = F[4 * r] * ax[0] + F[4 * r + 1] * ax[1] + F[4 * r + 2] * ax[2];
                o[3 + r] = F[4 * r + 3];
            }
        }
        double Rq[9];
        rodrigues(ax, q[(size_t)p * N + i], Rq);
#pragma unroll
        for (int r = 0; r < 3; ++r) { R[4 * r] = Rq[3 * r]; R[4 * r + 1] = Rq[3 * r + 1]; R[4 * r + 2] = Rq[3 * r + 2]; R[4 * r + 3] = 0; }
        mul_xf(F, R, T);
        mul_xf(T, cc.Lk[i], F);
    }
    double Xl[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) Xl[i] = X[i];
    mul_xf(F, Xl, T);   // T_cam = T_link * X   (RobotProjectionFactor.h:208)
    double* o = cam + (size_t)p * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) { o[3 * r] = T[4 * r]; o[3 * r + 1] = T[4 * r + 1]; o[3 * r + 2] = T[4 * r + 2]; o[9 + r] = T[4 * r + 3]; }
}

// ------------------------------------------------------------------ projection factor core
// Everything RobotProjectionFactor::project + evaluateError + Robust::WhitenSystem produce for one
// factor except J_q (which needs the chain).  R,t = camera pose in the world.
struct ProjOut {
    double b0, b1;       // whitened, reweighted right-hand side  b = -sqrt(w) r / sigma
    double jl[6];        // whitened, reweighted 2x3
    double jk[12];       // whitened, reweighted 2x6
    double err;          // 0.5 w |r/sigma|^2
    bool cheir;
};

template <bool WITH_J, bool WITH_JK>
__device__ __forceinline__ void project_factor(const Camera& cm, const double* R, const double* t, double lx, double ly, double lz,
                                               double zu, double zv, ProjOut& o) {
    const double d0 = lx - t[0], d1 = ly - t[1], d2 = lz - t[2];
    const double px = R[0] * d0 + R[3] * d1 + R[6] * d2;
    const double py = R[1] * d0 + R[4] * d1 + R[7] * d2;
    const double pz = R[2] * d0 + R[5] * d1 + R[8] * d2;
    const bool bad = pz < 0.01;                              // RobotProjectionFactor.h:213
    const double iz = 1.0 / pz, u = px * iz, v = py * iz;
    const double pu = bad ? 2.0 * cm.fx : cm.fx * u + cm.cx; // :219 returns (2fx, 2fx)
    const double pv = bad ? 2.0 * cm.fx : cm.fy * v + cm.cy;
    const double r0 = (pu - zu) * cm.inv_sigma_px, r1 = (pv - zv) * cm.inv_sigma_px;
    const double n2 = r0 * r0 + r1 * r1;
    const double w = cm.k2 / (cm.k2 + n2), sw = sqrt(w);
    o.b0 = -sw * r0; o.b1 = -sw * r1;
    o.err = 0.5 * w * n2;
    o.cheir = bad;
    if (WITH_J) {
        const double s = bad ? 0.0 : sw * cm.inv_sigma_px;   // zero Jacobians behind the camera (:215-217)
        const double su = s * cm.fx * iz, sv = s * cm.fy * iz;
        const double uu = bad ? 0.0 : u, vv = bad ? 0.0 : v;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            o.jl[c] = bad ? 0.0 : su * (R[3 * c] - uu * R[3 * c + 2]);
            o.jl[3 + c] = bad ? 0.0 : sv * (R[3 * c + 1] - vv * R[3 * c + 2]);
        }
        if (WITH_JK) {
            const double a = s * cm.fx, c2 = s * cm.fy, iz2 = bad ? 0.0 : iz;
            o.jk[0] = a * (uu * vv); o.jk[1] = a * (-1.0 - uu * uu); o.jk[2] = a * vv;
            o.jk[3] = a * (-iz2); o.jk[4] = 0.0; o.jk[5] = a * (iz2 * uu);
            o.jk[6] = c2 * (1.0 + vv * vv); o.jk[7] = c2 * (-uu * vv); o.jk[8] = c2 * (-uu);
            o.jk[9] = 0.0; o.jk[10] = c2 * (-iz2); o.jk[11] = c2 * (iz2 * vv);
        }
    }
}

// ------------------------------------------------------------------ warp-level bilinear accumulation
// acc(c, c2) += sum_r U[r][c] * V[r][c2] over the 32 observations of a tile, for the D x D symmetric
// result, each unordered pair owned by exactly one lane: lane column c accumulates c2 = (c + t) mod D,
// t = 0..D/2.  For D <= 16 the two half-warps split the two residual rows and are added at the end.
template <int D>
struct GramCfg {
    static constexpr int GS = D <= 16 ? 16 : 32;     // lanes per group
    static constexpr int G = 32 / GS;                // groups (rows handled in parallel)
    static constexpr int H = D / 2 + 1;              // offsets per lane
    static constexpr int LD = (D % 2) ? D : D + 1;   // odd leading dimension: conflict-free column reads
};

template <int D, bool SYM_UV>
__device__ __forceinline__ void gram_tile(const double* __restrict__ U, const double* __restrict__ V, int count, int lane,
                                          double (&acc)[GramCfg<D>::H]) {
    // U, V: [2][32][LD] in shared memory (V == U when SYM_UV)
    typedef GramCfg<D> Cf;
    const int g = lane / Cf::GS, c = lane % Cf::GS;
    if (c >= D) return;
    for (int o = 0; o < count; ++o) {
#pragma unroll
        for (int rr = 0; rr < 2 / Cf::G; ++rr) {
            const int r = Cf::G == 2 ? g : rr;
            const double* vrow = V + (r * 32 + o) * Cf::LD;
            const double mine = SYM_UV ? vrow[c] : U[(r * 32 + o) * Cf::LD + c];
#pragma unroll
            for (int t = 0; t < Cf::H; ++t) {
                int c2 = c + t;
                if (c2 >= D) c2 -= D;
                acc[t] = fma(mine, vrow[c2], acc[t]);
            }
        }
    }
}

template <int D>
__device__ __forceinline__ void gram_finish(double (&acc)[GramCfg<D>::H]) {
    if (GramCfg<D>::G == 2) {
#pragma unroll
        for (int t = 0; t < GramCfg<D>::H; ++t) acc[t] += __shfl_xor_sync(kFull, acc[t], 16);
    }
}

// ---- Warp totals of per-lane accumulators: 32 values per lane are summed across the warp by a butterfly
// "transpose-reduce" (31 shuffles for 32 sums, fixed tree); afterwards lane l holds the warp total of value number l.
// Used once per keyframe by K1 and once per warp by K1b.
template <int D>
__host__ __device__ constexpr int tri_row(int e) { int a = 0, rem = e; while (rem >= D - a) { rem -= D - a; ++a; } return a; }
template <int D>
__host__ __device__ constexpr int tri_col(int e) { int a = 0, rem = e; while (rem >= D - a) { rem -= D - a; ++a; } return a + rem; }

template <int D>
struct TriTable {   // compile-time (row, col) of the e-th upper-triangular entry, row-major
    int row[D * (D + 1) / 2 + 32], col[D * (D + 1) / 2 + 32];
    __host__ __device__ constexpr TriTable() : row(), col() {
        int e = 0;
        for (int a = 0; a < D; ++a)
            for (int b = a; b < D; ++b) { row[e] = a; col[e] = b; ++e; }
        for (; e < D * (D + 1) / 2 + 32; ++e) { row[e] = 0; col[e] = 0; }
    }
};

template <int W>
__device__ __forceinline__ void butterfly_step(double (&v)[32], int lane) {
    const bool upper = (lane & W) != 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double lo = v[i], hi = v[i + W];
        const double send = upper ? lo : hi, keep = upper ? hi : lo;
        v[i] = keep + __shfl_xor_sync(kFull, send, W);
    }
}
// v[i] on every lane -> return sum over lanes of v[lane]
__device__ __forceinline__ double transpose_reduce32(double (&v)[32], int lane) {
    butterfly_step<16>(v, lane); butterfly_step<8>(v, lane); butterfly_step<4>(v, lane); butterfly_step<2>(v, lane); butterfly_step<1>(v, lane);
    return v[0];
}

// ------------------------------------------------------------------ K1 (+K2b): linearise, pose-major
// One warp per keyframe (W warps for N >= 7).  The warp stages the keyframe's FK row (camera pose + joint axes /
// origins) in shared memory and walks the keyframe's observations 32 at a time, one projection factor per lane:
// residual, J_q, J_l, J_K, Robust-Cauchy whitening (RobotProjectionFactor.h:96-259 + SURVEY rows A1-A3).  J_q / J_l go
// to HBM as double2 planes.  The keyframe's blocks B_pp = sum J_q^T J_q, B_pK = sum J_q^T J_K, g_p = sum J_q^T b are
// the first NEq entries (rows 0..N-1) of the upper triangle of the Gram matrix of the whitened rows [J_q | J_K | b];
// every lane accumulates the products of ITS OWN factors for the warp's slice of those entries in registers across
// all tiles, and one butterfly transpose-reduce per 32 entries at the END of the keyframe (fixed tree) leaves entry
// l + 32 t on lane l.  The remaining 28 entries (B_KK, g_K, projection error) are global sums, not per-keyframe
// ones: proj_kk_kernel below accumulates them over a flat grid.
// History (profiles/README.md): reducing across the warp once per TILE -- shared-memory tiles, 4x4 register blocks,
// per-tile transpose-reduce -- was bound by LDS wavefronts / shuffle issue at ~1240 warp instructions per tile
// (31 % of the HBM roofline); splitting the 91 entries over two warps that both evaluate the factor doubled the
// fp64 work.  Per-lane accumulation needs 2 DFMAs per entry and tile and nothing else.
template <int N>
struct LinCfg {
    static constexpr int D = N + 7;
    static constexpr int NE = D * (D + 1) / 2;          // unique entries of the Gram matrix of [J_q | J_K | b]
    static constexpr int NEq = N * D - N * (N - 1) / 2; // ... of its rows 0..N-1 (the per-keyframe blocks)
    static constexpr int W = (NEq + 63) / 64;           // warps per keyframe (<= 64 accumulators per lane)
    static constexpr int EPW = (NEq + W - 1) / W;       // entries per warp
    static constexpr int NB = (EPW + 31) / 32;          // transpose-reduce batches at the end of a keyframe
    static constexpr int KF = W <= 2 ? 4 / W : 1;       // keyframes per CTA
    static constexpr int kWarps = W * KF;
    static constexpr int kThreads = 32 * kWarps;
    static constexpr int kPerWarpDoubles = 12 + 6 * N;
    static constexpr size_t kBytes = sizeof(double) * (kPerWarpDoubles * kWarps + KF);
};

// Link factors between consecutive configurations: kind 0 = DriftFactor(q_{p-1}, q_p) (DriftFactor.h:44-61, meas = q1Encoders -
// q0Encoders), kind 1 = VelocityFactor (VelocityFactor.h:40-62, meas = qDot).  has[k][p] != 0 (has[k][0] = has[k][P] = 0).
// residual (q_p - q_{p-1}) - meas[k][p], J0 = -I, J1 = +I, Diagonal noise: B_{p-1,p-1} += W, B_pp += W, B_{p-1,p} = -W, W = diag(1/sigma^2).
struct DriftTab {
    const uint8_t* has[2];    // null: no factors of that kind
    const double* meas[2];
    const double* w[2];       // [N] 1 / sigma^2
    __host__ __device__ bool any() const { return has[0] != nullptr || has[1] != nullptr; }
};
// whitened residual of the link (p-1, p) of kind k, joint i
__device__ __forceinline__ double drift_residual(const PriorConst& pc, const DriftTab& dt, int k, const double* __restrict__ q, int N, int p, int i) {
    return ((q[(size_t)p * N + i] - q[(size_t)(p - 1) * N + i]) - dt.meas[k][(size_t)p * N + i]) * pc.inv_sigma_link[k][i];
}

// whitened EncoderFactor residual of joint i (EncoderFactor.h:60-99): (q - enc) / sigma, SO(2)-wrapped if continuous
__device__ __forceinline__ double encoder_residual(const PriorConst& pc, double qq, double ee, int i) {
    double d = qq - ee;
    if ((pc.continuous_mask >> i) & 1) d = atan2(sin(qq - ee), cos(qq - ee));
    if (pc.use_dead_band) d = fabs(d) < pc.dead_band ? 0.0 : (d < 0 ? d + pc.dead_band : d - pc.dead_band);
    return d * pc.inv_sigma_enc[i];
}

template <int N, int ROLE, bool STORE_B>
__device__ __forceinline__ void linearize_role(const Camera& cm, const PriorConst& pc, int p, int64_t M, const int* __restrict__ pose_ptr,
                                               const int* __restrict__ p_lm, const double2* __restrict__ p_uv, const double* __restrict__ lm4,
                                               const double* __restrict__ q, const double* __restrict__ enc, bool with_enc, DriftTab dt,
                                               double2* __restrict__ J, double* __restrict__ Bpp, double* __restrict__ BpK, double* __restrict__ gp,
                                               double2* __restrict__ b_out, unsigned long long* __restrict__ cheir_count,
                                               const double* s_cam, const double* s_chain, int lane) {
    typedef LinCfg<N> Cf;
    constexpr int D = Cf::D, NEq = Cf::NEq, E0 = ROLE * Cf::EPW, E1 = (E0 + Cf::EPW < NEq) ? E0 + Cf::EPW : NEq, CNT = E1 - E0;
    constexpr bool kStore = ROLE == Cf::W - 1;           // this role writes the Jacobian planes
    constexpr TriTable<D> tri = TriTable<D>();
    double acc[CNT > 0 ? CNT : 1];
#pragma unroll
    for (int i = 0; i < CNT; ++i) acc[i] = 0.0;
    unsigned ncheir = 0;
    const int k0 = pose_ptr[p], k1 = pose_ptr[p + 1];
    // software pipeline: the landmark gather of tile t+1 (index -> position, two dependent loads) is issued before the
    // arithmetic of tile t
    const int klast = max(k1 - 1, k0);
    const bool any = k1 > k0;
    int j_nxt = any ? __ldg(p_lm + min(k0 + lane, klast)) : 0;
    double2 z_nxt = any ? ldg2(p_uv + min(k0 + lane, klast)) : make_double2(0.0, 0.0);
    d4 l_nxt = ldg256(lm4 + 4 * (size_t)j_nxt);
    int j_nxt2 = any ? __ldg(p_lm + min(k0 + 32 + lane, klast)) : 0;
    for (int base = k0; base < k1; base += 32) {
        const int k = base + lane;
        const double2 z = z_nxt;
        const double2 l01 = make_double2(l_nxt.x, l_nxt.y);
        const double lz = l_nxt.z;
        {   // prefetch tile t+1 (clamped indices keep the loads in range; values of inactive lanes are not used)
            z_nxt = ldg2(p_uv + min(k + 32, klast));
            l_nxt = ldg256(lm4 + 4 * (size_t)j_nxt2);
            j_nxt2 = __ldg(p_lm + min(k + 64, klast));
        }
        double r0[D], r1[D];   // the two whitened rows [J_q (N) | J_K (6) | b] of this lane's factor (zero if none)
#pragma unroll
        for (int i = 0; i < D; ++i) { r0[i] = 0.0; r1[i] = 0.0; }
        if (k < k1) {
            ProjOut o;
            project_factor<true, true>(cm, s_cam, s_cam + 9, l01.x, l01.y, lz, z.x, z.y, o);
            if (ROLE == 0) ncheir += o.cheir ? 1u : 0u;
#pragma unroll
            for (int i = 0; i < N; ++i) {   // column i of getLinearJacobian: a_i x (l - o_i)   (:240)
                const double* a = s_chain + 6 * i;
                const double rx = l01.x - a[3], ry = l01.y - a[4], rz = lz - a[5];
                const double cx = a[1] * rz - a[2] * ry, cy = a[2] * rx - a[0] * rz, cz = a[0] * ry - a[1] * rx;
                r0[i] = -(o.jl[0] * cx + o.jl[1] * cy + o.jl[2] * cz);       // J1 = -(Dpoint * Jrobot)   (:245)
                r1[i] = -(o.jl[3] * cx + o.jl[4] * cy + o.jl[5] * cz);
            }
            if (kStore) {   // packed planes [jq row0 | jq row1 | jl row0 | jl row1]
                double flat[2 * N + 6];
#pragma unroll
                for (int i = 0; i < N; ++i) { flat[i] = r0[i]; flat[N + i] = r1[i]; }
#pragma unroll
                for (int i = 0; i < 6; ++i) flat[2 * N + i] = o.jl[i];
#pragma unroll
                for (int i = 0; i < N + 3; ++i) __stcs(J + (size_t)i * M + k, make_double2(flat[2 * i], flat[2 * i + 1]));
                if (STORE_B) b_out[k] = make_double2(o.b0, o.b1);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) { r0[N + i] = o.jk[i]; r1[N + i] = o.jk[6 + i]; }
            r0[N + 6] = o.b0; r1[N + 6] = o.b1;
        }
#pragma unroll
        for (int i = 0; i < CNT; ++i) {   // indices are compile-time after unrolling: the table lookups fold to register names
            const int e = E0 + i;
            acc[i] = fma(r0[tri.row[e]], r0[tri.col[e]], fma(r1[tri.row[e]], r1[tri.col[e]], acc[i]));
        }
    }
    // warp totals: lane l ends up with entries E0 + l + 32 t
    double tot[Cf::NB];
#pragma unroll
    for (int t = 0; t < Cf::NB; ++t) {
        double v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (32 * t + i < CNT) ? acc[(32 * t + i < CNT) ? 32 * t + i : 0] : 0.0;
        tot[t] = transpose_reduce32(v, lane);
    }
    double* B = Bpp + (size_t)p * N * N;
    double* BK = BpK + (size_t)p * N * 6;
    double* g = gp + (size_t)p * N;
#pragma unroll
    for (int t = 0; t < Cf::NB; ++t) {
        const int e = E0 + 32 * t + lane;
        if (32 * t + lane >= CNT) continue;
        const int lo = tri_row<D>(e), hi = tri_col<D>(e);   // lo < N for every entry of this kernel
        double v = tot[t];
        if (hi < N) {
            if (lo == hi && with_enc) v += pc.inv_sigma_enc[lo] * pc.inv_sigma_enc[lo];
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (lo == hi && dt.has[k]) v += pc.inv_sigma_link[k][lo] * pc.inv_sigma_link[k][lo] * (double)((dt.has[k][p] ? 1 : 0) + (dt.has[k][p + 1] ? 1 : 0));
            B[lo * N + hi] = v; B[hi * N + lo] = v;
        } else if (hi < N + 6) BK[lo * 6 + (hi - N)] = v;
        else {                                               // hi == N+6: J_q^T b, plus the encoder factor's -J^T r (J = I)
            if (with_enc) v -= encoder_residual(pc, q[(size_t)p * N + lo], enc[(size_t)p * N + lo], lo) * pc.inv_sigma_enc[lo];
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (dt.has[k]) {                             // links: -J1^T r (this keyframe is q1) and -J0^T r (it is q0 of the next link)
                    if (dt.has[k][p]) v -= drift_residual(pc, dt, k, q, N, p, lo) * pc.inv_sigma_link[k][lo];
                    if (dt.has[k][p + 1]) v += drift_residual(pc, dt, k, q, N, p + 1, lo) * pc.inv_sigma_link[k][lo];
                }
            g[lo] = v;
        }
    }
    if (ROLE == 0) {
        ncheir = __reduce_add_sync(kFull, ncheir);
        if (lane == 0 && ncheir) atomicAdd(cheir_count, (unsigned long long)ncheir);   // statistics only
    }
}

template <int N, int ROLE, bool STORE_B>
struct LinRoleDispatch {
    template <typename... A>
    static __device__ __forceinline__ void run(int role, A&&... a) {
        if (role == ROLE) linearize_role<N, ROLE, STORE_B>(a...);
        else if constexpr (ROLE + 1 < LinCfg<N>::W) LinRoleDispatch<N, ROLE + 1, STORE_B>::run(role, a...);
    }
};

// cta_part row of this kernel: column 27 = the CTA's encoder-factor error, all other columns zero (the projection part
// of B_KK / g_K / error comes from proj_kk_kernel, in rows of the same layout)
template <int N, bool STORE_B>
__global__ void __launch_bounds__(LinCfg<N>::kThreads, 2)
linearize_pose_kernel(Camera cm, PriorConst pc, int P, int64_t M, const int* __restrict__ pose_ptr, const int* __restrict__ p_lm,
                      const double2* __restrict__ p_uv, const double* __restrict__ cam, const double* __restrict__ chain,
                      const double* __restrict__ lm4, const double* __restrict__ q, const double* __restrict__ enc,
                      const uint8_t* __restrict__ has_enc, int add_priors, double2* __restrict__ J, double* __restrict__ Bpp,
                      double* __restrict__ BpK, double* __restrict__ gp, double* __restrict__ cta_part, double2* __restrict__ b_out,
                      unsigned long long* __restrict__ cheir_count, DriftTab dt) {
    typedef LinCfg<N> Cf;
    if (!add_priors) { dt.has[0] = nullptr; dt.has[1] = nullptr; }   // like the encoder factors, link factors are added by rank 0 only
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kf = warp / Cf::W, role = warp % Cf::W;
    double* s_cam = smem + (size_t)warp * Cf::kPerWarpDoubles;   // 12
    double* s_chain = s_cam + 12;                                 // 6N
    double* s_enc = smem + (size_t)Cf::kWarps * Cf::kPerWarpDoubles;   // [KF] encoder-factor error of each keyframe
    const int p = blockIdx.x * Cf::KF + kf;
    if (role == 0 && lane == 0) s_enc[kf] = 0.0;
    if (p < P) {
        for (int i = lane; i < 12; i += 32) s_cam[i] = cam[(size_t)p * 12 + i];
        for (int i = lane; i < 6 * N; i += 32) s_chain[i] = chain[(size_t)p * 6 * N + i];
        __syncwarp();
        const bool with_enc = add_priors && (has_enc == nullptr || has_enc[p]);
        LinRoleDispatch<N, 0, STORE_B>::run(role, cm, pc, p, M, pose_ptr, p_lm, p_uv, lm4, q, enc, with_enc, dt, J, Bpp, BpK, gp, b_out, cheir_count,
                                             s_cam, s_chain, lane);
        if (role == 0 && (with_enc || dt.any())) {   // EncoderFactor / link-factor error 0.5 |r / sigma|^2 (a link is counted at its q1)
            double e2 = 0.0;
            if (lane < N) {
                if (with_enc) { const double wr = encoder_residual(pc, q[(size_t)p * N + lane], enc[(size_t)p * N + lane], lane); e2 = 0.5 * wr * wr; }
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (dt.has[k] && dt.has[k][p]) { const double wd = drift_residual(pc, dt, k, q, N, p, lane); e2 += 0.5 * wd * wd; }
            }
            e2 = warp_sum(e2);
            if (lane == 0) s_enc[kf] = e2;
        }
    }
    __syncthreads();
    // fixed-order combine of the CTA's keyframes -> one partial row per CTA
    if (threadIdx.x < kPartCols) {
        double s = 0.0;
        if (threadIdx.x == 27)
#pragma unroll
            for (int w = 0; w < Cf::KF; ++w) s += s_enc[w];
        cta_part[(size_t)blockIdx.x * kPartCols + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------ K1b: extrinsic block of the normal equations
// B_KK = sum J_K^T J_K (21 unique), g_K = sum J_K^T b (6) and the projection error sum 0.5 |b|^2 (1) are sums over
// ALL projection factors.  Flat grid-stride loop over the pose-major observation list (neighbouring lanes share the
// camera row, so its gather is an L1 broadcast); every thread keeps the 28 sums of its own factors in registers; one
// transpose-reduce per warp and a fixed-order combine per CTA at the end.  Row layout = linearize_finish_kernel's:
// [B_KK upper triangle row-major (21) | g_K (6) | error (1) | 0 ...].
constexpr int kKkThreads = 256;
__global__ void __launch_bounds__(kKkThreads)
proj_kk_kernel(Camera cm, int64_t M, const int* __restrict__ p_pose, const int* __restrict__ p_lm, const double2* __restrict__ p_uv,
               const double* __restrict__ cam, const double* __restrict__ lm4, double* __restrict__ blk_part) {
    constexpr TriTable<7> tri = TriTable<7>();
    double v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kKkThreads;
    for (int64_t k = (int64_t)blockIdx.x * kKkThreads + threadIdx.x; k < M; k += stride) {
        const int p = __ldg(p_pose + k), j = __ldg(p_lm + k);
        const double2 z = ldg2(p_uv + k);
        double R[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) { const d4 t = ldg256(cam + (size_t)p * 12 + 4 * i); R[4 * i] = t.x; R[4 * i + 1] = t.y; R[4 * i + 2] = t.z; R[4 * i + 3] = t.w; }
        const d4 lq = ldg256(lm4 + 4 * (size_t)j);
        ProjOut o;
        project_factor<true, true>(cm, R, R + 9, lq.x, lq.y, lq.z, z.x, z.y, o);
        double r0[7], r1[7];
#pragma unroll
        for (int i = 0; i < 6; ++i) { r0[i] = o.jk[i]; r1[i] = o.jk[6 + i]; }
        r0[6] = o.b0; r1[6] = o.b1;
#pragma unroll
        for (int e = 0; e < 28; ++e) v[e] = fma(r0[tri.row[e]], r0[tri.col[e]], fma(r1[tri.row[e]], r1[tri.col[e]], v[e]));
    }
    // TriTable<7> order: row a < 6 holds (a, a..5) then (a, 6); remap to [B_KK (21) | g_K (6) | err] while reducing
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = 0.0;
    {
        int nk = 0;
#pragma unroll
        for (int e = 0; e < 28; ++e) {
            const int a = tri.row[e], b = tri.col[e];
            if (b < 6) { w[nk] = v[e]; ++nk; }
            else if (a < 6) w[21 + a] = v[e];
            else w[27] = 0.5 * v[e];
        }
    }
    const double tot = transpose_reduce32(w, lane);   // lane l: warp total of column l
    __shared__ double s[kKkThreads / 32][32];
    s[warp][lane] = tot;
    __syncthreads();
    if (threadIdx.x < kPartCols) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kKkThreads / 32; ++i) t += s[i][threadIdx.x];
        blk_part[(size_t)blockIdx.x * kPartCols + threadIdx.x] = t;
    }
}

// ------------------------------------------------------------------ deterministic row reduction
// out[b][c] = sum over the rows of chunk b of in[r][c]; C <= 32 columns, fixed order.
__global__ void __launch_bounds__(256) reduce_rows_kernel(const double* __restrict__ in, int64_t R, int C, int ld_in, int rows_per_block,
                                                          double* __restrict__ out) {
    __shared__ double s[8][kPartCols];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(R, r0 + rows_per_block);
    double a = 0.0;
    if (lane < C)
        for (int64_t r = r0 + warp; r < r1; r += 8) a += in[r * ld_in + lane];
    s[warp][lane] = a;
    __syncthreads();
    if (warp == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s[w][lane];
        out[(size_t)blockIdx.x * kPartCols + lane] = lane < C ? t : 0.0;
    }
}

// Two independent row reductions in one launch (block 0: A, block 1: B); each sums all rows of its input, fixed order.
__global__ void __launch_bounds__(1024) reduce_rows_pair_kernel(const double* __restrict__ inA, int RA, int CA, int ldA, double* __restrict__ outA,
                                                                const double* __restrict__ inB, int RB, int CB, int ldB, double* __restrict__ outB,
                                                                const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double s[32][kPartCols];
    const double* in = blockIdx.x == 0 ? inA : inB;
    const int R = blockIdx.x == 0 ? RA : RB, C = blockIdx.x == 0 ? CA : CB, ld = blockIdx.x == 0 ? ldA : ldB;
    double* out = blockIdx.x == 0 ? outA : outB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double a0 = 0.0, a1 = 0.0;
    if (lane < C) {
        int r = warp;
        for (; r + 32 < R; r += 64) { a0 += in[(size_t)r * ld + lane]; a1 += in[(size_t)(r + 32) * ld + lane]; }
        if (r < R) a0 += in[(size_t)r * ld + lane];
    }
    s[warp][lane] = a0 + a1;
    __syncthreads();
    if (warp == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 32; ++w) t += s[w][lane];
        out[lane] = lane < C ? t : 0.0;
    }
}

// ------------------------------------------------------------------ K2a: landmark blocks (landmark-major)
// C_j = sum J_l^T J_l + prior, g_j = sum J_l^T b + prior, W_Kj = sum J_K^T J_l.
// EIGHT lanes per landmark (four landmarks per warp): lane s takes observations s, s+8, ... of the landmark's
// landmark-major segment (coalesced 128-byte reads of l_uv / l_pose), accumulates the 27 block entries of its own
// observations in registers, and a three-step fold over the 8 lanes (fixed tree, 28 shuffles) leaves entries
// 4s..4s+3 on lane s, which writes them.  The camera-side Jacobians are recomputed here from the [P][12] camera
// table (L2-resident) so that nothing per-observation has to be stored or permuted for this pass.
// (The first version ran one thread per landmark: 22 dependent gather round trips per thread, 7 % of the HBM roofline.)
constexpr int kLmLanes = 8;
constexpr int kLmPerBlock = 128 / kLmLanes;

template <int W, int LANEBIT>
__device__ __forceinline__ void fold_step(double (&v)[32], int lane) {
    const bool upper = (lane & LANEBIT) != 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double lo = v[i], hi = v[i + W];
        const double send = upper ? lo : hi, keep = upper ? hi : lo;
        v[i] = keep + __shfl_xor_sync(kFull, send, LANEBIT);
    }
}

__global__ void __launch_bounds__(128)
landmark_blocks_kernel(Camera cm, int L, const int* __restrict__ lm_ptr, const int* __restrict__ l_pose, const double2* __restrict__ l_uv,
                       const double* __restrict__ cam, const double* __restrict__ lm4, const double* __restrict__ lm_prior,
                       double* __restrict__ C, double* __restrict__ gl, double* __restrict__ WK, double* __restrict__ blk_err) {
    const int j = blockIdx.x * kLmPerBlock + (threadIdx.x >> 3);
    const int sub = threadIdx.x & 7, lane = threadIdx.x & 31;
    double v[32];   // [c (6) | g (3) | wk (18) | pad]
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.0;
    double lx = 0.0, ly = 0.0, lz = 0.0;
    if (j < L) {
        const double2 l01 = ldg2(reinterpret_cast<const double2*>(lm4 + 4 * (size_t)j));
        lx = l01.x; ly = l01.y; lz = __ldg(lm4 + 4 * (size_t)j + 2);
        const int k0 = lm_ptr[j], k1 = lm_ptr[j + 1];
        for (int k = k0 + sub; k < k1; k += kLmLanes) {
            const int p = __ldg(l_pose + k);
            const double2 z = ldg2(l_uv + k);
            double R[12];
#pragma unroll
            for (int i = 0; i < 3; ++i) { const d4 t = ldg256(cam + (size_t)p * 12 + 4 * i); R[4 * i] = t.x; R[4 * i + 1] = t.y; R[4 * i + 2] = t.z; R[4 * i + 3] = t.w; }
            ProjOut o;
            project_factor<true, true>(cm, R, R + 9, lx, ly, lz, z.x, z.y, o);
            v[0] += o.jl[0] * o.jl[0] + o.jl[3] * o.jl[3]; v[1] += o.jl[0] * o.jl[1] + o.jl[3] * o.jl[4];
            v[2] += o.jl[0] * o.jl[2] + o.jl[3] * o.jl[5]; v[3] += o.jl[1] * o.jl[1] + o.jl[4] * o.jl[4];
            v[4] += o.jl[1] * o.jl[2] + o.jl[4] * o.jl[5]; v[5] += o.jl[2] * o.jl[2] + o.jl[5] * o.jl[5];
#pragma unroll
            for (int a = 0; a < 3; ++a) v[6 + a] += o.jl[a] * o.b0 + o.jl[3 + a] * o.b1;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int a = 0; a < 3; ++a) v[9 + 3 * r + a] += o.jk[r] * o.jl[a] + o.jk[6 + r] * o.jl[3 + a];
        }
    }
    fold_step<16, 4>(v, lane); fold_step<8, 2>(v, lane); fold_step<4, 1>(v, lane);   // lane sub now holds entries 4 sub .. 4 sub + 3
    double err = 0.0;
    if (j < L) {
        // PriorFactor<Point3> with Robust(Cauchy k, Diagonal sigma_lm)  (ArmSlamCalib.cpp:705, :1648-1651)
        const double r0 = (lx - lm_prior[3 * (size_t)j]) * cm.inv_sigma_lm, r1 = (ly - lm_prior[3 * (size_t)j + 1]) * cm.inv_sigma_lm,
                     r2 = (lz - lm_prior[3 * (size_t)j + 2]) * cm.inv_sigma_lm;
        const double n2 = r0 * r0 + r1 * r1 + r2 * r2, w = cm.k2 / (cm.k2 + n2);
        const double d = w * cm.inv_sigma_lm * cm.inv_sigma_lm;
        if (sub == 0) err = 0.5 * w * n2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = 4 * sub + i;
            double t = v[i];
            if (e == 0 || e == 3 || e == 5) t += d;
            if (e == 6) t -= w * r0 * cm.inv_sigma_lm;
            if (e == 7) t -= w * r1 * cm.inv_sigma_lm;
            if (e == 8) t -= w * r2 * cm.inv_sigma_lm;
            if (e < 6) C[6 * (size_t)j + e] = t;
            else if (e < 9) gl[3 * (size_t)j + (e - 6)] = t;
            else if (e < 27) WK[18 * (size_t)j + (e - 9)] = t;
        }
    }
    // block partial of the prior error (fixed tree)
    __shared__ double s[4];
    err = warp_sum(err);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = err;
    __syncthreads();
    if (threadIdx.x == 0) blk_err[blockIdx.x] = s[0] + s[1] + s[2] + s[3];
}

// closed-form inverse of the symmetric 3x3 (xx,xy,xz,yy,yz,zz); false if not positive definite
__device__ __forceinline__ bool spd3_inv(const double* c, double lambda, double* inv) {
    const double a = c[0] + lambda, b = c[1], d = c[2], e = c[3] + lambda, f = c[4], g = c[5] + lambda;
    const double d1 = e - b * b / a;
    const double l21 = (f - d * b / a) / d1;
    const double d2 = g - d * d / a - l21 * l21 * d1;
    const double c00 = e * g - f * f, c01 = d * f - b * g, c02 = b * f - d * e;
    const double id = 1.0 / (a * c00 + b * c01 + d * c02);
    inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
    inv[3] = (a * g - d * d) * id; inv[4] = (b * d - a * f) * id; inv[5] = (a * e - b * b) * id;
    return (a > 0) && (d1 > 0) && (d2 > 0);
}

__device__ __forceinline__ void sym3_mul(const double* s, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = s[0] * x + s[1] * y + s[2] * z; oy = s[1] * x + s[3] * y + s[4] * z; oz = s[2] * x + s[4] * y + s[5] * z;
}

// ------------------------------------------------------------------ K2a': damp + invert landmark blocks
// Cinv_j = (C_j + lambda I)^-1, cg_j = Cinv_j g_j, and the per-block partials of
// S_KK -= W_Kj Cinv_j W_Kj^T (21 sym) and rhs_K -= W_Kj cg_j (6).   One thread per landmark.
__global__ void __launch_bounds__(128)
landmark_invert_kernel(int L, double lambda, const double* __restrict__ C, const double* __restrict__ gl, const double* __restrict__ WK,
                       double* __restrict__ Cinv, double* __restrict__ u4, double* __restrict__ blk_part, int* __restrict__ fail_idx) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double part[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) part[i] = 0.0;
    if (j < L) {
        double ci[6];
        if (!spd3_inv(C + 6 * (size_t)j, lambda, ci)) atomicMin(fail_idx, j);
#pragma unroll
        for (int i = 0; i < 6; ++i) Cinv[6 * (size_t)j + i] = ci[i];
        double gx, gy, gz;
        sym3_mul(ci, gl[3 * (size_t)j], gl[3 * (size_t)j + 1], gl[3 * (size_t)j + 2], gx, gy, gz);
        u4[4 * (size_t)j] = gx; u4[4 * (size_t)j + 1] = gy; u4[4 * (size_t)j + 2] = gz; u4[4 * (size_t)j + 3] = 0.0;
        double W[18], Y[18];
#pragma unroll
        for (int i = 0; i < 18; ++i) W[i] = WK[18 * (size_t)j + i];
#pragma unroll
        for (int r = 0; r < 6; ++r) sym3_mul(ci, W[3 * r], W[3 * r + 1], W[3 * r + 2], Y[3 * r], Y[3 * r + 1], Y[3 * r + 2]);
        int idx = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = r; c < 6; ++c) part[idx++] = Y[3 * r] * W[3 * c] + Y[3 * r + 1] * W[3 * c + 1] + Y[3 * r + 2] * W[3 * c + 2];
#pragma unroll
        for (int r = 0; r < 6; ++r) part[21 + r] = W[3 * r] * gx + W[3 * r + 1] * gy + W[3 * r + 2] * gz;
    }
    __shared__ double s[4][kPartCols];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 27; ++i) {
        const double t = warp_sum(part[i]);
        if (lane == 0) s[warp][i] = t;
    }
    __syncthreads();
    if (threadIdx.x < kPartCols) blk_part[(size_t)blockIdx.x * kPartCols + threadIdx.x] = threadIdx.x < 27 ? s[0][threadIdx.x] + s[1][threadIdx.x] + s[2][threadIdx.x] + s[3][threadIdx.x] : 0.0;
}

// ------------------------------------------------------------------ load one observation's Jacobian planes
template <int N>
__device__ __forceinline__ void load_planes(const double2* __restrict__ J, int64_t M, int k, double (&f)[2 * N + 6]) {
#pragma unroll
    for (int i = 0; i < N + 3; ++i) {
        const double2 t = __ldcs(J + (size_t)i * M + k);   // read once per pass: stream through, keep L2 for the gathered tables
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}

// ------------------------------------------------------------------ Schur-diagonal blocks + reduced rhs (pose-major)
// Per keyframe: Dloc_p = - sum_m J_q^T (J_l Cinv_j J_l^T) J_q  and  rloc_p = - sum_m J_q^T J_l (Cinv_j g_j)
// (the parts of S_pp and of the reduced right-hand side that come from this rank's observations).
template <int N>
struct PrecSmem {
    static constexpr int LD = GramCfg<N>::LD;
    static constexpr int kPerWarpDoubles = 2 * 2 * 32 * LD + 64;
    static constexpr size_t kBytes = sizeof(double) * kPerWarpDoubles * kWarpsPerCta;
};

template <int N>
__global__ void __launch_bounds__(kCtaThreads)
pose_schur_diag_kernel(int P, int64_t M, const int* __restrict__ pose_ptr, const int* __restrict__ p_lm, const double2* __restrict__ J,
                       const double* __restrict__ Cinv, const double* __restrict__ u4, double* __restrict__ Dloc, double* __restrict__ rloc) {
    typedef GramCfg<N> Cf;
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWarpsPerCta + warp;
    if (p >= P) return;
    double* s_U = smem + (size_t)warp * PrecSmem<N>::kPerWarpDoubles;   // Q J_q  [2][32][LD]
    double* s_V = s_U + 2 * 32 * Cf::LD;                                // J_q    [2][32][LD]
    double* s_e = s_V + 2 * 32 * Cf::LD;                                // e = J_l cg  [32][2]
    double acc[Cf::H];
#pragma unroll
    for (int t = 0; t < Cf::H; ++t) acc[t] = 0.0;
    double racc = 0.0;
    const int k0 = pose_ptr[p], k1 = pose_ptr[p + 1];
    for (int base = k0; base < k1; base += 32) {
        const int k = base + lane, count = min(32, k1 - base);
        if (k < k1) {
            double f[2 * N + 6];
            load_planes<N>(J, M, k, f);
            const int j = __ldg(p_lm + k);
            const double* ci = Cinv + 6 * (size_t)j;
            const double* jl = f + 2 * N;
            double t0x, t0y, t0z, t1x, t1y, t1z;
            sym3_mul(ci, jl[0], jl[1], jl[2], t0x, t0y, t0z);
            sym3_mul(ci, jl[3], jl[4], jl[5], t1x, t1y, t1z);
            const double q00 = jl[0] * t0x + jl[1] * t0y + jl[2] * t0z, q01 = jl[0] * t1x + jl[1] * t1y + jl[2] * t1z;
            const double q11 = jl[3] * t1x + jl[4] * t1y + jl[5] * t1z;
            const double2 c01 = ldg2(reinterpret_cast<const double2*>(u4 + 4 * (size_t)j));
            const double cz = __ldg(u4 + 4 * (size_t)j + 2);
            s_e[2 * lane] = jl[0] * c01.x + jl[1] * c01.y + jl[2] * cz;
            s_e[2 * lane + 1] = jl[3] * c01.x + jl[4] * c01.y + jl[5] * cz;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                s_V[(0 * 32 + lane) * Cf::LD + i] = f[i];
                s_V[(1 * 32 + lane) * Cf::LD + i] = f[N + i];
                s_U[(0 * 32 + lane) * Cf::LD + i] = q00 * f[i] + q01 * f[N + i];
                s_U[(1 * 32 + lane) * Cf::LD + i] = q01 * f[i] + q11 * f[N + i];
            }
        }
        __syncwarp();
        gram_tile<N, false>(s_U, s_V, count, lane, acc);
        {   // rhs: lanes of group g take row g (or both rows when one group)
            const int g = lane / Cf::GS, c = lane % Cf::GS;
            if (c < N)
                for (int o = 0; o < count; ++o) {
#pragma unroll
                    for (int rr = 0; rr < 2 / Cf::G; ++rr) {
                        const int r = Cf::G == 2 ? g : rr;
                        racc = fma(s_V[(r * 32 + o) * Cf::LD + c], s_e[2 * o + r], racc);
                    }
                }
        }
        __syncwarp();
    }
    gram_finish<N>(acc);
    if (Cf::G == 2) racc += __shfl_xor_sync(kFull, racc, 16);
    if (lane < N) {
        const int c = lane;
        double* Dp = Dloc + (size_t)p * N * N;
#pragma unroll
        for (int t = 0; t < Cf::H; ++t) {
            int c2 = c + t;
            if (c2 >= N) c2 -= N;
            Dp[c * N + c2] = -acc[t]; Dp[c2 * N + c] = -acc[t];
        }
        rloc[(size_t)p * N + c] = -racc;
    }
}

// Minv_p = (B_pp + lambda I + Dsum_p)^-1 by in-shared-memory Gauss-Jordan (SPD, no pivoting); one warp per keyframe.
// rhs_p = g_p + rsum_p.  Dsum/rsum are the (all-reduced) outputs of pose_schur_diag_kernel.
template <int N>
__global__ void __launch_bounds__(kCtaThreads)
pose_invert_kernel(int P, double lambda, const double* __restrict__ Bpp, const double* __restrict__ gp, const double* __restrict__ Dsum,
                   const double* __restrict__ rsum, double* __restrict__ Minv, double* __restrict__ rhs, int* __restrict__ fail_idx) {
    __shared__ double sA[kWarpsPerCta][N][2 * N + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWarpsPerCta + warp;
    if (p >= P) return;
    double(*A)[2 * N + 1] = sA[warp];
    for (int i = lane; i < N * N; i += 32) {
        const int r = i / N, c = i % N;
        A[r][c] = Bpp[(size_t)p * N * N + i] + Dsum[(size_t)p * N * N + i] + (r == c ? lambda : 0.0);
        A[r][N + c] = r == c ? 1.0 : 0.0;
    }
    if (lane < N) rhs[(size_t)p * N + lane] = gp[(size_t)p * N + lane] + rsum[(size_t)p * N + lane];
    __syncwarp();
    bool ok = true;
    for (int k = 0; k < N; ++k) {
        const double piv = A[k][k];
        ok = ok && (piv > 0);
        const double ip = 1.0 / piv;
        __syncwarp();
        for (int c = lane; c < 2 * N; c += 32) A[k][c] *= ip;
        __syncwarp();
        for (int i = lane; i < N * 2 * N; i += 32) {
            const int r = i / (2 * N), c = i % (2 * N);
            if (r != k && c != k) A[r][c] -= A[r][k] * A[k][c];
        }
        __syncwarp();
        for (int r = lane; r < N; r += 32) if (r != k) A[r][k] = 0.0;
        __syncwarp();
    }
    for (int i = lane; i < N * N; i += 32) Minv[(size_t)p * N * N + i] = A[i / N][N + i % N];
    if (!ok && lane == 0) atomicMin(fail_idx, p);
}

// ------------------------------------------------------------------ cluster-Jacobi preconditioner
// Preconditioner blocks are the diagonal blocks of S over "clusters" = runs of consecutive, co-visible keyframes
// (built on the host).  With per-keyframe blocks the intra-cluster coupling through shared landmarks is what
// makes PCG crawl on arm trajectories (hundreds of keyframes looking at the same points).
//
// W_k = J_q^T J_l (N x 3, row-major) per observation, pose-major AoS [M][3N]
template <int N>
__global__ void __launch_bounds__(256) obs_w_kernel(int64_t M, const double2* __restrict__ J, double* __restrict__ Wm) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    double f[2 * N + 6];
    load_planes<N>(J, M, (int)k, f);
    const double* jl = f + 2 * N;
    double* o = Wm + (size_t)k * 3 * N;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) o[3 * a + c] = f[a] * jl[c] + f[N + a] * jl[3 + c];
}

// Row block of keyframe p inside its cluster block of S:
//   S[p][p'] = [p == p'] (B_pp + lambda I) - sum over landmarks j seen by both of W_m Cinv_j W_m'^T
// and the keyframe's part of the reduced right-hand side, rloc_p = - sum_m W_m (Cinv_j g_j).
// One warp per keyframe; the N x n row block lives in shared memory; observations are visited in a fixed order.
// run_lo/run_hi[k]: the landmark-major range holding the same landmark's observations from this cluster.
template <int N>
__global__ void __launch_bounds__(kCtaThreads)
cluster_schur_kernel(int P, double lambda, int add_diag, int nmax, const int* __restrict__ pose_ptr, const int* __restrict__ p_lm,
                     const int* __restrict__ run_lo, const int* __restrict__ run_hi, const int* __restrict__ l_pose,
                     const int* __restrict__ l_perm, const double* __restrict__ Wm, const double* __restrict__ Cinv,
                     const double* __restrict__ cg4, const double* __restrict__ Bpp, const int* __restrict__ pose_cluster,
                     const int* __restrict__ cl_ptr, const long long* __restrict__ cl_off, double* __restrict__ Sc, double* __restrict__ rloc) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWarpsPerCta + warp;
    if (p >= P) return;
    double* T = smem + (size_t)warp * (N * nmax + 3 * N);
    double* Y = T + N * nmax;
    const int c = pose_cluster[p], c0 = cl_ptr[c], n = (cl_ptr[c + 1] - c0) * N;
    for (int i = lane; i < N * n; i += 32) T[i] = 0.0;
    __syncwarp();
    if (add_diag)
        for (int e = lane; e < N * N; e += 32) {
            const int a = e / N, b = e % N;
            T[a * n + (p - c0) * N + b] = Bpp[(size_t)p * N * N + e] + (a == b ? lambda : 0.0);
        }
    __syncwarp();
    double racc = 0.0;   // lane a < N
    const int k0 = pose_ptr[p], k1 = pose_ptr[p + 1];
    for (int k = k0; k < k1; ++k) {
        const int j = __ldg(p_lm + k);
        const double* ci = Cinv + 6 * (size_t)j;
        const double* W = Wm + (size_t)k * 3 * N;
        for (int i = lane; i < 3 * N; i += 32) {
            const int a = i / 3, cc = i % 3;
            const double w0 = W[3 * a], w1 = W[3 * a + 1], w2 = W[3 * a + 2];
            const double c0_ = cc == 0 ? ci[0] : (cc == 1 ? ci[1] : ci[2]);
            const double c1_ = cc == 0 ? ci[1] : (cc == 1 ? ci[3] : ci[4]);
            const double c2_ = cc == 0 ? ci[2] : (cc == 1 ? ci[4] : ci[5]);
            Y[i] = w0 * c0_ + w1 * c1_ + w2 * c2_;
        }
        if (lane < N) {
            const double* g = cg4 + 4 * (size_t)j;
            racc += W[3 * lane] * g[0] + W[3 * lane + 1] * g[1] + W[3 * lane + 2] * g[2];
        }
        __syncwarp();
        const int lo = run_lo[k], hi = run_hi[k];
        for (int kk = lo; kk < hi; ++kk) {
            const int pp = __ldg(l_pose + kk);
            const double* W2 = Wm + (size_t)__ldg(l_perm + kk) * 3 * N;
            for (int e = lane; e < N * N; e += 32) {
                const int a = e / N, b = e % N;
                const double v = Y[3 * a] * W2[3 * b] + Y[3 * a + 1] * W2[3 * b + 1] + Y[3 * a + 2] * W2[3 * b + 2];
                T[a * n + (pp - c0) * N + b] -= v;
            }
        }
        __syncwarp();
    }
    double* out = Sc + cl_off[c] + (size_t)(p - c0) * N * n;
    for (int i = lane; i < N * n; i += 32) out[i] = T[i];
    if (lane < N) rloc[(size_t)p * N + lane] = -racc;
}

// Cluster-major variant of the same computation: one CTA per cluster walks the cluster's (landmark, cluster) incidences
// in a fixed order and accumulates the whole n x n block (and the cluster's share of the reduced right-hand side) in shared
// memory.  Inside one incidence every S entry is touched by exactly one thread (the host checks that a landmark is seen at
// most once per keyframe), so the result is deterministic.  ~10x faster than the keyframe-major kernel above, whose
// pair loop is a chain of dependent gathers.
template <int N>
__global__ void __launch_bounds__(256)
cluster_schur_cta_kernel(double lambda, int add_diag, int gmax, const int* __restrict__ cl_ptr, const long long* __restrict__ cl_off,
                         const int* __restrict__ cl_inc_ptr, const int* __restrict__ cl_inc, const int* __restrict__ inc_lo,
                         const int* __restrict__ inc_hi, const int* __restrict__ inc_lm, const int* __restrict__ l_pose,
                         const int* __restrict__ l_perm, const double* __restrict__ Wm, const double* __restrict__ Cinv,
                         const double* __restrict__ cg4, const double* __restrict__ Bpp, double* __restrict__ Sc, double* __restrict__ rloc) {
    extern __shared__ double sm[];
    const int c = blockIdx.x, c0 = cl_ptr[c], npose = cl_ptr[c + 1] - c0, n = npose * N, ld = n | 1;
    double* S = sm;                         // [n][ld]
    double* rl = S + (size_t)n * ld;        // [n]
    double* Wb = rl + n;                    // [gmax][3N]
    double* Yb = Wb + gmax * 3 * N;         // [gmax][3N]
    int* pp = reinterpret_cast<int*>(Yb + gmax * 3 * N);   // [gmax]
    const int tid = threadIdx.x;
    for (int i = tid; i < n * ld; i += 256) S[i] = 0.0;
    for (int i = tid; i < n; i += 256) rl[i] = 0.0;
    __syncthreads();
    if (add_diag)
        for (int i = tid; i < npose * N * N; i += 256) {
            const int p = i / (N * N), a = (i / N) % N, b = i % N;
            S[(p * N + a) * ld + p * N + b] = Bpp[(size_t)(c0 + p) * N * N + a * N + b] + (a == b ? lambda : 0.0);
        }
    __syncthreads();
    for (int q = cl_inc_ptr[c]; q < cl_inc_ptr[c + 1]; ++q) {
        const int inc = cl_inc[q];
        const int lo = inc_lo[inc], r = inc_hi[inc] - lo, j = inc_lm[inc];
        for (int t = tid; t < r * 3 * N; t += 256) Wb[t] = Wm[(size_t)__ldg(l_perm + lo + t / (3 * N)) * 3 * N + t % (3 * N)];
        if (tid < r) pp[tid] = __ldg(l_pose + lo + tid) - c0;
        __syncthreads();
        const double* ci = Cinv + 6 * (size_t)j;
        for (int t = tid; t < r * 3 * N; t += 256) {
            const int base = t - t % 3, cc = t % 3;
            const double c0_ = cc == 0 ? ci[0] : (cc == 1 ? ci[1] : ci[2]);
            const double c1_ = cc == 0 ? ci[1] : (cc == 1 ? ci[3] : ci[4]);
            const double c2_ = cc == 0 ? ci[2] : (cc == 1 ? ci[4] : ci[5]);
            Yb[t] = Wb[base] * c0_ + Wb[base + 1] * c1_ + Wb[base + 2] * c2_;
        }
        if (tid < r * N) {   // reduced right-hand side: - W_m (Cinv g)_j
            const int i = tid / N, a = tid % N;
            const double* g = cg4 + 4 * (size_t)j;
            rl[pp[i] * N + a] -= Wb[i * 3 * N + 3 * a] * g[0] + Wb[i * 3 * N + 3 * a + 1] * g[1] + Wb[i * 3 * N + 3 * a + 2] * g[2];
        }
        __syncthreads();
        {   // thread = fixed (a, b) entry of the N x N block; thread groups stride over the r*r keyframe pairs
            constexpr int NN = N * N, NGRP = 256 / NN > 0 ? 256 / NN : 1;
            const int ab = tid % NN, grp = tid / NN, a = ab / N, b = ab % N;
            if (grp < NGRP)
                for (int i = 0; i < r; ++i) {
                    const double* y = Yb + i * 3 * N + 3 * a;
                    const double y0 = y[0], y1 = y[1], y2 = y[2];
                    double* Srow = S + (pp[i] * N + a) * ld + b;
                    for (int i2 = grp; i2 < r; i2 += NGRP) {
                        const double* w = Wb + i2 * 3 * N + 3 * b;
                        Srow[pp[i2] * N] -= y0 * w[0] + y1 * w[1] + y2 * w[2];
                    }
                }
        }
        __syncthreads();
    }
    double* out = Sc + cl_off[c];
    for (int i = tid; i < n * n; i += 256) out[i] = S[(i / n) * ld + i % n];
    for (int i = tid; i < n; i += 256) rloc[(size_t)c0 * N + i] = rl[i];
}

// In-place Gauss-Jordan inverse of one cluster block (SPD, no pivoting) in shared memory; one CTA per cluster.
// Also forms the reduced right-hand side of the cluster: rhs = g + rsum.
__global__ void __launch_bounds__(256)
cluster_invert_kernel(int N, const int* __restrict__ cl_ptr, const long long* __restrict__ cl_off, const double* __restrict__ Sc,
                      const double* __restrict__ g, const double* __restrict__ rsum, double* __restrict__ Mc, double* __restrict__ rhs,
                      int* __restrict__ fail_idx) {
    extern __shared__ double A[];
    __shared__ double colk[256];
    __shared__ int bad;
    const int c = blockIdx.x, c0 = cl_ptr[c], n = (cl_ptr[c + 1] - c0) * N, ld = n | 1;
    const double* src = Sc + cl_off[c];
    if (threadIdx.x == 0) bad = 0;
    for (int i = threadIdx.x; i < n * n; i += 256) A[(i / n) * ld + i % n] = src[i];
    for (int i = threadIdx.x; i < n; i += 256) rhs[(size_t)c0 * N + i] = g[(size_t)c0 * N + i] + rsum[(size_t)c0 * N + i];
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const double piv = A[k * ld + k];
        if (threadIdx.x == 0 && !(piv > 0)) bad = 1;
        const double ip = 1.0 / piv;
        for (int i = threadIdx.x; i < n; i += 256) colk[i] = A[i * ld + k];
        __syncthreads();
        for (int cc = threadIdx.x; cc < n; cc += 256) A[k * ld + cc] = cc == k ? ip : A[k * ld + cc] * ip;
        __syncthreads();
        for (int i = threadIdx.x; i < n * n; i += 256) {
            const int r = i / n, cc = i % n;
            if (r != k) A[r * ld + cc] = (cc == k ? 0.0 : A[r * ld + cc]) - colk[r] * A[k * ld + cc];
        }
        __syncthreads();
    }
    double* dst = Mc + cl_off[c];
    for (int i = threadIdx.x; i < n * n; i += 256) dst[i] = A[(i / n) * ld + i % n];
    if (threadIdx.x == 0 && bad) atomicMin(fail_idx, c0);
}

// PCG vector update with the cluster preconditioner; one CTA (128 threads, n <= 128) per cluster:
//   init: x = 0, r = rhs, z = M r, p = z          else: x += alpha p, r -= alpha y, z = M r
// part[c][0] = r_c . z_c
__global__ void __launch_bounds__(128)
pcg_cluster_update_kernel(int init, int N, const int* __restrict__ cl_ptr, const long long* __restrict__ cl_off, const double* __restrict__ Mc,
                          const double* __restrict__ rhs, const double* __restrict__ y, double* __restrict__ x, double* __restrict__ r,
                          double* __restrict__ z, double* __restrict__ p, const double* __restrict__ scal, const int* __restrict__ done,
                          double* __restrict__ part, double* __restrict__ rc, PeerComm pcm) {
    if (!init && *done) return;
    __shared__ double rs[128];
    __shared__ double red[4];
    const int c = blockIdx.x, t = threadIdx.x;
    const size_t base = (size_t)cl_ptr[c] * N;
    const int n = (cl_ptr[c + 1] - cl_ptr[c]) * N;
    double rm = 0.0;
    if (t < n) {
        const size_t i = base + t;
        if (init == 1) { rm = rhs[i]; x[i] = 0.0; }
        else if (init == 2) { rm = rhs[i] - y[i]; }   // warm start: x holds the previous solution, y = S x
        else {   // multi-GPU: y = sum over ranks of the partial products, read from the peers' exchange slots (published before pcg_mid ran)
            const double alpha = scal[0];
            const double yi = pcm.world > 1 ? peer_sum(pcm, *pcm.ctr, (long long)i) : y[i];
            x[i] += alpha * p[i]; rm = r[i] - alpha * yi;
        }
        r[i] = rm;
    }
    rs[t] = rm;
    __syncthreads();
    if (rc && t < N) {   // restriction Z^T r for this cluster
        double a = 0.0;
        for (int i = t; i < n; i += N) a += rs[i];
        rc[(size_t)c * N + t] = a;
    }
    double zm = 0.0;
    if (t < n) {
        const double* Mcol = Mc + cl_off[c] + t;   // column t == row t (symmetric): coalesced across threads
        for (int col = 0; col < n; ++col) zm = fma(Mcol[(size_t)col * n], rs[col], zm);
        z[base + t] = zm;
        if (init) p[base + t] = zm;
    }
    const double d = warp_sum(rm * zm);
    if ((t & 31) == 0) red[t >> 5] = d;
    __syncthreads();
    if (t == 0) part[c] = (red[0] + red[1]) + (red[2] + red[3]);
}

// K rows at PCG start: x_K = 0, r_K = rhs_K, z_K = MinvK r_K, p_K = z_K; scal[7] = r_K . z_K   (one warp)
__global__ void pcg_init_k_kernel(int64_t PN, const double* __restrict__ rhs, const double* __restrict__ MinvK, double* __restrict__ x,
                                  double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ kdot,
                                  double* __restrict__ rcK, const double* __restrict__ yK) {
    const int lane = threadIdx.x;
    const double rk = lane < 6 ? rhs[PN + lane] - (yK ? yK[lane] : 0.0) : 0.0;   // yK != null: warm start, x_K kept
    double zk = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double ri = __shfl_sync(kFull, rk, i);
        if (lane < 6) zk = fma(MinvK[6 * lane + i], ri, zk);
    }
    if (lane < 6) { if (!yK) x[PN + lane] = 0.0; r[PN + lane] = rk; z[PN + lane] = zk; p[PN + lane] = zk; if (rcK) rcK[lane] = rk; }
    const double d = warp_sum(lane < 6 ? rk * zk : 0.0);
    if (lane == 0) *kdot = d;
}

// ------------------------------------------------------------------ coarse level of the two-level preconditioner
// Coarse space Z: one uniform joint-offset vector per cluster (N columns each) plus the extrinsic (6 columns).
// These are the slow modes of arm calibration -- a whole group of keyframes and the landmarks it sees shifting
// together -- which cluster-Jacobi cannot damp.  E = Z^T S Z is assembled block by block in a fixed order, factorised
// and inverted by cuSOLVER (plain library call, once per damping value), and applied additively: M^-1 = M_cl^-1 + Z E^-1 Z^T.
//
// "incidence" = (landmark j, cluster c) with at least one observation; V = sum of W_m over them, T = V Cinv_j.
template <int N>
__global__ void __launch_bounds__(128)
inc_v_kernel(int nInc, const int* __restrict__ inc_lo, const int* __restrict__ inc_hi, const int* __restrict__ inc_lm,
             const int* __restrict__ l_perm, const double* __restrict__ Wm, const double* __restrict__ Cinv, double* __restrict__ V,
             double* __restrict__ T) {
    __shared__ double sv[4][3 * N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inc = blockIdx.x * 4 + warp;
    if (inc >= nInc) return;
    const int lo = inc_lo[inc], hi = inc_hi[inc];
    for (int e = lane; e < 3 * N; e += 32) {
        double a = 0.0;
        for (int kk = lo; kk < hi; ++kk) a += Wm[(size_t)__ldg(l_perm + kk) * 3 * N + e];
        sv[warp][e] = a;
        V[(size_t)inc * 3 * N + e] = a;
    }
    __syncwarp();
    const double* ci = Cinv + 6 * (size_t)inc_lm[inc];
    for (int e = lane; e < 3 * N; e += 32) {
        const int a = e / 3, cc = e % 3;
        const double c0_ = cc == 0 ? ci[0] : (cc == 1 ? ci[1] : ci[2]);
        const double c1_ = cc == 0 ? ci[1] : (cc == 1 ? ci[3] : ci[4]);
        const double c2_ = cc == 0 ? ci[2] : (cc == 1 ? ci[4] : ci[5]);
        T[(size_t)inc * 3 * N + e] = sv[warp][3 * a] * c0_ + sv[warp][3 * a + 1] * c1_ + sv[warp][3 * a + 2] * c2_;
    }
}

__global__ void set_scalar_kernel(double* p, double v) { *p = v; }

// ------------------------------------------------------------------ K3: implicit Schur complement  y = S x
// J_q is never read back in the PCG passes.  Column i of an observation's J_q is -J_l (a_i x (l - o_i))
// (RobotProjectionFactor.h:240-245), so with the keyframe's joint axes a_i / origins o_i (FK table, `chain`) and the
// landmark position l at the linearisation point (`lml`):
//   J_q x_p            = -J_l (A x l - B),            A = sum_i x_i a_i,  B = sum_i x_i (a_i x o_i)      (per keyframe)
//   sum_m J_q^T e_m    :  component i = -(a_i . U - m_i . T),  U = sum_m l_m x t_m,  T = sum_m t_m,  t_m = J_l^T e_m,  m_i = a_i x o_i
// Per observation the passes read the three J_l planes (48 B) and one 32-byte landmark sector instead of all N+3
// planes (16 (N+3) B): at N = 6 the implicit product moves 0.36 GB per PCG iteration instead of 0.74 GB.
template <int N>
__device__ __forceinline__ void load_jl(const double2* __restrict__ J, int64_t M, int k, double (&jl)[6]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double2 t = __ldcs(J + (size_t)(N + i) * M + k);   // read once per pass: stream through
        jl[2 * i] = t.x; jl[2 * i + 1] = t.y;
    }
}

// pass A (pose-major): v_k = J_l^T (J_q x_p)            [optionally first p_p <- z_p + beta p_p]
template <int N, bool UPDATE_P>
__global__ void __launch_bounds__(kCtaThreads)
schur_pass_a_kernel(int P, int64_t M, const int* __restrict__ pose_ptr, const int* __restrict__ p_lm, const double2* __restrict__ J,
                    double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ scal, const int* __restrict__ done,
                    double* __restrict__ v4, const double* __restrict__ zc, const int* __restrict__ pose_agg, const int* __restrict__ p_inv,
                    const double* __restrict__ chain, const double* __restrict__ lml) {
    if (done && *done) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWarpsPerCta + warp;
    if (p >= P) return;
    double ab[6] = {0, 0, 0, 0, 0, 0};   // lane i < N: x_i a_i | x_i (a_i x o_i)
    if (lane < N) {
        double mine = x[(size_t)p * N + lane];
        if (UPDATE_P) {   // p <- z + beta p  (z = cluster part + coarse part Z zc); scal[1] = beta
            double zz = z[(size_t)p * N + lane];
            if (zc) zz += zc[(size_t)pose_agg[p] * N + lane];
            mine = zz + scal[1] * mine;
            x[(size_t)p * N + lane] = mine;
        }
        const double* c = chain + (size_t)p * 6 * N + 6 * lane;
        const double a0 = c[0], a1 = c[1], a2 = c[2], o0 = c[3], o1 = c[4], o2 = c[5];
        ab[0] = mine * a0; ab[1] = mine * a1; ab[2] = mine * a2;
        ab[3] = mine * (a1 * o2 - a2 * o1); ab[4] = mine * (a2 * o0 - a0 * o2); ab[5] = mine * (a0 * o1 - a1 * o0);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) ab[i] = warp_sum(ab[i]);
    const int k0 = pose_ptr[p], k1 = pose_ptr[p + 1];
    // (a one-tile-ahead software pipeline of the streamed loads was measured slower here: 64 instead of 48 registers)
    for (int k = k0 + lane; k < k1; k += 32) {
        double jl[6];
        load_jl<N>(J, M, k, jl);
        const int j = __ldg(p_lm + k);
        const int slot = __ldg(p_inv + k);
        const d4 l = ldg256(lml + 4 * (size_t)j);
        const double w0 = ab[1] * l.z - ab[2] * l.y - ab[3], w1 = ab[2] * l.x - ab[0] * l.z - ab[4], w2 = ab[0] * l.y - ab[1] * l.x - ab[5];
        const double s0 = -(jl[0] * w0 + jl[1] * w1 + jl[2] * w2), s1 = -(jl[3] * w0 + jl[4] * w1 + jl[5] * w2);
        // v goes straight to its landmark-major slot (one full 32-byte sector), so pass B reads its segment contiguously
        stg256(v4 + 4 * (size_t)slot, jl[0] * s0 + jl[3] * s1, jl[1] * s0 + jl[4] * s1, jl[2] * s0 + jl[5] * s1, 0.0);
    }
}

// Sum the rows of part[R][W] (W = 1 or 8) with all 1024 threads of the block: each thread strides over rows with
// four independent accumulators (memory-level parallelism), then a fixed shared-memory tree.  out[W] valid after return.
template <bool CG>
__device__ __forceinline__ double ld_part(const double* p) { return CG ? __ldcg(p) : *p; }

template <int W, int T = kReduceThreads, bool CG = true>   // CG = false: `part` is in shared memory
__device__ __forceinline__ void block_sum_cols(const double* part, int R, double* out, double* scratch) {
    // T threads.  Loads go to L2 (__ldcg): in the fused "last CTA" callers the partials were written by other SMs of the
    // same launch.
    constexpr int RL = T / W;
    const int t = threadIdx.x, col = t % W, rl = t / W;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int r = rl;
    for (; r + 3 * RL < R; r += 4 * RL) {
        a0 += ld_part<CG>(part + (size_t)r * W + col); a1 += ld_part<CG>(part + (size_t)(r + RL) * W + col);
        a2 += ld_part<CG>(part + (size_t)(r + 2 * RL) * W + col); a3 += ld_part<CG>(part + (size_t)(r + 3 * RL) * W + col);
    }
    for (; r < R; r += RL) a0 += ld_part<CG>(part + (size_t)r * W + col);
    scratch[t] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    for (int s = RL / 2; s > 0; s >>= 1) {
        if (rl < s) scratch[t] += scratch[t + s * W];
        __syncthreads();
    }
    if (t < W) out[t] = scratch[t];
    __syncthreads();
}

// ------------------------------------------------------------------ single-block helpers
// sum rows [0,R) of a [R][kPartCols] partial array, fixed order; call with 256 threads. Result valid in all threads for col < 32.
__device__ inline double block_reduce_rows(const double* __restrict__ part, int R, int col_lane, double (*s)[kPartCols]) {
    const int warp = threadIdx.x >> 5;
    double a = 0.0;
    for (int r = warp; r < R; r += 8) a += part[(size_t)r * kPartCols + col_lane];
    __syncthreads();
    s[warp][col_lane] = a;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s[w][col_lane];
    return t;
}

// Multi-GPU, after pass C: reduce this rank's CTA partials of pass C / pass B into the exchange slot's row area, then
// publish the new sequence number to every peer (release at system scope: the slot is complete when a peer sees it).
__global__ void __launch_bounds__(1024)
pcg_signal_kernel(PeerComm pcm, const double* __restrict__ partC, int nC, const double* __restrict__ partB, int nB, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double s[32][kPartCols];
    const unsigned seq = *pcm.ctr + 1u;
    double* rows = pcm.x[pcm.rank] + (size_t)(seq & 1u) * pcm.stride + pcm.row_off;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int which = 0; which < 2; ++which) {
        const double* in = which == 0 ? partC : partB;
        const int R = which == 0 ? nC : nB, C = which == 0 ? 7 : 6;
        double a0 = 0.0, a1 = 0.0;
        if (lane < C) {
            int r = warp;
            for (; r + 32 < R; r += 64) { a0 += in[(size_t)r * kPcgCols + lane]; a1 += in[(size_t)(r + 32) * kPcgCols + lane]; }
            if (r < R) a0 += in[(size_t)r * kPcgCols + lane];
        }
        __syncthreads();
        s[warp][lane] = a0 + a1;
        __syncthreads();
        if (warp == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 32; ++w) t += s[w][lane];
            rows[which * kPartCols + lane] = lane < C ? t : 0.0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *pcm.ctr = seq;
        __threadfence_system();
        for (int r = 0; r < pcm.world; ++r)
            if (r != pcm.rank) st_release_sys_u32(pcm.peer_flags[r] + pcm.rank, seq);
    }
}

// Device-side PCG scalars: scal[0]=alpha, [1]=beta, [2]=rz, [3]=rz0, [4]=pSp, [5]=iterations, [6]=breakdown flag
// Finish y = S p: gathers the K rows, alpha, and updates the K part of x, r, z.  One block of 256 threads.
//   partC: CTA partials of pass C [nC][32] (cols 0..5 = sum B_pK^T p_p, col 6 = sum p_p.y_p)
//   partB: block partials of pass B [nB][32] (cols 0..5 = sum W_Kj u_j)
// Multi-GPU: the partial sums have been all-reduced already and arrive as a single row each (nC = nB = 1).
struct MidArgs {
    int P, N;
    double lambda_val;
    const double* lambda_dev;
    const double *partC, *partB;
    int nC, nB;
    const double *BKK, *MinvK;
    double *p, *x, *r, *z, *yK_out, *scal;
    int* done;
    double* rcK;
};

// T threads of one CTA.  In shared memory: scratch[T], sc/sb, sh.
template <int T>
__device__ __forceinline__ void pcg_mid_body(const MidArgs& a) {
    const double lambda = a.lambda_dev ? *a.lambda_dev : a.lambda_val;
    __shared__ double scratch[T];
    __shared__ double sc[kPcgCols], sb[kPcgCols];
    __shared__ double sh[16];
    const int lane = threadIdx.x & 31, P = a.P, N = a.N;
    block_sum_cols<kPcgCols, T>(a.partC, a.nC, sc, scratch);
    block_sum_cols<kPcgCols, T>(a.partB, a.nB, sb, scratch);
    const double c = lane < kPcgCols ? sc[lane] : 0.0;
    const double b = lane < kPcgCols ? sb[lane] : 0.0;
    double* pK = a.p + (size_t)P * N;
    if (threadIdx.x < 32) {
        double yk = 0.0;
        if (lane < 6) {
            yk = lambda * pK[lane] + c - b;
#pragma unroll
            for (int i = 0; i < 6; ++i) yk += a.BKK[6 * lane + i] * pK[i];
        }
        double dot = lane < 6 ? yk * pK[lane] : 0.0;
        dot = warp_sum(dot);
        const double cdot = __shfl_sync(kFull, c, 6);
        const double pSp = cdot + dot;
        const double alpha = a.scal[2] / pSp;
        if (lane == 0) { a.scal[0] = alpha; a.scal[4] = pSp; if (!(pSp > 0)) { a.scal[6] = 1.0; } }
        double rk = 0.0;
        if (lane < 6) {
            a.yK_out[lane] = yk;
            a.x[(size_t)P * N + lane] += alpha * pK[lane];
            rk = a.r[(size_t)P * N + lane] - alpha * yk;
            a.r[(size_t)P * N + lane] = rk;
            if (a.rcK) a.rcK[lane] = rk;
            sh[lane] = rk;
        }
        __syncwarp();
        if (lane < 6) {
            double zk = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) zk += a.MinvK[6 * lane + i] * sh[i];
            a.z[(size_t)P * N + lane] = zk;
            sh[8 + lane] = zk * rk;
        }
        __syncwarp();
        if (lane == 0) a.scal[7] = sh[8] + sh[9] + sh[10] + sh[11] + sh[12] + sh[13];   // r_K . z_K
    }
}

__global__ void __launch_bounds__(kReduceThreads)
pcg_mid_kernel(MidArgs a, PeerComm pcm, double* __restrict__ rows_local) {
    if (*a.done) return;
    if (pcm.world > 1) {
        // wait until every peer has published this iteration's partial product, then sum the K-row partials of all ranks
        const unsigned seq = *pcm.ctr;
        if (threadIdx.x < pcm.world && threadIdx.x != pcm.rank) {
            const unsigned* f = pcm.peer_flags[pcm.rank] + threadIdx.x;
            const long long t0 = clock64();
            while ((int)(ld_acquire_sys_u32(f) - seq) < 0) {
                if (clock64() - t0 > 8000000000ll) { *pcm.fail = 1; break; }   // ~4 s: a peer died; report instead of hanging
                __nanosleep(64);
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * kPartCols) rows_local[threadIdx.x] = peer_sum(pcm, seq, pcm.row_off + threadIdx.x);
        __syncthreads();
    }
    pcg_mid_body<kReduceThreads>(a);
}

// pass B (landmark-major): u_j = Cinv_j (sum_k v + W_Kj^T x_K) [or cg_j - that, for back-substitution];
// block partials of sum_j W_Kj u_j.   Eight lanes per landmark (as in K2a): lane s sums slots s, s+8, ... of the
// landmark's contiguous v segment (one 32-byte sector per load, 256 B per group and step), lane r < 6 owns row r of
// W_Kj; xor-folds over the group in a fixed order.
constexpr int kPassBThreads = 256;
constexpr int kPassBLmPerBlock = kPassBThreads / kLmLanes;
template <bool BACKSUB>
__global__ void __launch_bounds__(kPassBThreads)
schur_pass_b_kernel(int L, const int* __restrict__ lm_ptr, const double* __restrict__ v4,
                    const double* __restrict__ WK, const double* __restrict__ Cinv, const double* __restrict__ xK,
                    const int* __restrict__ done, const double* __restrict__ cg4, double* __restrict__ u4, double* __restrict__ blk_part) {
    if (done && *done) return;
    const int j = blockIdx.x * kPassBLmPerBlock + (threadIdx.x >> 3);
    const int sub = threadIdx.x & 7;
    double tx = 0, ty = 0, tz = 0, w0 = 0, w1 = 0, w2 = 0;
    if (j < L) {
        const int k0 = lm_ptr[j], k1 = lm_ptr[j + 1];
        for (int k = k0 + sub; k < k1; k += kLmLanes) {
            const d4 a = ldg256(v4 + 4 * (size_t)k);
            tx += a.x; ty += a.y; tz += a.z;
        }
        if (sub < 6) {
            w0 = WK[18 * (size_t)j + 3 * sub]; w1 = WK[18 * (size_t)j + 3 * sub + 1]; w2 = WK[18 * (size_t)j + 3 * sub + 2];
            const double xk = xK[sub];
            tx += w0 * xk; ty += w1 * xk; tz += w2 * xk;
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { tx += __shfl_xor_sync(kFull, tx, o); ty += __shfl_xor_sync(kFull, ty, o); tz += __shfl_xor_sync(kFull, tz, o); }
    double wu = 0.0;
    if (j < L) {
        const double* ci = Cinv + 6 * (size_t)j;
        double ux, uy, uz;
        sym3_mul(ci, tx, ty, tz, ux, uy, uz);
        if (BACKSUB) { const d4 c = ldg256(cg4 + 4 * (size_t)j); ux = c.x - ux; uy = c.y - uy; uz = c.z - uz; }
        if (sub == 0) stg256(u4 + 4 * (size_t)j, ux, uy, uz, 0.0);
        wu = w0 * ux + w1 * uy + w2 * uz;   // lane r < 6: (W_Kj u_j)_r ; lanes 6, 7: 0
    }
    if (BACKSUB) return;
    // block partial of sum_j W_Kj u_j: same-row lanes of the warp's four landmarks, then the four warps
    wu += __shfl_xor_sync(kFull, wu, 8);
    wu += __shfl_xor_sync(kFull, wu, 16);
    __shared__ double s[kPassBThreads / 32][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 8) s[warp][lane] = wu;
    __syncthreads();
    if (threadIdx.x < kPcgCols) {
        double t = 0.0;
        if (threadIdx.x < 6)
#pragma unroll
            for (int w = 0; w < kPassBThreads / 32; ++w) t += s[w][threadIdx.x];
        blk_part[(size_t)blockIdx.x * kPcgCols + threadIdx.x] = t;
    }
}

// pass C (pose-major): yloc_p = - sum_m J_q^T (J_l u_j); then (if add_diag) y_p = (B_pp + lambda) x_p + B_pK x_K + yloc_p.
// CTA partial row: [ sum_p B_pK^T x_p (6) | sum_p x_p . y_p (1) ]
template <int N>
__global__ void __launch_bounds__(kCtaThreads)
schur_pass_c_kernel(int P, int64_t M, double lambda_val, int add_diag, const int* __restrict__ pose_ptr, const int* __restrict__ p_lm,
                    const double2* __restrict__ J, const double* __restrict__ u4, const double* __restrict__ x, const double* __restrict__ Bpp,
                    const double* __restrict__ BpK, const int* __restrict__ done, double* __restrict__ y, double* __restrict__ cta_part,
                    const double* __restrict__ lambda_dev, const double* __restrict__ chain, const double* __restrict__ lml,
                    const unsigned* __restrict__ xctr, double* __restrict__ xbase, long long xstride, DriftTab dt) {
    if (done && *done) return;
    if (xctr) y = xbase + (size_t)((*xctr + 1u) & 1u) * xstride;   // multi-GPU: the next exchange slot of this rank
    const double lambda = lambda_dev ? *lambda_dev : lambda_val;   // device-resident inside the captured PCG graph
    __shared__ double s_part[kWarpsPerCta][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kWarpsPerCta + warp;
    if (lane < 8) s_part[warp][lane] = 0.0;
    if (p < P) {
        double ut[6] = {0, 0, 0, 0, 0, 0};   // U = sum l x t | T = sum t over this lane's observations
        const int k0 = pose_ptr[p], k1 = pose_ptr[p + 1];
        for (int k = k0 + lane; k < k1; k += 32) {
            double jl[6];
            load_jl<N>(J, M, k, jl);
            const int j = __ldg(p_lm + k);
            const d4 u = ldg256(u4 + 4 * (size_t)j);
            const d4 l = ldg256(lml + 4 * (size_t)j);
            const double e0 = jl[0] * u.x + jl[1] * u.y + jl[2] * u.z, e1 = jl[3] * u.x + jl[4] * u.y + jl[5] * u.z;
            const double t0 = jl[0] * e0 + jl[3] * e1, t1 = jl[1] * e0 + jl[4] * e1, t2 = jl[2] * e0 + jl[5] * e1;
            ut[0] += l.y * t2 - l.z * t1; ut[1] += l.z * t0 - l.x * t2; ut[2] += l.x * t1 - l.y * t0;
            ut[3] += t0; ut[4] += t1; ut[5] += t2;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) ut[i] = warp_sum(ut[i]);
        double mine = 0.0;   // lane c keeps component c
        if (lane < N) {
            const double* c = chain + (size_t)p * 6 * N + 6 * lane;
            const double a0 = c[0], a1 = c[1], a2 = c[2], o0 = c[3], o1 = c[4], o2 = c[5];
            const double m0 = a1 * o2 - a2 * o1, m1 = a2 * o0 - a0 * o2, m2 = a0 * o1 - a1 * o0;
            mine = (a0 * ut[0] + a1 * ut[1] + a2 * ut[2]) - (m0 * ut[3] + m1 * ut[4] + m2 * ut[5]);
        }
        double xm = lane < N ? x[(size_t)p * N + lane] : 0.0;
        if (add_diag) {
            double xp[N];
#pragma unroll
            for (int i = 0; i < N; ++i) xp[i] = __shfl_sync(kFull, xm, i);
            if (lane < N) {
                const double* B = Bpp + (size_t)p * N * N + lane * N;
                const double* BK = BpK + (size_t)p * N * 6 + lane * 6;
                const double* xK = x + (size_t)P * N;
                double t = lambda * xm;
#pragma unroll
                for (int i = 0; i < N; ++i) t = fma(B[i], xp[i], t);
#pragma unroll
                for (int i = 0; i < 6; ++i) t = fma(BK[i], xK[i], t);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    if (dt.has[k]) {   // B_{p,p-1} x_{p-1} + B_{p,p+1} x_{p+1} = -W (x_{p-1} [link p] + x_{p+1} [link p+1])
                        const double w = dt.w[k][lane];
                        if (dt.has[k][p]) t -= w * x[(size_t)(p - 1) * N + lane];
                        if (dt.has[k][p + 1]) t -= w * x[(size_t)(p + 1) * N + lane];
                    }
                mine += t;
            }
            // B_pK^T x_p : lane c < 6 takes column c
            double bk = 0.0;
            if (lane < 6) {
#pragma unroll
                for (int i = 0; i < N; ++i) bk = fma(BpK[(size_t)p * N * 6 + i * 6 + lane], xp[i], bk);
                s_part[warp][lane] = bk;
            }
        }
        if (lane < N) y[(size_t)p * N + lane] = mine;
        const double dot = warp_sum(lane < N ? xm * mine : 0.0);
        if (lane == 0) s_part[warp][6] = dot;
    }
    __syncthreads();
    if (threadIdx.x < kPcgCols) {
        double s = 0.0;
        if (threadIdx.x < 7)
#pragma unroll
            for (int w = 0; w < kWarpsPerCta; ++w) s += s_part[w][threadIdx.x];
        cta_part[(size_t)blockIdx.x * kPcgCols + threadIdx.x] = s;
    }
}

// K rows of y = S x from the pass B / pass C partial sums (used for the warm-start residual).  One block of 1024 threads.
__global__ void __launch_bounds__(kReduceThreads)
schur_k_rows_kernel(int P, int N, double lambda, const double* __restrict__ partC, int nC, const double* __restrict__ partB, int nB,
                    const double* __restrict__ BKK, const double* __restrict__ x, double* __restrict__ yK) {
    __shared__ double scratch[kReduceThreads];
    __shared__ double sc[kPcgCols], sb[kPcgCols];
    block_sum_cols<kPcgCols>(partC, nC, sc, scratch);
    block_sum_cols<kPcgCols>(partB, nB, sb, scratch);
    if (threadIdx.x < 6) {
        const double* xK = x + (size_t)P * N;
        double yk = lambda * xK[threadIdx.x] + sc[threadIdx.x] - sb[threadIdx.x];
        for (int i = 0; i < 6; ++i) yk += BKK[6 * threadIdx.x + i] * xK[i];
        yK[threadIdx.x] = yk;
    }
}

// x_p += alpha p_p; r_p -= alpha y_p; z_p = Minv_p r_p; CTA partial of r_p . z_p.  One warp per keyframe.
template <int N>
__global__ void __launch_bounds__(kCtaThreads)
pcg_update_kernel(int P, const double* __restrict__ p, const double* __restrict__ y, const double* __restrict__ Minv, double* __restrict__ x,
                  double* __restrict__ r, double* __restrict__ z, const double* __restrict__ scal, const int* __restrict__ done,
                  double* __restrict__ cta_part) {
    if (*done) return;
    __shared__ double s_part[kWarpsPerCta];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pp = blockIdx.x * kWarpsPerCta + warp;
    double dot = 0.0;
    if (pp < P) {
        const double alpha = scal[0];
        double rm = 0.0;
        if (lane < N) {
            const size_t i = (size_t)pp * N + lane;
            x[i] += alpha * p[i];
            rm = r[i] - alpha * y[i];
            r[i] = rm;
        }
        double zm = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double ri = __shfl_sync(kFull, rm, i);
            if (lane < N) zm = fma(Minv[(size_t)pp * N * N + lane * N + i], ri, zm);
        }
        if (lane < N) z[(size_t)pp * N + lane] = zm;
        dot = warp_sum(lane < N ? rm * zm : 0.0);
    }
    if (lane == 0) s_part[warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) s += s_part[w];
        cta_part[blockIdx.x] = s;
    }
}

// Z^T y over the keyframe rows: ry[c N + t] = sum of y[p N + t] over the keyframes p of cluster c (K entries of ry stay 0).
// One warp per cluster; feeds the side-stream V-cycle t = V (Z^T y) while pcg_mid / the cluster update run on the main stream.
__global__ void __launch_bounds__(128) coarse_restrict_y_kernel(int nC, int N, const int* __restrict__ cl_ptr, const double* __restrict__ y,
                                                                 double* __restrict__ ry, const int* __restrict__ done, PeerComm pcm) {
    if (done && *done) return;
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= nC || lane >= N) return;
    double a = 0.0;
    if (pcm.world > 1) {   // multi-GPU: y = sum over ranks of the exchange slots (launched after pcg_mid, which waited for the peers)
        const unsigned seq = *pcm.ctr;
        for (int p = cl_ptr[c]; p < cl_ptr[c + 1]; ++p) a += peer_sum(pcm, seq, (long long)p * N + lane);
    } else {
        for (int p = cl_ptr[c]; p < cl_ptr[c + 1]; ++p) a += y[(size_t)p * N + lane];
    }
    ry[(size_t)c * N + lane] = a;
}

// Coarse correction zc = B [rc; rc_K] of the two-level preconditioner from the V-cycle output t = V rc (coarse_ml.cuh):
//   z_K = S_k^-1 (src_K - E_Kc t),   z = t - Y z_K            (bordering of the extrinsic columns around the V-cycle)
// direct mode (alpha == null):  zc  = [z; z_K]                  with src_K = K entries of the restriction rc
// linear mode (alpha != null):  zc -= alpha [z; z_K]            with src_K = y_K and t = V (Z^T y): the correction is advanced by
//                               linearity, so the V-cycle of an iteration runs on a side stream beside pcg_mid / the cluster update
struct BorderArgs {
    const double* t;       // [n1] V-cycle output (null: no coarse level)
    const double* EK;      // [n1][6]
    const double* Y;       // [n1][6]
    const double* Ski;     // [36]
    const double* srcK;    // [6]
    const double* alpha;   // null = direct mode
    double* zc;            // [n1 + 6]
    int n1;
};

template <int T>
__device__ __forceinline__ void coarse_border_apply(const BorderArgs& b) {
    __shared__ double sg[T / 32][8];
    __shared__ double szk[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double g[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < b.n1; i += T) {
        const double ti = __ldcg(b.t + i);
#pragma unroll
        for (int r = 0; r < 6; ++r) g[r] = fma(b.EK[(size_t)i * 6 + r], ti, g[r]);
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) { const double w = warp_sum(g[r]); if (lane == 0) sg[warp][r] = w; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double tot = 0.0;
        for (int w = 0; w < T / 32; ++w) tot += sg[w][threadIdx.x];
        sg[0][threadIdx.x] = __ldcg(b.srcK + threadIdx.x) - tot;   // only row 0 of sg is read below; every thread < 6 wrote after reading its column
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double zk = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) zk = fma(b.Ski[6 * threadIdx.x + r], sg[0][r], zk);
        szk[threadIdx.x] = zk;
    }
    __syncthreads();
    const double al = b.alpha ? *b.alpha : 0.0;
    for (int i = threadIdx.x; i < b.n1; i += T) {
        double z = __ldcg(b.t + i);
#pragma unroll
        for (int r = 0; r < 6; ++r) z = fma(-b.Y[(size_t)i * 6 + r], szk[r], z);
        b.zc[i] = b.alpha ? b.zc[i] - al * z : z;
    }
    if (threadIdx.x < 6) b.zc[b.n1 + threadIdx.x] = b.alpha ? b.zc[b.n1 + threadIdx.x] - al * szk[threadIdx.x] : szk[threadIdx.x];
    __syncthreads();
}

// rz_new, beta, convergence test, p_K <- z_K + beta p_K.  One CTA of T threads.
struct EndArgs {
    int P, N;
    const double* part;
    int nPart;
    double tol2;
    int max_iters;
    double *p, *z, *scal;
    int* done;
    const double *rc, *zc;   // restriction Z^T r (length n1 + 6) and coarse correction (same length); null: no coarse level
    int nz;                  // n1 + 6
    BorderArgs border;       // border.t != null: finish the coarse correction here first
};

template <int T>
__device__ __forceinline__ void pcg_end_body(const EndArgs& a) {
    __shared__ double scratch[T];
    __shared__ double so[1];
    const int lane = threadIdx.x & 31, P = a.P, N = a.N;
    if (a.border.t) coarse_border_apply<T>(a.border);
    block_sum_cols<1, T>(a.part, a.nPart, so, scratch);
    double t = so[0];
    if (a.rc) {   // r . (Z zc) = (Z^T r) . zc ; and the K rows of z get their coarse part
        __shared__ double prod[T];
        double acc = 0.0;
        for (int i = threadIdx.x; i < a.nz; i += T) acc += __ldcg(a.rc + i) * __ldcg(a.zc + i);
        prod[threadIdx.x] = acc;
        __syncthreads();
        block_sum_cols<1, T, false>(prod, T, so, scratch);
        t += so[0];
        if (threadIdx.x < 6) a.z[(size_t)P * N + threadIdx.x] += __ldcg(a.zc + a.nz - 6 + threadIdx.x);
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        const double rz_new = t + a.scal[7];
        const double beta = rz_new / a.scal[2];
        const double it = a.scal[5] + 1.0;
        if (lane < 6) a.p[(size_t)P * N + lane] = a.z[(size_t)P * N + lane] + beta * a.p[(size_t)P * N + lane];
        __syncwarp();
        if (lane == 0) {
            a.scal[1] = beta; a.scal[2] = rz_new; a.scal[5] = it;
            // scal[6]: 0 = fine, 1 = p^T S p <= 0 (pcg_mid), 2 = r^T M^-1 r negative or NaN (indefinite / garbage
            // preconditioner), 3 = iteration limit reached above the tolerance.  Only rz_new in [0, tol^2 rz0] is convergence.
            const bool conv = rz_new >= 0.0 && !(rz_new > a.tol2 * a.scal[3]);
            if (!conv && !(rz_new > 0.0) && a.scal[6] == 0.0) a.scal[6] = 2.0;
            if (!conv && it >= a.max_iters && a.scal[6] == 0.0) a.scal[6] = 3.0;
            if (conv || a.scal[6] != 0.0) *a.done = 1;
        }
    }
}

__global__ void __launch_bounds__(kReduceThreads) pcg_end_kernel(EndArgs a) {
    if (*a.done) return;
    pcg_end_body<kReduceThreads>(a);
}

// PCG start: x = 0, r = rhs, z = Minv r, p = z, CTA partials of r.z  (one warp per keyframe); K part by block 0 / warp 0 extra.
template <int N>
__global__ void __launch_bounds__(kCtaThreads)
pcg_init_kernel(int P, const double* __restrict__ rhs, const double* __restrict__ Minv, const double* __restrict__ MinvK, double* __restrict__ x,
                double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ cta_part) {
    __shared__ double s_part[kWarpsPerCta + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pp = blockIdx.x * kWarpsPerCta + warp;
    double dot = 0.0;
    if (pp < P) {
        double rm = lane < N ? rhs[(size_t)pp * N + lane] : 0.0;
        double zm = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double ri = __shfl_sync(kFull, rm, i);
            if (lane < N) zm = fma(Minv[(size_t)pp * N * N + lane * N + i], ri, zm);
        }
        if (lane < N) {
            const size_t i = (size_t)pp * N + lane;
            x[i] = 0.0; r[i] = rm; z[i] = zm; p[i] = zm;
        }
        dot = warp_sum(lane < N ? rm * zm : 0.0);
    }
    if (lane == 0) s_part[warp] = dot;
    if (blockIdx.x == 0 && warp == 0) {
        double rk = lane < 6 ? rhs[(size_t)P * N + lane] : 0.0, zk = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double ri = __shfl_sync(kFull, rk, i);
            if (lane < 6) zk = fma(MinvK[6 * lane + i], ri, zk);
        }
        if (lane < 6) {
            const size_t i = (size_t)P * N + lane;
            x[i] = 0.0; r[i] = rk; z[i] = zk; p[i] = zk;
        }
        const double dk = warp_sum(lane < 6 ? rk * zk : 0.0);
        if (lane == 0) s_part[kWarpsPerCta] = dk;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) s += s_part[w];
        if (blockIdx.x == 0) s += s_part[kWarpsPerCta];
        cta_part[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(kReduceThreads)
pcg_init_end_kernel(const double* __restrict__ part, int nPart, const double* __restrict__ kdot, double* __restrict__ scal,
                    int* __restrict__ done, const double* __restrict__ rc, const double* __restrict__ zc, int nz, double* __restrict__ zK,
                    double* __restrict__ pK, double lambda, double rz_ref, double tol2, BorderArgs border) {
    __shared__ double scratch[kReduceThreads];
    __shared__ double so[1];
    if (border.t) coarse_border_apply<kReduceThreads>(border);   // zc = B rc from the V-cycle output
    block_sum_cols<1>(part, nPart, so, scratch);
    double t = so[0];
    if (rc) {
        __shared__ double prod[kReduceThreads];
        double a = 0.0;
        for (int i = threadIdx.x; i < nz; i += kReduceThreads) a += rc[i] * zc[i];
        prod[threadIdx.x] = a;
        __syncthreads();
        block_sum_cols<1, kReduceThreads, false>(prod, kReduceThreads, so, scratch);
        t += so[0];
        if (threadIdx.x < 6) { const double v = zK[threadIdx.x] + zc[nz - 6 + threadIdx.x]; zK[threadIdx.x] = v; pK[threadIdx.x] = v; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (kdot) t += *kdot;
        // rz_ref > 0: warm start, convergence is measured against the preconditioned norm of the first right-hand side
        const double ref = rz_ref > 0.0 ? rz_ref : t;
        scal[0] = 0.0; scal[1] = 0.0; scal[2] = t; scal[3] = ref; scal[4] = 0.0; scal[5] = 0.0; scal[6] = 0.0; scal[7] = 0.0;
        scal[8] = lambda;   // read by the kernels of the captured PCG graph
        if (!(t >= 0.0)) scal[6] = 2.0;   // negative or NaN r^T M^-1 r: the preconditioner is not positive definite
        *done = (t > tol2 * ref && t > 0.0) ? 0 : 1;   // nothing to do (zero rhs, or the warm start already meets the tolerance)
    }
}

// ------------------------------------------------------------------ reduced-system K rows (single block)
// S_KK = B_KK + lambda I - sum_j W_Kj Cinv W_Kj^T ; rhs_K = g_K - sum_j W_Kj cg_j ; MinvK = S_KK^-1
__global__ void __launch_bounds__(256)
k_rows_kernel(double lambda, const double* __restrict__ part, int nPart, const double* __restrict__ BKK, const double* __restrict__ gK,
              double* __restrict__ MinvK, double* __restrict__ rhsK, double* __restrict__ SKK_out, int* __restrict__ fail_idx) {
    __shared__ double s[8][kPartCols];
    __shared__ double A[6][13];
    const int lane = threadIdx.x & 31;
    const double t = block_reduce_rows(part, nPart, lane, s);
    if (threadIdx.x < 32) {
        // unpack the 21 symmetric entries
        if (lane < 21) {
            int r = 0, rem = lane;
            while (rem >= 6 - r) { rem -= 6 - r; ++r; }
            const int c = r + rem;
            const double v = BKK[6 * r + c] - t + (r == c ? lambda : 0.0);
            A[r][c] = v; A[c][r] = v;
            SKK_out[6 * r + c] = v; SKK_out[6 * c + r] = v;
        }
        if (lane >= 21 && lane < 27) rhsK[lane - 21] = gK[lane - 21] - t;
        for (int i = lane; i < 36; i += 32) A[i / 6][6 + i % 6] = (i / 6 == i % 6) ? 1.0 : 0.0;
        __syncwarp();
        bool ok = true;
        for (int k = 0; k < 6; ++k) {
            const double piv = A[k][k];
            ok = ok && (piv > 0);
            const double ip = 1.0 / piv;
            __syncwarp();
            if (lane < 12) A[k][lane] *= ip;
            __syncwarp();
            for (int i = lane; i < 72; i += 32) {
                const int r = i / 12, c = i % 12;
                if (r != k && c != k) A[r][c] -= A[r][k] * A[k][c];
            }
            __syncwarp();
            if (lane < 6 && lane != k) A[lane][k] = 0.0;
            __syncwarp();
        }
        for (int i = lane; i < 36; i += 32) MinvK[i] = A[i / 6][6 + i % 6];
        if (!ok && lane == 0) atomicMin(fail_idx, 0);
    }
}

// ------------------------------------------------------------------ linearisation finish (single block)
// B_KK (full 6x6), g_K, total error from the CTA partials of the linearise kernel, the landmark-prior
// error partials and the extrinsic prior (PriorFactor<Pose3>, J = I; ArmSlamCalib.cpp:684-685).
__global__ void __launch_bounds__(256)
linearize_finish_kernel(PriorConst pc, int add_priors, const double* __restrict__ X, const double* __restrict__ cta_part, int nPart,
                        const double* __restrict__ lm_err, int nLmErr, double* __restrict__ BKK, double* __restrict__ gK, double* __restrict__ err_out) {
    __shared__ double s[8][kPartCols];
    __shared__ double sl[256];
    const int lane = threadIdx.x & 31;
    const double t = block_reduce_rows(cta_part, nPart, lane, s);
    double e = 0.0;
    for (int i = threadIdx.x; i < nLmErr; i += 256) e += lm_err[i];
    sl[threadIdx.x] = e;
    __syncthreads();
    if (threadIdx.x < 32) {
        double xi[6] = {0, 0, 0, 0, 0, 0};
        if (add_priors) pose3_local(pc.chart, pc.x_prior, X, xi);
        if (lane < 21) {
            int r = 0, rem = lane;
            while (rem >= 6 - r) { rem -= 6 - r; ++r; }
            const int c = r + rem;
            const double v = t + ((r == c && add_priors) ? pc.inv_sigma_x[r] * pc.inv_sigma_x[r] : 0.0);
            BKK[6 * r + c] = v; BKK[6 * c + r] = v;
        }
        if (lane >= 21 && lane < 27) {
            const int i = lane - 21;
            gK[i] = t - (add_priors ? xi[i] * pc.inv_sigma_x[i] * pc.inv_sigma_x[i] : 0.0);
        }
        if (lane == 27) {
            double le = 0.0;
            for (int i = 0; i < 256; ++i) le += sl[i];
            double pe = 0.0;
            for (int i = 0; i < 6; ++i) { const double w = xi[i] * pc.inv_sigma_x[i]; pe += 0.5 * w * w; }
            *err_out = t + le + pe;
        }
    }
}

// ------------------------------------------------------------------ K1e: error only
// projection factors, one thread per observation (pose-major so neighbouring lanes share the camera row)
__global__ void __launch_bounds__(256)
error_proj_kernel(Camera cm, int64_t M, const int* __restrict__ p_pose, const int* __restrict__ p_lm, const double2* __restrict__ p_uv,
                  const double* __restrict__ cam, const double* __restrict__ lm4, double* __restrict__ blk_err) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (k < M) {
        const int p = __ldg(p_pose + k), j = __ldg(p_lm + k);
        const double2 z = ldg2(p_uv + k);
        double R[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) { const d4 t = ldg256(cam + (size_t)p * 12 + 4 * i); R[4 * i] = t.x; R[4 * i + 1] = t.y; R[4 * i + 2] = t.z; R[4 * i + 3] = t.w; }
        const d4 lq = ldg256(lm4 + 4 * (size_t)j);
        ProjOut o;
        project_factor<false, false>(cm, R, R + 9, lq.x, lq.y, lq.z, z.x, z.y, o);
        e = o.err;
    }
    __shared__ double s[8];
    e = warp_sum(e);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) blk_err[blockIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// encoder factors (thread per keyframe) and landmark priors (thread per landmark), block partials
__global__ void __launch_bounds__(256)
error_prior_kernel(Camera cm, PriorConst pc, int N, int P, int L, int with_enc, const double* __restrict__ q, const double* __restrict__ enc,
                   const uint8_t* __restrict__ has_enc, const double* __restrict__ lm4, const double* __restrict__ lm_prior,
                   double* __restrict__ blk_err, DriftTab dt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < P) {
        if (with_enc && (has_enc == nullptr || has_enc[i])) {
            for (int c = 0; c < N; ++c) {
                const double qq = q[(size_t)i * N + c], ee = enc[(size_t)i * N + c];
                double d = qq - ee;
                if ((pc.continuous_mask >> c) & 1) d = atan2(sin(qq - ee), cos(qq - ee));
                if (pc.use_dead_band) d = fabs(d) < pc.dead_band ? 0.0 : (d < 0 ? d + pc.dead_band : d - pc.dead_band);
                const double w = d * pc.inv_sigma_enc[c];
                e += 0.5 * w * w;
            }
        }
        for (int k = 0; k < 2; ++k)
            if (with_enc && dt.has[k] && dt.has[k][i])   // Drift / Velocity factor on (q_{i-1}, q_i), rank 0 only like the encoders
                for (int c = 0; c < N; ++c) { const double w = drift_residual(pc, dt, k, q, N, i, c); e += 0.5 * w * w; }
    } else if (i < P + L) {
        const int j = i - P;
        const double r0 = (lm4[4 * (size_t)j] - lm_prior[3 * (size_t)j]) * cm.inv_sigma_lm, r1 = (lm4[4 * (size_t)j + 1] - lm_prior[3 * (size_t)j + 1]) * cm.inv_sigma_lm,
                     r2 = (lm4[4 * (size_t)j + 2] - lm_prior[3 * (size_t)j + 2]) * cm.inv_sigma_lm;
        const double n2 = r0 * r0 + r1 * r1 + r2 * r2;
        e = 0.5 * (cm.k2 / (cm.k2 + n2)) * n2;
    }
    __shared__ double s[8];
    e = warp_sum(e);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) blk_err[blockIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// total = sum of block partials (+ extrinsic prior on rank 0).  One block, fixed order.
__global__ void __launch_bounds__(256)
error_finish_kernel(PriorConst pc, int with_prior, const double* __restrict__ X, const double* __restrict__ blk_err, int n, double* __restrict__ out) {
    __shared__ double sl[256];
    double e = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) e += blk_err[i];
    sl[threadIdx.x] = e;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sl[threadIdx.x] += sl[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double pe = 0.0;
        if (with_prior) {
            double xi[6];
            pose3_local(pc.chart, pc.x_prior, X, xi);
            for (int i = 0; i < 6; ++i) { const double w = xi[i] * pc.inv_sigma_x[i]; pe += 0.5 * w * w; }
        }
        *out = sl[0] + pe;
    }
}

// ------------------------------------------------------------------ K4: retract  (Values::retract)
// q' = q + s_q dq, l' = l + s_l dl (padded rows), X' = X (+) dX     (RobotConfig.h:112-126, Point3, Pose3 chart)
// delta = cu * du + cn * dn lets the Dogleg blend reuse the kernel (du may be null).
__global__ void __launch_bounds__(256)
retract_kernel(int chart, int64_t nq, int L, const double* __restrict__ q, const double* __restrict__ lm4, const double* __restrict__ X,
               double cn, const double* __restrict__ nq_, const double* __restrict__ nl4, const double* __restrict__ nK, double cu,
               const double* __restrict__ uq, const double* __restrict__ ul3, const double* __restrict__ uK, double* __restrict__ q2,
               double* __restrict__ l2, double* __restrict__ X2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) q2[i] = q[i] + (cn * nq_[i] + (uq ? cu * uq[i] : 0.0));
    if (i < (int64_t)L * 4) {
        const int c = (int)(i & 3);
        const int64_t j = i >> 2;
        l2[i] = c == 3 ? 0.0 : lm4[i] + (cn * nl4[i] + (ul3 ? cu * ul3[3 * j + c] : 0.0));
    }
    if (i == 0) {
        double xi[6], O[12], Xl[12];
        for (int c = 0; c < 6; ++c) xi[c] = cn * nK[c] + (uK ? cu * uK[c] : 0.0);
        for (int c = 0; c < 12; ++c) Xl[c] = X[c];
        pose3_retract(chart, Xl, xi, O);
        for (int c = 0; c < 12; ++c) X2[c] = O[c];
    }
}

// ------------------------------------------------------------------ small dot products (deterministic)
// out[slot] = sum a[i] * b[i]; block partials then a fixed tree by the caller's finish kernel.
__global__ void __launch_bounds__(256)
dot_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, int stride_a, int width_a, int stride_b, double* __restrict__ blk) {
    // element e of the logical vector lives at a[(e / width_a) * stride_a + e % width_a]; same for b with width_a
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (e < n) {
        const int64_t row = e / width_a;
        const int col = (int)(e % width_a);
        v = a[row * stride_a + col] * b[row * stride_b + col];
    }
    __shared__ double s[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) blk[blockIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

__global__ void __launch_bounds__(256) sum_kernel(const double* __restrict__ blk, int n, double* __restrict__ out) {
    __shared__ double sl[256];
    double e = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) e += blk[i];
    sl[threadIdx.x] = e;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sl[threadIdx.x] += sl[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sl[0];
}

// ------------------------------------------------------------------ K5: Dogleg quadratic forms
// per landmark: g_j^T C_j g_j + 2 g_K^T W_Kj g_j   (undamped C)
__global__ void __launch_bounds__(256)
dogleg_landmark_quad_kernel(int L, const double* __restrict__ C, const double* __restrict__ gl, const double* __restrict__ WK,
                            const double* __restrict__ gK, double* __restrict__ blk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (j < L) {
        const double x = gl[3 * (size_t)j], y = gl[3 * (size_t)j + 1], z = gl[3 * (size_t)j + 2];
        double ox, oy, oz;
        sym3_mul(C + 6 * (size_t)j, x, y, z, ox, oy, oz);
        v = x * ox + y * oy + z * oz;
        const double* W = WK + 18 * (size_t)j;
        for (int r = 0; r < 6; ++r) v += 2.0 * gK[r] * (W[3 * r] * x + W[3 * r + 1] * y + W[3 * r + 2] * z);
    }
    __shared__ double s[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) blk[blockIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// copy gl [L][3] into a padded [L][4] vector
__global__ void pad3to4_kernel(int L, const double* __restrict__ in3, double* __restrict__ out4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)L * 4) { const int c = (int)(i & 3); out4[i] = c == 3 ? 0.0 : in3[3 * (i >> 2) + c]; }
}

__global__ void unpad4to3_kernel(int L, const double* __restrict__ in4, double* __restrict__ out3) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)L * 3) out3[i] = in4[4 * (i / 3) + i % 3];
}

// gather pose-major per-observation data back to the caller's observation order (parity accessor)
template <int N>
__global__ void unpack_blocks_kernel(int64_t M, const int* __restrict__ p_orig, const double2* __restrict__ J, const double2* __restrict__ b,
                                     double* __restrict__ b_out, double* __restrict__ jq_out, double* __restrict__ jl_out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const int64_t m = p_orig[k];
    double f[2 * N + 6];
    load_planes<N>(J, M, (int)k, f);
    for (int i = 0; i < 2 * N; ++i) jq_out[m * 2 * N + i] = f[i];
    for (int i = 0; i < 6; ++i) jl_out[m * 6 + i] = f[2 * N + i];
    b_out[2 * m] = b[k].x; b_out[2 * m + 1] = b[k].y;
}

}  // namespace asck
