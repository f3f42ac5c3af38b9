// Host side of the B200 batch optimiser: the C ABI of include/asc.h, problem packing, the
// LM / Dogleg controllers and the kernel launch sequences.
// There is no CPU fallback in this file: every compute entry point needs an sm_100 device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only; the library is dlopen()ed when a communicator is requested
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/asc.h"
#include "host_math.h"
#include "kernels.cuh"
#include "coarse_ml.cuh"

using namespace asck;

// DOFs with compiled kernels.  data/dof_experiments.py sweeps 4..19; BASELINE configs use 6..12.
#ifndef ASC_DOF_LIST
#define ASC_DOF_LIST(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)
#endif

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) return fail(ASC_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ---- NCCL through dlopen: no link-time dependency, and the already-loaded libnccl (torch's) is reused
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.lib) return ASC_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(ASC_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(lib, "ncclAllGather");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(ASC_ERR_COMM, "libnccl is missing required symbols");
    g_nccl.lib = lib;
    return ASC_OK;
}

int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ASC_DEBUG_SYNC=1: synchronise after the named launch and report the first failing kernel (compute-sanitizer stand-in)
#define DBG_SYNC(h, name)                                                                                             \
    do {                                                                                                              \
        static const bool dbg__ = getenv("ASC_DEBUG_SYNC") != nullptr;                                                \
        if (dbg__) {                                                                                                  \
            cudaError_t e__ = cudaStreamSynchronize((h)->stream);                                                     \
            if (e__ != cudaSuccess) return fail(ASC_ERR_CUDA, "kernel %s failed: %s", name, cudaGetErrorString(e__)); \
        }                                                                                                             \
    } while (0)

}  // namespace

namespace {

// Stable parallel counting sort: dest[i] = position of element i when elements are grouped by key(i) (ties keep their
// order); ptr[b] = start of bucket b.  Host packing of a 2M-observation graph is on the end-to-end path of every optimize().
template <typename KeyFn>
void parallel_bucket_sort(int64_t n, int nb, KeyFn key, std::vector<int>& ptr, int* dest) {
    const int Tmax = std::max(1, std::min(omp_get_max_threads(), 16));
    std::vector<int> hist((size_t)Tmax * nb, 0);
    ptr.assign((size_t)nb + 1, 0);
    // The team may be SMALLER than requested (nested region with nesting off, OMP_DYNAMIC, OMP_THREAD_LIMIT, a cgroup
    // limit): every slice is derived from the team size the runtime actually delivered, never from Tmax.
#pragma omp parallel num_threads(Tmax)
    {
        const int T = omp_get_num_threads(), t = omp_get_thread_num();
        const int64_t lo = n * t / T, hi = n * (t + 1) / T;
        int* h = hist.data() + (size_t)t * nb;
        for (int64_t i = lo; i < hi; ++i) h[key(i)]++;
#pragma omp barrier
#pragma omp for schedule(static)
        for (int b = 0; b < nb; ++b) {
            int s = 0;
            for (int u = 0; u < T; ++u) { const int c = hist[(size_t)u * nb + b]; hist[(size_t)u * nb + b] = s; s += c; }
            ptr[b + 1] = s;   // bucket total for now
        }
#pragma omp single
        {
            int run = 0;
            for (int b = 0; b < nb; ++b) { const int c = ptr[b + 1]; ptr[b] = run; run += c; }
            ptr[nb] = run;
        }
        for (int64_t i = lo; i < hi; ++i) { const int b = key(i); dest[i] = ptr[b] + h[b]++; }
    }
}

}  // namespace


// ---- host side of the multilevel coarse solve (coarse_ml.cuh): patterns, aggregates and index lists, built once per packed graph
namespace {

struct MlHostLevel {
    int n = 0;
    std::vector<int> row_ptr, col, diag_pos;   // BSR pattern, both triangles, columns ascending
    std::vector<double> w;                     // static edge weight per block (shared-landmark pair count, summed on coarsening)
    std::vector<int> agg, mptr, midx;          // aggregation to the next level
    std::vector<int> cptr, cidx;               // children of each block of the NEXT level among this level's blocks
    int n_next = 0;
};

// Greedy heavy-edge aggregation: seeds in order of decreasing weighted degree; an aggregate grows by the unassigned neighbour
// with the largest total weight to its current members (ties: smallest index) until it holds `gs` nodes.  Deterministic.
void ml_aggregate(const MlHostLevel& L, int gs, std::vector<int>& agg, int& na) {
    const int n = L.n;
    agg.assign(n, -1);
    na = 0;
    std::vector<double> deg(n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int k = L.row_ptr[i]; k < L.row_ptr[i + 1]; ++k)
            if (L.col[k] != i) deg[i] += L.w[k];
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return deg[a] > deg[b]; });
    std::vector<double> cand(n, 0.0);
    std::vector<int> touched;
    for (int seed : order) {
        if (agg[seed] >= 0) continue;
        touched.clear();
        int size = 0, cur = seed;
        for (;;) {
            agg[cur] = na; ++size;
            for (int k = L.row_ptr[cur]; k < L.row_ptr[cur + 1]; ++k) {
                const int j = L.col[k];
                if (agg[j] < 0) { if (cand[j] == 0.0) touched.push_back(j); cand[j] += L.w[k] > 0 ? L.w[k] : 1e-300; }
            }
            if (size >= gs) break;
            int best = -1; double bw = 0.0;
            for (int j : touched)
                if (agg[j] < 0 && (cand[j] > bw || (cand[j] == bw && best >= 0 && j < best))) { best = j; bw = cand[j]; }
            if (best < 0) break;
            cur = best;
        }
        for (int j : touched) cand[j] = 0.0;
        ++na;
    }
}

// next level's pattern = image of this level's blocks under the aggregation; child lists in block order
void ml_coarsen_pattern(MlHostLevel& L, MlHostLevel& C) {
    const int na = L.n_next;
    C.n = na;
    std::vector<std::vector<std::pair<int, int>>> rows(na);   // per coarse row: (coarse col, fine block)
    for (int i = 0; i < L.n; ++i)
        for (int k = L.row_ptr[i]; k < L.row_ptr[i + 1]; ++k) rows[L.agg[i]].push_back({L.agg[L.col[k]], k});
    C.row_ptr.assign(na + 1, 0); C.col.clear(); C.w.clear(); C.diag_pos.assign(na, -1);
    L.cptr.assign(1, 0); L.cidx.clear();
    for (int a = 0; a < na; ++a) {
        auto& r = rows[a];
        std::stable_sort(r.begin(), r.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first < y.first; });
        size_t q = 0;
        while (q < r.size()) {
            const int cb = r[q].first;
            double wsum = 0.0;
            if (cb == a) C.diag_pos[a] = (int)C.col.size();
            while (q < r.size() && r[q].first == cb) { L.cidx.push_back(r[q].second); wsum += L.w[r[q].second]; ++q; }
            C.col.push_back(cb); C.w.push_back(wsum);
            L.cptr.push_back((int)L.cidx.size());
        }
        C.row_ptr[a + 1] = (int)C.col.size();
    }
    L.mptr.assign(na + 1, 0);
    for (int i = 0; i < L.n; ++i) L.mptr[L.agg[i] + 1]++;
    for (int a = 0; a < na; ++a) L.mptr[a + 1] += L.mptr[a];
    L.midx.assign(L.n, 0);
    std::vector<int> fill(L.mptr.begin(), L.mptr.end() - 1);
    for (int i = 0; i < L.n; ++i) L.midx[fill[L.agg[i]]++] = i;
}

}  // namespace

struct asc_handle {
    asc_params prm;
    int dev = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;              // side stream: cluster blocks run beside the coarse factorisation (single GPU)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // stage boundaries inside one damped solve (asc_iter_log)
    double ms_prepare = 0, ms_coarse = 0, ms_pcg = 0, ms_error = 0;                        // accumulated over the inner steps of one iterate()
    int N = 0, P = 0, L = 0;
    int64_t M = 0;
    bool chain_set = false, cam_set = false, problem_set = false, values_set = false, linearized = false;
    double lin_lambda = -1.0;   // lambda the current Cinv/Minv were prepared for (<0: none)
    // host copies
    std::vector<double> t_p2j, axis, t_c2j, t_base;
    std::vector<uint8_t> continuous;
    Camera cm{};
    PriorConst pc{};
    // multi-GPU
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    // peer-memory exchange of the PCG loop (kernels.cuh, PeerComm): one IPC-shared buffer per rank
    PeerComm pcm{};
    bool p2p_ok = false;
    char* xbuf = nullptr;           // this rank's exchange buffer [flags 256 B | slot 0 | slot 1]
    size_t xbuf_bytes = 0;
    void* peer_map[kMaxPeers] = {nullptr};   // cudaIpcOpenMemHandle mappings of the peers' buffers
    // failure report
    char fail_kind = 0;
    int64_t fail_index = -1;
    // device memory
    std::vector<void*> allocs;
    char* arena = nullptr;          // one grow-only device arena: re-packing a graph does not cudaFree/cudaMalloc
    size_t arena_bytes = 0;
    int *pose_ptr = nullptr, *p_lm = nullptr, *p_pose = nullptr, *p_orig = nullptr;
    int *lm_ptr = nullptr, *l_pose = nullptr, *l_perm = nullptr, *p_inv = nullptr;
    double2 *p_uv = nullptr, *l_uv = nullptr, *J = nullptr, *b_dbg = nullptr;
    double *enc = nullptr, *lm_prior = nullptr;
    uint8_t* has_enc = nullptr;
    // DriftFactor links (asc_set_drift): device tables live outside the arena (set after asc_set_problem)
    uint8_t* d_has_link[2] = {nullptr, nullptr};   // [P + 1] per kind (0 drift, 1 velocity)
    double *d_link_meas[2] = {nullptr, nullptr}, *d_link_w[2] = {nullptr, nullptr};
    bool link_set[2] = {false, false};
    DriftTab drift() const {
        DriftTab t;
        for (int k = 0; k < 2; ++k) { t.has[k] = link_set[k] ? d_has_link[k] : nullptr; t.meas[k] = d_link_meas[k]; t.w[k] = d_link_w[k]; }
        return t;
    }
    double *q = nullptr, *lm4 = nullptr, *X = nullptr, *q_try = nullptr, *lm4_try = nullptr, *X_try = nullptr;
    double *cam = nullptr, *chain = nullptr;
    double *lml = nullptr;       // landmark positions at the linearisation point (stable pointer: lm4 / lm4_try swap on accept)
    double *poseblk = nullptr;   // [Bpp | BpK | BKK(36) | gp | gK(6) | err | pad]   (one all-reduce)
    double *Bpp = nullptr, *BpK = nullptr, *BKK = nullptr, *gr = nullptr, *err_lin = nullptr;
    double *C = nullptr, *gl = nullptr, *WK = nullptr, *Cinv = nullptr, *cg4 = nullptr, *u4 = nullptr, *v4 = nullptr, *dl4 = nullptr, *gl4 = nullptr;
    double *precblk = nullptr;   // [Dloc | rloc | krow(32)]
    double *Dloc = nullptr, *rloc = nullptr, *krow = nullptr;
    double *Minv = nullptr, *MinvK = nullptr, *rhs = nullptr;
    double *vx = nullptr, *vr = nullptr, *vz = nullptr, *vp = nullptr;
    double *ybuf = nullptr;      // [y (PN) | yK(6) pad to 8 | rowC(32) | rowB(32)]
    double *partA = nullptr, *partB = nullptr, *partC = nullptr, *partL = nullptr, *partR = nullptr, *errblk = nullptr;
    int nCtaPose = 0, nBlkLm = 0, nErrBlk = 0;
    int nBlkB = 0;               // CTAs of schur_pass_b_kernel (kPassBLmPerBlock landmarks each)
    int nBlkLmK2 = 0;            // CTAs of landmark_blocks_kernel (kLmPerBlock landmarks each)
    double* lmerr = nullptr;     // their landmark-prior error partials
    double *scal = nullptr, *status = nullptr;
    int *flags = nullptr;        // [0] pcg done, [1] landmark fail idx, [2] pose fail idx, [3] K fail
    unsigned long long* cheir = nullptr;
    // measurement hooks
    long long launches = 0;
    cudaEvent_t tm[2] = {nullptr, nullptr};
    int prof_id = 0, prof_n = 0;
    std::vector<cudaEvent_t> prof_ev;
    double *q_save = nullptr, *lm4_save = nullptr, *X_save = nullptr, *stage3 = nullptr;
    // cluster-Jacobi preconditioner
    bool use_clusters = false, dup_obs = false;   // dup_obs: some landmark is observed twice from one keyframe
    int gmax = 1;
    int nC = 0, nmax = 0;
    size_t scTotal = 0, regionA = 0;
    int *cl_ptr = nullptr, *pose_cluster = nullptr, *p_run_lo = nullptr, *p_run_hi = nullptr;
    long long* cl_off = nullptr;
    double *Wm = nullptr, *Mc = nullptr;
    // coarse level (two-level additive preconditioner; multilevel V-cycle on its block-sparse operator, coarse_ml.cuh)
    bool use_coarse = false;
    int n1 = 0, nz = 0, nInc = 0, nDest = 0;   // n1 = nC * N cluster unknowns of the coarse space, nz = n1 + 6
    int* pose_agg = nullptr;                   // keyframe -> coarse node (= its cluster)
    double *ryc = nullptr;                     // side branch: Z^T y
    cudaEvent_t ev_c = nullptr, ev_q = nullptr;
    int *inc_lo = nullptr, *inc_hi = nullptr, *inc_lm = nullptr, *dest_ptr = nullptr, *pos_lo = nullptr, *pos_up = nullptr;
    int *pair_a = nullptr, *pair_b = nullptr, *cl_inc_ptr = nullptr, *cl_inc = nullptr;
    double *Vinc = nullptr, *Tinc = nullptr, *rc = nullptr, *zc = nullptr, *SKK = nullptr;
    struct MlLevel {                // one level of the hierarchy; the last one is the dense top level
        int n = 0, nnzb = 0, n_next = 0;
        int *row_ptr = nullptr, *col = nullptr, *diag_pos = nullptr;          // BSR pattern (both triangles)
        int2* bpos = nullptr;                                                 // where each block's elements live (transposed rows, coarse_ml.cuh)
        int *agg = nullptr, *mptr = nullptr, *midx = nullptr;                 // node -> aggregate; members of each aggregate
        int *cptr = nullptr, *cidx = nullptr;                                 // blocks of THIS level that sum into each block of the next
        double *Eb = nullptr, *Dinv = nullptr, *lmax = nullptr;
        double *r = nullptr, *za = nullptr, *zb = nullptr, *d = nullptr, *rs = nullptr;   // V-cycle vectors (r of level 0 is the caller's)
    };
    std::vector<MlLevel> ml;
    double *mlblk = nullptr;        // [Eb of level 0 | EK]: one all-reduce in a multi-GPU run
    double *EK = nullptr, *Yb = nullptr, *Ski = nullptr;   // extrinsic border of level 1, Y = V E_cK, S_k^-1
    double *Atop = nullptr, *ztop = nullptr, *mlpart = nullptr, *mlcol = nullptr;
    unsigned* mlbar = nullptr;      // grid-barrier counters of the persistent V-cycle kernel ([0..1]) and of the top-level inversion ([2..3])
    double* mlpiv = nullptr;        // pivot-row broadcast buffer of the top-level inversion
    unsigned long long* ml_trace = nullptr;   // ASC_ML_TRACE: barrier timestamps of the last V-cycle launch (debugging)
    int ntop = 0, ml_degree = 4;
    double ml_frac = 10.0;
    const double* ml_out = nullptr;  // where the V-cycle leaves its result (static for a given hierarchy and degree)
    int coarse_last_iters = 0;
    int outer_iteration = 0;
    bool warm_ok = false;           // vx holds the solution of an earlier damping value of the SAME linearisation
    double warm_rz_ref = 0.0;       // preconditioned norm^2 of that first right-hand side
    // CUDA graph of one chunk of PCG iterations (single GPU; all kernel arguments are handle-lifetime pointers)
    cudaGraphExec_t pcg_graph = nullptr;
    int pcg_graph_key = -1, pcg_graph_launches = 0;        // LM/Dogleg iterate() count inside the running asc_optimize call
    double* h_status = nullptr;  // pinned: status(32) + scal(8)
    int* h_flags = nullptr;      // pinned
};

namespace {

template <typename T>
int dev_alloc(asc_handle* h, T** out, size_t count) {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    CUDA_TRY(cudaMemsetAsync(p, 0, std::max<size_t>(count, 1) * sizeof(T), h->stream));
    h->allocs.push_back(p);
    *out = (T*)p;
    return ASC_OK;
}

void free_device(asc_handle* h) {
    for (void* p : h->allocs) cudaFree(p);
    h->allocs.clear();
}

struct ArenaReq { void** dst; size_t bytes; };

// lays the requested buffers out in the handle's arena (256-byte aligned), growing it only when needed
int commit_arena(asc_handle* h, const std::vector<ArenaReq>& reqs) {
    size_t total = 0;
    for (const ArenaReq& r : reqs) total += (r.bytes + 255) & ~(size_t)255;
    if (total > h->arena_bytes) {
        if (h->arena) CUDA_TRY(cudaFree(h->arena));
        h->arena = nullptr; h->arena_bytes = 0;
        const size_t want = total + total / 8 + (1 << 20);
        CUDA_TRY(cudaMalloc((void**)&h->arena, want));
        CUDA_TRY(cudaMemsetAsync(h->arena, 0, want, h->stream));
        h->arena_bytes = want;
    }
    size_t off = 0;
    for (const ArenaReq& r : reqs) { *r.dst = h->arena + off; off += (r.bytes + 255) & ~(size_t)255; }
    return ASC_OK;
}

int allreduce(asc_handle* h, double* buf, size_t count) {
    if (h->world <= 1) return ASC_OK;
    ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess) return fail(ASC_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return ASC_OK;
}


// Multi-GPU: every rank must build the same coarse hierarchy, but a rank only sees the cluster pairs its own landmarks couple.
// Union of the sorted key lists of all ranks (weights summed), through two all-gathers of padded device buffers.
int union_block_keys(asc_handle* h, std::vector<long long>& keys, std::vector<double>& w) {
    if (!g_nccl.AllGather) return fail(ASC_ERR_COMM, "libnccl has no ncclAllGather");
    const int G = h->world;
    long long* d_cnt = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_cnt, sizeof(long long) * G));
    long long mine = (long long)keys.size();
    CUDA_TRY(cudaMemcpyAsync(d_cnt + h->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    if (g_nccl.AllGather(d_cnt + h->rank, d_cnt, 1, ncclInt64, h->comm, h->stream) != ncclSuccess) { cudaFree(d_cnt); return fail(ASC_ERR_COMM, "ncclAllGather failed"); }
    std::vector<long long> cnt(G);
    CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(long long) * G, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(d_cnt);
    size_t nmax = 1;
    for (int r = 0; r < G; ++r) nmax = std::max(nmax, (size_t)cnt[r]);
    long long* d_keys = nullptr; double* d_w = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_keys, sizeof(long long) * nmax * G));
    CUDA_TRY(cudaMalloc((void**)&d_w, sizeof(double) * nmax * G));
    CUDA_TRY(cudaMemcpyAsync(d_keys + nmax * h->rank, keys.data(), sizeof(long long) * keys.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(d_w + nmax * h->rank, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice, h->stream));
    if (g_nccl.AllGather(d_keys + nmax * h->rank, d_keys, nmax, ncclInt64, h->comm, h->stream) != ncclSuccess ||
        g_nccl.AllGather(d_w + nmax * h->rank, d_w, nmax, ncclDouble, h->comm, h->stream) != ncclSuccess) {
        cudaFree(d_keys); cudaFree(d_w);
        return fail(ASC_ERR_COMM, "ncclAllGather failed");
    }
    std::vector<long long> ak(nmax * G);
    std::vector<double> aw(nmax * G);
    CUDA_TRY(cudaMemcpyAsync(ak.data(), d_keys, sizeof(long long) * nmax * G, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(aw.data(), d_w, sizeof(double) * nmax * G, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(d_keys); cudaFree(d_w);
    // G-way merge of sorted lists
    std::vector<size_t> pos(G, 0);
    keys.clear(); w.clear();
    for (;;) {
        long long best = -1;
        for (int r = 0; r < G; ++r)
            if (pos[r] < (size_t)cnt[r]) { const long long k = ak[nmax * r + pos[r]]; if (best < 0 || k < best) best = k; }
        if (best < 0) break;
        double ws = 0.0;
        for (int r = 0; r < G; ++r)
            if (pos[r] < (size_t)cnt[r] && ak[nmax * r + pos[r]] == best) { ws += aw[nmax * r + pos[r]]; ++pos[r]; }
        keys.push_back(best); w.push_back(ws);
    }
    return ASC_OK;
}

inline void prof_start(asc_handle* h, int id) {
    if (h->prof_id == id && 2 * h->prof_n + 1 < (int)h->prof_ev.size()) cudaEventRecord(h->prof_ev[2 * h->prof_n], h->stream);
}
inline void prof_stop(asc_handle* h, int id) {
    if (h->prof_id == id && 2 * h->prof_n + 1 < (int)h->prof_ev.size()) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], h->stream); ++h->prof_n; }
}

template <int N>
ChainConst<N> make_chain_const(const asc_handle* h) {
    ChainConst<N> cc;
    asc::Xf base = asc::xf_from12(h->t_base.data());
    asc::Xf f0 = asc::xf_mul(base, asc::xf_from12(&h->t_p2j[0]));
    asc::xf_to12(f0, cc.F0);
    for (int i = 0; i < N; ++i) {
        asc::Xf li = asc::xf_inv(asc::xf_from12(&h->t_c2j[12 * i]));
        if (i + 1 < N) li = asc::xf_mul(li, asc::xf_from12(&h->t_p2j[12 * (i + 1)]));
        asc::xf_to12(li, cc.Lk[i]);
        for (int k = 0; k < 3; ++k) cc.axis[i][k] = h->axis[3 * i + k];
    }
    return cc;
}

// ------------------------------------------------------------------ launch sequences (templated on DOF)
template <int N>
struct Engine {
    static int fk(asc_handle* h, const double* q, const double* X, bool with_chain) {
        const ChainConst<N> cc = make_chain_const<N>(h);
        const int grid = cdiv(h->P, 128);
        if (with_chain) fk_kernel<N, true><<<grid, 128, 0, h->stream>>>(cc, q, X, h->P, h->cam, h->chain);
        else fk_kernel<N, false><<<grid, 128, 0, h->stream>>>(cc, q, X, h->P, h->cam, h->chain);
        CUDA_TRY(cudaGetLastError());
        return ASC_OK;
    }

    static int setup_attrs(const asc_handle* h) {
        static bool done[64] = {false};   // the opt-in is per device (and per template instance)
        const int dev = h->dev < 64 ? h->dev : 63;
        if (done[dev] && h->dev < 63) return ASC_OK;
        CUDA_TRY(cudaFuncSetAttribute(pose_schur_diag_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PrecSmem<N>::kBytes));
        CUDA_TRY(cudaFuncSetAttribute(cluster_schur_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(cluster_schur_cta_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(cluster_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        done[dev] = true;
        return ASC_OK;
    }

    // FK + K1/K2b + K2a + finish.  Leaves the undamped blocks and the error on the device.
    static int linearize(asc_handle* h, bool store_b, float* ms_lin, float* ms_asm) {
        int rc = setup_attrs(h);
        if (rc) return rc;
        const int add_priors = h->rank == 0 ? 1 : 0;
        CUDA_TRY(cudaMemsetAsync(h->cheir, 0, sizeof(unsigned long long), h->stream));
        CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
        rc = fk(h, h->q, h->X, true);
        if (rc) return rc;
        prof_start(h, 1);
        const int nCtaLin = cdiv(h->P, LinCfg<N>::KF);
        if (store_b)
            linearize_pose_kernel<N, true><<<nCtaLin, LinCfg<N>::kThreads, LinCfg<N>::kBytes, h->stream>>>(
                h->cm, h->pc, h->P, h->M, h->pose_ptr, h->p_lm, h->p_uv, h->cam, h->chain, h->lm4, h->q, h->enc, h->has_enc, add_priors,
                h->J, h->Bpp, h->BpK, h->gr, h->partA, h->b_dbg, h->cheir, h->drift());
        else
            linearize_pose_kernel<N, false><<<nCtaLin, LinCfg<N>::kThreads, LinCfg<N>::kBytes, h->stream>>>(
                h->cm, h->pc, h->P, h->M, h->pose_ptr, h->p_lm, h->p_uv, h->cam, h->chain, h->lm4, h->q, h->enc, h->has_enc, add_priors,
                h->J, h->Bpp, h->BpK, h->gr, h->partA, nullptr, h->cheir, h->drift());
        CUDA_TRY(cudaGetLastError());
        DBG_SYNC(h, "fk + linearize_pose");
        prof_stop(h, 1);
        CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
        // extrinsic block B_KK, g_K and the projection error: flat pass, partial rows appended after K1's
        const int nKk = std::max(1, std::min(cdiv(h->M, kKkThreads), 148 * 4));
        prof_start(h, 7);
        proj_kk_kernel<<<nKk, kKkThreads, 0, h->stream>>>(h->cm, h->M, h->p_pose, h->p_lm, h->p_uv, h->cam, h->lm4, h->partA + (size_t)nCtaLin * kPartCols);
        CUDA_TRY(cudaGetLastError());
        DBG_SYNC(h, "proj_kk");
        prof_stop(h, 7);
        prof_start(h, 2);
        landmark_blocks_kernel<<<h->nBlkLmK2, 128, 0, h->stream>>>(h->cm, h->L, h->lm_ptr, h->l_pose, h->l_uv, h->cam, h->lm4, h->lm_prior, h->C,
                                                                  h->gl, h->WK, h->lmerr);
        CUDA_TRY(cudaGetLastError());
        DBG_SYNC(h, "landmark_blocks");
        prof_stop(h, 2);
        h->launches += 9;
        const int nRows = nCtaLin + nKk;
        const int nR1 = std::min(64, nRows), rpb = cdiv(nRows, nR1);
        reduce_rows_kernel<<<cdiv(nRows, rpb), 256, 0, h->stream>>>(h->partA, nRows, 28, kPartCols, rpb, h->partR);
        CUDA_TRY(cudaGetLastError());
        linearize_finish_kernel<<<1, 256, 0, h->stream>>>(h->pc, add_priors, h->X, h->partR, cdiv(nRows, rpb), h->lmerr, h->nBlkLmK2, h->BKK,
                                                         h->gr + (size_t)h->P * N, h->err_lin);
        CUDA_TRY(cudaGetLastError());
        if (h->world > 1) {
            const size_t count = (size_t)h->P * (N * N + 6 * N) + 36 + (size_t)h->P * N + 6 + 1;
            rc = allreduce(h, h->poseblk, count);
            if (rc) return rc;
        }
        if (h->L) pad3to4_kernel<<<cdiv((int64_t)h->L * 4, 256), 256, 0, h->stream>>>(h->L, h->gl, h->gl4);
        CUDA_TRY(cudaGetLastError());
        if (h->L) CUDA_TRY(cudaMemcpyAsync(h->lml, h->lm4, sizeof(double) * 4 * h->L, cudaMemcpyDeviceToDevice, h->stream));
        if (h->use_clusters && h->M) {
            obs_w_kernel<N><<<cdiv(h->M, 256), 256, 0, h->stream>>>(h->M, h->J, h->Wm);
            CUDA_TRY(cudaGetLastError());
            DBG_SYNC(h, "linearize finish + obs_w");
        }
        CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
        h->linearized = true;
        h->lin_lambda = -1.0;
        h->warm_ok = false;
        if (ms_lin || ms_asm) {
            CUDA_TRY(cudaEventSynchronize(h->ev[2]));
            if (ms_lin) CUDA_TRY(cudaEventElapsedTime(ms_lin, h->ev[0], h->ev[1]));
            if (ms_asm) CUDA_TRY(cudaEventElapsedTime(ms_asm, h->ev[1], h->ev[2]));
        }
        return ASC_OK;
    }

    // damped landmark inverses, Schur-diagonal preconditioner, reduced right-hand side
    static int prepare(asc_handle* h, double lambda) {
        CUDA_TRY(cudaEventRecord(h->ev_stage[0], h->stream));
        const int big = 0x7fffffff;
        int init_flags[16] = {0, big, big, big, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [8]: peer-exchange timeout
        h->launches += 5;
        CUDA_TRY(cudaMemcpyAsync(h->flags, init_flags, sizeof init_flags, cudaMemcpyHostToDevice, h->stream));
        landmark_invert_kernel<<<h->nBlkLm, 128, 0, h->stream>>>(h->L, lambda, h->C, h->gl, h->WK, h->Cinv, h->cg4, h->partL, h->flags + 1);
        CUDA_TRY(cudaGetLastError());
        DBG_SYNC(h, "landmark_invert");
        // Single GPU: the cluster blocks (cluster_schur + cluster_invert, ~3 ms on C2) depend only on the landmark inverses, and
        // so does the coarse operator (assembly + multilevel set-up): the former run on a side stream beside the latter and
        // are joined at the end of prepare().  Multi-GPU keeps one stream (the blocks go through an all-reduce).
        static const bool no_side = getenv("ASC_NO_SIDE_STREAM") != nullptr;
        const bool side = h->world == 1 && h->use_clusters && !no_side;
        cudaStream_t cs = side ? h->stream2 : h->stream;
        if (side) {
            CUDA_TRY(cudaEventRecord(h->ev_fork, h->stream));
            CUDA_TRY(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
        }
        if (h->use_clusters && !h->dup_obs) {
            const size_t sm = sizeof(double) * ((size_t)h->nmax * (h->nmax | 1) + h->nmax + 2 * (size_t)h->gmax * 3 * N) + sizeof(int) * (h->gmax + 2);
            cluster_schur_cta_kernel<N><<<h->nC, 256, sm, cs>>>(lambda, h->rank == 0 ? 1 : 0, h->gmax, h->cl_ptr, h->cl_off, h->cl_inc_ptr, h->cl_inc,
                                                                      h->inc_lo, h->inc_hi, h->inc_lm, h->l_pose, h->l_perm, h->Wm, h->Cinv, h->cg4, h->Bpp,
                                                                      h->Dloc, h->rloc);
        } else if (h->use_clusters) {
            const size_t sm = sizeof(double) * kWarpsPerCta * ((size_t)N * h->nmax + 3 * N);
            cluster_schur_kernel<N><<<h->nCtaPose, kCtaThreads, sm, cs>>>(h->P, lambda, h->rank == 0 ? 1 : 0, h->nmax, h->pose_ptr, h->p_lm,
                                                                               h->p_run_lo, h->p_run_hi, h->l_pose, h->l_perm, h->Wm, h->Cinv, h->cg4,
                                                                               h->Bpp, h->pose_cluster, h->cl_ptr, h->cl_off, h->Dloc, h->rloc);
        } else {
            pose_schur_diag_kernel<N><<<h->nCtaPose, kCtaThreads, PrecSmem<N>::kBytes, h->stream>>>(h->P, h->M, h->pose_ptr, h->p_lm, h->J, h->Cinv, h->cg4,
                                                                                                  h->Dloc, h->rloc);
        }
        CUDA_TRY(cudaGetLastError());
        if (side) {   // the inverse of the blocks follows directly; the main stream waits for ev_join before it returns
            const size_t smi = sizeof(double) * (size_t)h->nmax * (h->nmax | 1);
            cluster_invert_kernel<<<h->nC, 256, smi, cs>>>(N, h->cl_ptr, h->cl_off, h->Dloc, h->gr, h->rloc, h->Mc, h->rhs, h->flags + 2);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(h->ev_join, h->stream2));
        }
        DBG_SYNC(h, "cluster_schur / pose_schur_diag");
        const int nR1 = std::min(64, h->nBlkLm), rpb = cdiv(h->nBlkLm, nR1);
        reduce_rows_kernel<<<cdiv(h->nBlkLm, rpb), 256, 0, h->stream>>>(h->partL, h->nBlkLm, 27, kPartCols, rpb, h->partR);
        CUDA_TRY(cudaGetLastError());
        const double* kpart = h->partR;
        int nk = cdiv(h->nBlkLm, rpb);
        if (h->world > 1) {
            reduce_rows_kernel<<<1, 256, 0, h->stream>>>(h->partR, nk, 27, kPartCols, nk, h->krow);
            CUDA_TRY(cudaGetLastError());
            int rc = allreduce(h, h->precblk, h->regionA + (size_t)h->P * N + kPartCols);
            if (rc) return rc;
            kpart = h->krow;
            nk = 1;
        }
        k_rows_kernel<<<1, 256, 0, h->stream>>>(lambda, kpart, nk, h->BKK, h->gr + (size_t)h->P * N, h->MinvK, h->rhs + (size_t)h->P * N, h->SKK, h->flags + 3);
        CUDA_TRY(cudaGetLastError());
        DBG_SYNC(h, "k_rows");
        if (h->use_clusters) {
            const size_t sm = sizeof(double) * (size_t)h->nmax * (h->nmax | 1);
            if (!side) cluster_invert_kernel<<<h->nC, 256, sm, h->stream>>>(N, h->cl_ptr, h->cl_off, h->Dloc, h->gr, h->rloc, h->Mc, h->rhs, h->flags + 2);
            DBG_SYNC(h, "cluster_invert");
            CUDA_TRY(cudaEventRecord(h->ev_stage[1], h->stream));
            if (h->use_coarse) {
                int rcc = coarse_refresh(h, lambda);
                if (rcc) return rcc;
            }
            CUDA_TRY(cudaEventRecord(h->ev_stage[2], h->stream));
        } else {
            pose_invert_kernel<N><<<h->nCtaPose, kCtaThreads, 0, h->stream>>>(h->P, lambda, h->Bpp, h->gr, h->Dloc, h->rloc, h->Minv, h->rhs, h->flags + 2);
            CUDA_TRY(cudaEventRecord(h->ev_stage[1], h->stream));
            CUDA_TRY(cudaEventRecord(h->ev_stage[2], h->stream));
        }
        CUDA_TRY(cudaGetLastError());
        if (side) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        h->lin_lambda = lambda;
        return ASC_OK;
    }

    // one application y = S x written into ybuf; K rows handled by the caller's kernels.
    static int schur_apply_launch(asc_handle* h, double lambda, bool update_p, double* x, const int* done, const double* lambda_dev = nullptr,
                                  bool to_slot = false) {
        h->launches += 3;
        prof_start(h, 3);
        if (update_p)
            schur_pass_a_kernel<N, true><<<h->nCtaPose, kCtaThreads, 0, h->stream>>>(h->P, h->M, h->pose_ptr, h->p_lm, h->J, x, h->vz, h->scal, done, h->v4,
                                                                                   h->use_coarse ? h->zc : nullptr, h->pose_agg, h->p_inv, h->chain, h->lml);
        else
            schur_pass_a_kernel<N, false><<<h->nCtaPose, kCtaThreads, 0, h->stream>>>(h->P, h->M, h->pose_ptr, h->p_lm, h->J, x, h->vz, h->scal, done, h->v4, nullptr,
                                                                                    nullptr, h->p_inv, h->chain, h->lml);
        CUDA_TRY(cudaGetLastError());
        prof_stop(h, 3);
        prof_start(h, 4);
        schur_pass_b_kernel<false><<<h->nBlkB, kPassBThreads, 0, h->stream>>>(h->L, h->lm_ptr, h->v4, h->WK, h->Cinv, x + (size_t)h->P * N, done,
                                                                             nullptr, h->u4, h->partB);
        CUDA_TRY(cudaGetLastError());
        prof_stop(h, 4);
        prof_start(h, 5);
        schur_pass_c_kernel<N><<<h->nCtaPose, kCtaThreads, 0, h->stream>>>(h->P, h->M, lambda, h->rank == 0 ? 1 : 0, h->pose_ptr, h->p_lm, h->J, h->u4, x,
                                                                         h->Bpp, h->BpK, done, h->ybuf, h->partC, lambda_dev, h->chain, h->lml,
                                                                         to_slot ? h->pcm.ctr : nullptr, to_slot ? h->pcm.x[h->rank] : nullptr, h->pcm.stride,
                                                                         h->drift());
        CUDA_TRY(cudaGetLastError());
        prof_stop(h, 5);
        return ASC_OK;
    }

    // ---- multilevel coarse solve (coarse_ml.cuh)
    static MlLevelDev level_dev(const asc_handle::MlLevel& L) {
        MlLevelDev d;
        d.n = L.n; d.row_ptr = L.row_ptr; d.col = L.col; d.Eb = L.Eb; d.bpos = L.bpos; d.Dinv = L.Dinv; d.lmax = L.lmax;
        return d;
    }

    // t = V rin : one symmetric V-cycle on the cluster part of the coarse operator, ONE persistent kernel (coarse_ml.cuh).  The
    // result is left at h->ml_out (the buffer sequence is static for a given hierarchy and degree).  Capturable, runs on any stream.
    static int vcycle(asc_handle* h, const double* rin, const int* done, cudaStream_t st) {
        const int nl = (int)h->ml.size() - 1;   // smoothed levels; ml[nl] is the dense top level
        const int m = h->ml_degree;
        if (nl > kCycleMaxLevels) return fail(ASC_ERR_INVALID, "internal: too many coarse levels");
        static const int ctas_env = getenv("ASC_ML_CTAS") ? atoi(getenv("ASC_ML_CTAS")) : 0;
        MlCycleArgs a{};
        a.nl = nl; a.m = m; a.ntop = h->ntop; a.frac = h->ml_frac;
        for (int l = 0; l < nl; ++l) {
            asc_handle::MlLevel& L = h->ml[l];
            a.lv[l] = level_dev(L); a.agg[l] = L.agg; a.mptr[l] = L.mptr; a.midx[l] = L.midx; a.n_next[l] = L.n_next;
            a.za[l] = L.za; a.zb[l] = L.zb; a.d[l] = L.d; a.rs[l] = L.rs;
        }
        a.r0 = rin;
        for (int l = 1; l <= nl; ++l) a.r[l] = h->ml[l].r;
        a.Atop = h->Atop; a.ztop = h->ztop; a.bar = h->mlbar; a.fail = h->flags + 9; a.done = done;
        a.trace = h->ml_trace;
        const int ctas = ctas_env > 0 ? std::min(ctas_env, 148) : 128;
        ml_vcycle_kernel<N><<<ctas, kCycleThreads, 0, st>>>(a);
        CUDA_TRY(cudaGetLastError());
        h->launches += 1;
        const double* out = h->ztop;
        if (nl > 0) {   // output buffer of level 0: pre-smoothing ends in bufs[(m-1)&1], post-smoothing flips m more times
            const double* zi = ((m - 1) & 1) ? h->ml[0].zb : h->ml[0].za;
            for (int k = 0; k < m; ++k) zi = zi == h->ml[0].za ? h->ml[0].zb : h->ml[0].za;
            out = zi;
        }
        if (h->ml_out && h->ml_out != out) return fail(ASC_ERR_INVALID, "internal: V-cycle output buffer moved");
        h->ml_out = out;
        return ASC_OK;
    }

    // Assemble the level-1 operator for the current system, set up every level (diagonal-block inverses, spectral bound,
    // Galerkin block sums), invert the dense top level, and eliminate the extrinsic border: Y = V E_cK, S_k^-1.
    static int coarse_refresh(asc_handle* h, double lambda) {
        const int nl = (int)h->ml.size() - 1;
        asc_handle::MlLevel& L0 = h->ml[0];
        CUDA_TRY(cudaMemsetAsync(h->mlbar, 0, sizeof(unsigned) * 4, h->stream));
        if (getenv("ASC_ML_TRACE") && !h->ml_trace) CUDA_TRY(cudaMalloc((void**)&h->ml_trace, sizeof(unsigned long long) * 64 * 148 * 2));
        inc_v_kernel<N><<<cdiv(std::max(h->nInc, 1), 4), 128, 0, h->stream>>>(h->nInc, h->inc_lo, h->inc_hi, h->inc_lm, h->l_perm, h->Wm, h->Cinv, h->Vinc, h->Tinc);
        DBG_SYNC(h, "inc_v");
        ml_blocks_l1_kernel<N><<<cdiv(std::max(h->nDest, 1), 4), 128, 0, h->stream>>>(h->nDest, h->dest_ptr, h->pos_lo, h->pos_up, h->pair_a, h->pair_b, h->Vinc, h->Tinc, L0.bpos, L0.Eb);
        DBG_SYNC(h, "ml_blocks_l1");
        if (h->rank == 0) ml_diag_l1_kernel<<<cdiv((int64_t)h->nC * N * N, 256), 256, 0, h->stream>>>(h->nC, N, lambda, h->cl_ptr, h->Bpp, L0.diag_pos, L0.bpos, L0.Eb);
        ml_border_l1_kernel<<<cdiv((int64_t)h->nC * N * 6, 256), 256, 0, h->stream>>>(h->nC, N, h->rank == 0 ? 1 : 0, h->cl_ptr, h->cl_inc_ptr, h->cl_inc, h->inc_lm, h->Tinc,
                                                                                   h->WK, h->BpK, h->EK);
        CUDA_TRY(cudaGetLastError());
        DBG_SYNC(h, "ml_diag_l1 / ml_border_l1");
        h->launches += 4;
        if (h->world > 1) {   // sum the shard contributions: block-sparse operator + border (C2: 8 MB; the dense matrix of round 1 was 288 MB)
            int rc2 = allreduce(h, h->mlblk, (size_t)L0.nnzb * N * N + (size_t)h->n1 * 6);
            if (rc2) return rc2;
        }
        for (int l = 0; l < nl; ++l) {
            asc_handle::MlLevel& L = h->ml[l];
            asc_handle::MlLevel& C = h->ml[l + 1];
            ml_dinv_kernel<N><<<cdiv(L.n, 64), 64, 0, h->stream>>>(L.n, L.diag_pos, L.bpos, L.Eb, L.Dinv, h->flags + 4);
            // lambda_max(D^-1 E): 16 power steps from a fixed start vector, estimate = |w_16| / |w_15|
            const int grid = cdiv(L.n, 4);
            ml_power_init_kernel<<<cdiv(L.n * N, 256), 256, 0, h->stream>>>(L.n * N, L.za);
            double *v = L.za, *w = L.zb;
            constexpr int kPowerSteps = 16;
            for (int k = 0; k < kPowerSteps; ++k) {
                double* part = k == kPowerSteps - 2 ? h->mlpart : (k == kPowerSteps - 1 ? h->mlpart + grid : nullptr);
                ml_power_step_kernel<N><<<grid, 128, 0, h->stream>>>(L.n, L.row_ptr, L.col, L.Eb, L.Dinv, v, w, part);
                std::swap(v, w);
            }
            ml_power_finish_kernel<<<1, 256, 0, h->stream>>>(h->mlpart, h->mlpart + grid, grid, L.lmax, h->ml_frac, h->ml_degree);
            ml_coarsen_kernel<<<C.nnzb, N * N, 0, h->stream>>>(C.nnzb, N * N, L.cptr, L.cidx, L.bpos, L.Eb, C.bpos, C.Eb);
            CUDA_TRY(cudaGetLastError());
            DBG_SYNC(h, "ml level set-up");
            h->launches += kPowerSteps + 4;
        }
        {   // dense top level
            asc_handle::MlLevel& T = h->ml[nl];
            CUDA_TRY(cudaMemsetAsync(h->Atop, 0, sizeof(double) * (size_t)h->ntop * h->ntop, h->stream));
            ml_top_assemble_kernel<<<T.n, N * N, 0, h->stream>>>(T.n, N, T.row_ptr, T.col, T.bpos, T.Eb, h->Atop);
            const int top_ctas = cdiv(h->ntop, kTopThreads / 32);
            ml_top_invert_kernel<<<top_ctas, kTopThreads, sizeof(double) * (size_t)h->ntop * (kTopThreads / 32 + 1), h->stream>>>(h->Atop, h->ntop, h->mlpiv, h->mlbar + 2,
                                                                                                                              h->flags + 4);
            ml_symmetrize_kernel<<<cdiv((int64_t)h->ntop * h->ntop, 256), 256, 0, h->stream>>>(h->Atop, h->ntop);
            DBG_SYNC(h, "ml top level");
            h->launches += 3;
        }
        // extrinsic border: Y = V E_cK column by column, then S_k = S_KK - E_Kc Y
        for (int c = 0; c < 6; ++c) {
            ml_border_column_kernel<<<cdiv(h->n1, 256), 256, 0, h->stream>>>(h->n1, c, h->EK, h->mlcol);
            int rcv = vcycle(h, h->mlcol, nullptr, h->stream);
            if (rcv) return rcv;
            ml_store_column_kernel<<<cdiv(h->n1, 256), 256, 0, h->stream>>>(h->n1, c, h->ml_out, h->Yb);
            h->launches += 2;
        }
        ml_border_schur_kernel<<<1, 256, 0, h->stream>>>(h->n1, h->EK, h->Yb, h->SKK, h->Ski, h->flags + 5);
        CUDA_TRY(cudaGetLastError());
        DBG_SYNC(h, "ml border");
        h->launches += 1;
        return ASC_OK;
    }

    static BorderArgs border_args(const asc_handle* h, const double* srcK, const double* alpha) {
        BorderArgs b;
        b.t = h->ml_out; b.EK = h->EK; b.Y = h->Yb; b.Ski = h->Ski; b.srcK = srcK; b.alpha = alpha; b.zc = h->zc; b.n1 = h->n1;
        return b;
    }

    // block-Jacobi (+ multilevel coarse level) PCG on the reduced system; solution left in vx.  Returns iterations through *iters.
    static int pcg(asc_handle* h, double lambda, int* iters) {
        const size_t PN = (size_t)h->P * N;
        const int nUpd = h->use_clusters ? h->nC : h->nCtaPose;
        h->launches += 3;
        const double tol2 = h->prm.pcg_rel_tol * h->prm.pcg_rel_tol;
        // A rejected LM step is re-solved with a 10x larger lambda on the same linearisation; while lambda is far below the
        // spectrum of S the solution hardly moves, so start from it (GTSAM refactorises from scratch each time).
        const bool warm = h->warm_ok && h->use_clusters && h->warm_rz_ref > 0.0;
        PeerComm single{};   // world = 0: kernels take their single-GPU path
        single.world = 1;
        BorderArgs no_border{};
        if (warm) {
            int rcw = schur_apply_launch(h, lambda, false, h->vx, nullptr, nullptr);   // ybuf = S x0 (keyframe rows)
            if (rcw) return rcw;
            if (h->world > 1) {   // one collective outside the loop: sum the shard products and the K-row partials
                double* rowC = h->ybuf + PN + 8;
                double* rowB = rowC + kPartCols;
                reduce_rows_pair_kernel<<<2, 1024, 0, h->stream>>>(h->partC, h->nCtaPose, 7, kPcgCols, rowC, h->partB, h->nBlkB, 6, kPcgCols, rowB, nullptr);
                CUDA_TRY(cudaGetLastError());
                rcw = allreduce(h, h->ybuf, PN + 8 + 2 * kPartCols);
                if (rcw) return rcw;
                schur_k_rows_kernel<<<1, kReduceThreads, 0, h->stream>>>(h->P, N, lambda, rowC, 1, rowB, 1, h->BKK, h->vx, h->ybuf + PN);
                h->launches += 1;
            } else {
                schur_k_rows_kernel<<<1, kReduceThreads, 0, h->stream>>>(h->P, N, lambda, h->partC, h->nCtaPose, h->partB, h->nBlkB, h->BKK, h->vx, h->ybuf + PN);
            }
            CUDA_TRY(cudaGetLastError());
            h->launches += 1;
        }
        if (h->use_clusters) {
            double* rc_ = h->use_coarse ? h->rc : nullptr;
            pcg_cluster_update_kernel<<<h->nC, 128, 0, h->stream>>>(warm ? 2 : 1, N, h->cl_ptr, h->cl_off, h->Mc, h->rhs, h->ybuf, h->vx, h->vr, h->vz, h->vp, h->scal,
                                                                  h->flags, h->partA, rc_, single);
            pcg_init_k_kernel<<<1, 32, 0, h->stream>>>((int64_t)PN, h->rhs, h->MinvK, h->vx, h->vr, h->vz, h->vp, h->status + 40,
                                                      rc_ ? rc_ + h->n1 : nullptr, warm ? h->ybuf + PN : nullptr);
            CUDA_TRY(cudaGetLastError());
            if (h->use_coarse) {   // zc = B rc: V-cycle on the cluster part, border finished inside pcg_init_end
                int rc3 = vcycle(h, h->rc, nullptr, h->stream);
                if (rc3) return rc3;
            }
            pcg_init_end_kernel<<<1, kReduceThreads, 0, h->stream>>>(h->partA, nUpd, h->status + 40, h->scal, h->flags, rc_, h->zc, h->nz, h->vz + PN, h->vp + PN, lambda,
                                                         warm ? h->warm_rz_ref : 0.0, tol2, h->use_coarse ? border_args(h, h->rc + h->n1, nullptr) : no_border);
        } else {
            pcg_init_kernel<N><<<h->nCtaPose, kCtaThreads, 0, h->stream>>>(h->P, h->rhs, h->Minv, h->MinvK, h->vx, h->vr, h->vz, h->vp, h->partA);
            CUDA_TRY(cudaGetLastError());
            pcg_init_end_kernel<<<1, kReduceThreads, 0, h->stream>>>(h->partA, nUpd, nullptr, h->scal, h->flags, nullptr, nullptr, 0, nullptr, nullptr, lambda, 0.0, tol2, no_border);
        }
        CUDA_TRY(cudaGetLastError());
        static const int chunk_env = getenv("ASC_PCG_CHUNK") ? atoi(getenv("ASC_PCG_CHUNK")) : 0;
        const int chunk = chunk_env > 0 ? chunk_env : 8;   // iterations per graph launch / host check of the done flag
        int launched = 0;
        const double* lam_dev = h->scal + 8;
        const bool p2p = h->world > 1 && h->p2p_ok && h->use_clusters;   // peer-memory exchange instead of ncclAllReduce
        const bool profiling = h->prof_id != 0 && 2 * h->prof_n + 1 < (int)h->prof_ev.size();   // event-bracketed launches cannot be captured
        // Coarse V-cycle off the critical path (single GPU): t = V (Z^T y) runs on a side branch of the graph beside pcg_mid and the
        // cluster update, and pcg_end advances the coarse correction by linearity: zc <- zc - alpha B [Z^T y; y_K].
        static const bool no_side_v = getenv("ASC_NO_SIDE_VCYCLE") != nullptr;
        const bool side_v = !no_side_v && (h->world == 1 || p2p) && h->use_clusters && h->use_coarse;
        // Residual replacement for the coarse correction: the first iteration of every chunk recomputes zc = B Z^T r from the residual
        // (V-cycle on the main stream), the others advance it by linearity.  The recursion alone accumulates rounding errors of
        // size eps * cond(B) * |zc_0| that do not shrink with the residual and stall CG on ill-conditioned coarse operators (small graphs,
        // small lambda: measured 134 instead of 93 iterations, and one solve that never reached 1e-12).
        auto one_iteration = [&](bool side_v) -> int {
            double* rowC = h->ybuf + PN + 8;
            double* rowB = rowC + kPartCols;
            MidArgs mid{};
            mid.P = h->P; mid.N = N; mid.lambda_val = lambda; mid.lambda_dev = lam_dev;
            mid.partC = h->partC; mid.nC = h->nCtaPose; mid.partB = h->partB; mid.nB = h->nBlkB;
            mid.BKK = h->BKK; mid.MinvK = h->MinvK; mid.p = h->vp; mid.x = h->vx; mid.r = h->vr; mid.z = h->vz;
            mid.yK_out = h->ybuf + PN; mid.scal = h->scal; mid.done = h->flags;
            mid.rcK = h->use_coarse ? h->rc + h->n1 : nullptr;
            int rc = schur_apply_launch(h, lambda, true, h->vp, h->flags, lam_dev, p2p);
            if (rc) return rc;
            auto fork_vcycle = [&]() -> int {   // t = V (Z^T y) on the side stream
                CUDA_TRY(cudaEventRecord(h->ev_c, h->stream));
                CUDA_TRY(cudaStreamWaitEvent(h->stream2, h->ev_c, 0));
                coarse_restrict_y_kernel<<<cdiv(h->nC, 4), 128, 0, h->stream2>>>(h->nC, N, h->cl_ptr, h->ybuf, h->ryc, h->flags, p2p ? h->pcm : single);
                CUDA_TRY(cudaGetLastError());
                h->launches += 1;
                int rcv = vcycle(h, h->ryc, h->flags, h->stream2);
                if (rcv) return rcv;
                CUDA_TRY(cudaEventRecord(h->ev_q, h->stream2));
                return ASC_OK;
            };
            if (side_v && !p2p) { rc = fork_vcycle(); if (rc) return rc; }   // single GPU: right after pass C
            if (p2p) {
                pcg_signal_kernel<<<1, 1024, 0, h->stream>>>(h->pcm, mid.partC, mid.nC, mid.partB, mid.nB, h->flags);
                CUDA_TRY(cudaGetLastError());
                h->launches += 1;
                mid.partC = rowC; mid.partB = rowB; mid.nC = 1; mid.nB = 1;
            } else if (h->world > 1) {
                reduce_rows_pair_kernel<<<2, 1024, 0, h->stream>>>(mid.partC, mid.nC, 7, kPcgCols, rowC, mid.partB, mid.nB, 6, kPcgCols, rowB, h->flags);
                CUDA_TRY(cudaGetLastError());
                h->launches += 1;
                rc = allreduce(h, h->ybuf, PN + 8 + 2 * kPartCols);
                if (rc) return rc;
                mid.partC = rowC; mid.partB = rowB; mid.nC = 1; mid.nB = 1;
            }
            pcg_mid_kernel<<<1, kReduceThreads, 0, h->stream>>>(mid, p2p ? h->pcm : single, rowC);
            CUDA_TRY(cudaGetLastError());
            h->launches += 1;
            if (side_v && p2p) { rc = fork_vcycle(); if (rc) return rc; }    // multi-GPU: after pcg_mid has seen every peer's partial product
            EndArgs end{};
            end.P = h->P; end.N = N; end.part = h->partA; end.nPart = nUpd; end.tol2 = tol2; end.max_iters = h->prm.pcg_max_iters;
            end.p = h->vp; end.z = h->vz; end.scal = h->scal; end.done = h->flags; end.rc = h->use_coarse ? h->rc : nullptr; end.zc = h->zc;
            end.nz = h->nz;
            if (h->use_clusters) {
                prof_start(h, 9);
                pcg_cluster_update_kernel<<<h->nC, 128, 0, h->stream>>>(0, N, h->cl_ptr, h->cl_off, h->Mc, h->rhs, h->ybuf, h->vx, h->vr, h->vz, h->vp,
                                                                      h->scal, h->flags, h->partA, h->use_coarse ? h->rc : nullptr, p2p ? h->pcm : single);
                prof_stop(h, 9);
                if (side_v) {   // join; pcg_end advances zc by linearity
                    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_q, 0));
                    end.border = border_args(h, h->ybuf + PN, h->scal);
                } else if (h->use_coarse) {   // zc = B rc recomputed from the new residual
                    prof_start(h, 8);
                    int rc3 = vcycle(h, h->rc, h->flags, h->stream);
                    prof_stop(h, 8);
                    if (rc3) return rc3;
                    end.border = border_args(h, h->rc + h->n1, nullptr);
                }
            } else {
                pcg_update_kernel<N><<<h->nCtaPose, kCtaThreads, 0, h->stream>>>(h->P, h->vp, h->ybuf, h->Minv, h->vx, h->vr, h->vz, h->scal, h->flags, h->partA);
            }
            CUDA_TRY(cudaGetLastError());
            pcg_end_kernel<<<1, kReduceThreads, 0, h->stream>>>(end);
            CUDA_TRY(cudaGetLastError());
            h->launches += 2;
            return ASC_OK;
        };
        // one chunk = `chunk` iterations; kernels return immediately once the device-side done flag is set
        static const bool no_graph = getenv("ASC_NO_GRAPH") != nullptr;
        // multi-GPU: with the peer-memory exchange the loop holds kernels only and is captured like the single-GPU one (a captured
        // ncclAllReduce hung on this stack: NCCL 2.28.9, driver 580); ASC_NO_GRAPH_MGPU=1 falls back to plain launches
        static const bool no_graph_mgpu = getenv("ASC_NO_GRAPH_MGPU") != nullptr;
        const bool can_graph = !no_graph && (h->world == 1 || (p2p && !no_graph_mgpu)) && !profiling;
        const int graph_key = (h->use_clusters ? 1 : 0) | (h->use_coarse ? 2 : 0) | (p2p ? 8 : 0) | (side_v ? 64 : 0);
        for (;;) {
            if (can_graph) {
                if (!h->pcg_graph || h->pcg_graph_key != graph_key) {
                    if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }
                    const long long before = h->launches;
                    cudaGraph_t graph = nullptr;
                    CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                    int rc = ASC_OK;
                    for (int c = 0; c < chunk && !rc; ++c) rc = one_iteration(side_v && c > 0);
                    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
                    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                    if (ce != cudaSuccess) return fail(ASC_ERR_CUDA, "PCG graph capture failed: %s", cudaGetErrorString(ce));
                    CUDA_TRY(cudaGraphInstantiate(&h->pcg_graph, graph, 0));
                    cudaGraphDestroy(graph);
                    h->pcg_graph_launches = (int)(h->launches - before);
                    h->launches = before;
                    h->pcg_graph_key = graph_key;
                }
                CUDA_TRY(cudaGraphLaunch(h->pcg_graph, h->stream));
                h->launches += h->pcg_graph_launches;
            } else {
                for (int c = 0; c < chunk; ++c) { int rc = one_iteration(side_v && c > 0); if (rc) return rc; }
            }
            launched += chunk;
            CUDA_TRY(cudaMemcpyAsync(h->h_flags, h->flags, 16 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(cudaMemcpyAsync(h->h_status + 32, h->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            if (h->h_flags[0] || launched >= h->prm.pcg_max_iters + chunk) break;
        }
        if (iters) *iters = (int)h->h_status[32 + 5];
        if (!warm) h->warm_rz_ref = h->h_status[32 + 3];
        return ASC_OK;
    }

    // delta_l = Cinv (g_l - W^T delta_r) -> dl4
    static int backsub(asc_handle* h) {
        h->launches += 2;
        schur_pass_a_kernel<N, false><<<h->nCtaPose, kCtaThreads, 0, h->stream>>>(h->P, h->M, h->pose_ptr, h->p_lm, h->J, h->vx, nullptr, nullptr, nullptr, h->v4, nullptr,
                                                                                nullptr, h->p_inv, h->chain, h->lml);
        CUDA_TRY(cudaGetLastError());
        schur_pass_b_kernel<true><<<h->nBlkB, kPassBThreads, 0, h->stream>>>(h->L, h->lm_ptr, h->v4, h->WK, h->Cinv, h->vx + (size_t)h->P * N, nullptr,
                                                                   h->cg4, h->dl4, nullptr);
        CUDA_TRY(cudaGetLastError());
        return ASC_OK;
    }

    // Multi-GPU: the failure flags that steer the host (non-PD landmark block of ONE shard, a peer-exchange timeout seen by ONE
    // rank) are rank-local; every rank must take the same branch or the next collective deadlocks.  One small all-reduce makes
    // the decision global: afterwards h_flags[1..5, 8, 9] are "somebody failed" indicators on every rank (own index kept if it has one).
    static int agree_on_failures(asc_handle* h) {
        if (h->world <= 1) return ASC_OK;
        const int big = 0x7fffffff;
        double* d = h->status + 48;
        double v[8] = {h->h_flags[1] != big ? 1.0 : 0.0, h->h_flags[2] != big ? 1.0 : 0.0, h->h_flags[3] != big ? 1.0 : 0.0,
                       h->h_flags[4] != 0 ? 1.0 : 0.0, h->h_flags[5] != 0 ? 1.0 : 0.0, h->h_flags[8] != 0 ? 1.0 : 0.0, h->h_flags[9] != 0 ? 1.0 : 0.0, 0.0};
        CUDA_TRY(cudaMemcpyAsync(d, v, sizeof v, cudaMemcpyHostToDevice, h->stream));
        int rc = allreduce(h, d, 8);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(v, d, sizeof v, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (v[0] > 0 && h->h_flags[1] == big) h->h_flags[1] = -1;   // failed on another rank: index unknown here
        if (v[1] > 0 && h->h_flags[2] == big) h->h_flags[2] = -1;
        if (v[2] > 0 && h->h_flags[3] == big) h->h_flags[3] = -1;
        if (v[3] > 0) h->h_flags[4] = 1;
        if (v[4] > 0) h->h_flags[5] = 1;
        if (v[5] > 0) h->h_flags[8] = 1;
        if (v[6] > 0) h->h_flags[9] = 1;   // a grid-barrier watchdog is timing-dependent: one rank may see it alone
        return ASC_OK;
    }

    static int check_solve_flags(asc_handle* h) {
        const int big = 0x7fffffff;
        if (h->h_flags[8] != 0) return fail(ASC_ERR_COMM, "peer-memory exchange timed out waiting for a peer rank");
        if (h->h_flags[9] != 0) return fail(ASC_ERR_CUDA, "the V-cycle kernel's grid barrier timed out (its CTAs were not co-resident)");
        if (h->h_flags[1] != big) { h->fail_kind = 'l'; h->fail_index = h->h_flags[1]; return fail(ASC_ERR_INDETERMINATE, "landmark block %d is not positive definite", h->h_flags[1]); }
        if (h->h_flags[2] != big) { h->fail_kind = 'q'; h->fail_index = h->h_flags[2]; return fail(ASC_ERR_INDETERMINATE, "pose block %d is not positive definite", h->h_flags[2]); }
        if (h->h_flags[3] != big) { h->fail_kind = 'K'; h->fail_index = 0; return fail(ASC_ERR_INDETERMINATE, "extrinsic block is not positive definite"); }
        if (h->use_coarse && (h->h_flags[4] != 0 || h->h_flags[5] != 0)) { h->fail_kind = 'q'; h->fail_index = 0; return fail(ASC_ERR_INDETERMINATE, "coarse-level blocks are not positive definite (levels %d, border %d)", h->h_flags[4], h->h_flags[5]); }
        const double bd = h->h_status[32 + 6];
        if (bd == 3.0) return fail(ASC_ERR_NOT_CONVERGED, "PCG stopped at pcg_max_iters = %d above pcg_rel_tol", h->prm.pcg_max_iters);
        if (bd != 0.0) { h->fail_kind = 'K'; h->fail_index = 0; return fail(ASC_ERR_INDETERMINATE, bd == 1.0 ? "PCG breakdown: p^T S p <= 0" : "PCG breakdown: r^T M^-1 r is negative or not finite"); }
        return ASC_OK;
    }

    // full damped solve: prepare + PCG + back-substitution.  rc ASC_ERR_INDETERMINATE on breakdown.
    static int solve(asc_handle* h, double lambda, int* iters) {
        int rc = prepare(h, lambda);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_stage[3], h->stream));
        int its_local = 0;
        rc = pcg(h, lambda, &its_local);
        if (iters) *iters = its_local;
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_stage[4], h->stream));
        {   // stage times of this solve (the PCG loop has synchronised the stream already)
            CUDA_TRY(cudaEventSynchronize(h->ev_stage[4]));
            float a = 0, b = 0, c = 0, d = 0;
            CUDA_TRY(cudaEventElapsedTime(&a, h->ev_stage[0], h->ev_stage[1]));
            CUDA_TRY(cudaEventElapsedTime(&b, h->ev_stage[1], h->ev_stage[2]));
            CUDA_TRY(cudaEventElapsedTime(&c, h->ev_stage[2], h->ev_stage[3]));
            CUDA_TRY(cudaEventElapsedTime(&d, h->ev_stage[3], h->ev_stage[4]));
            h->ms_prepare += a + c; h->ms_coarse += b; h->ms_pcg += d;
        }
        h->coarse_last_iters = its_local;
        rc = agree_on_failures(h);
        if (rc) return rc;
        rc = check_solve_flags(h);
        if (rc) return rc;
        h->warm_ok = true;
        return backsub(h);
    }

    // (S^-1)_KK of the undamped reduced system: six PCG solves with unit right-hand sides on the K0 rows
    static int marginal(asc_handle* h, double* cov36, int* iters_total) {
        const size_t PN = (size_t)h->P * N;
        int rc = prepare(h, 0.0);
        if (rc) return rc;
        int total = 0;
        for (int i = 0; i < 6; ++i) {
            CUDA_TRY(cudaMemsetAsync(h->rhs, 0, sizeof(double) * (PN + 8), h->stream));
            set_scalar_kernel<<<1, 1, 0, h->stream>>>(h->rhs + PN + i, 1.0);
            CUDA_TRY(cudaGetLastError());
            h->warm_ok = false;
            int its = 0;
            rc = pcg(h, 0.0, &its);
            total += its;
            if (rc) return rc;
            rc = agree_on_failures(h);
            if (rc) return rc;
            rc = check_solve_flags(h);
            if (rc == ASC_ERR_INDETERMINATE) {
                h->fail_kind = 'K'; h->fail_index = 0;
                return fail(ASC_ERR_INDETERMINATE, "the undamped reduced system is not positive definite");
            }
            if (rc) return rc;
            double col[6];
            CUDA_TRY(cudaMemcpyAsync(col, h->vx + PN, sizeof col, cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            for (int r = 0; r < 6; ++r) cov36[6 * r + i] = col[r];
        }
        for (int r = 0; r < 6; ++r)   // S^-1 is symmetric; average away the PCG tolerance
            for (int c = r + 1; c < 6; ++c) { const double m = 0.5 * (cov36[6 * r + c] + cov36[6 * c + r]); cov36[6 * r + c] = m; cov36[6 * c + r] = m; }
        if (iters_total) *iters_total = total;
        h->warm_ok = false;   // rhs no longer belongs to the linearisation
        return ASC_OK;
    }

    // graph.error(values) at (q, lm4, X); synchronises and returns the value
    static int error(asc_handle* h, const double* q, const double* lm4, const double* X, double* out) {
        int rc = fk(h, q, X, false);
        if (rc) return rc;
        const int nb1 = cdiv(h->M, 256), nb2 = cdiv((int64_t)h->P + h->L, 256);
        h->launches += 4;
        prof_start(h, 6);
        if (nb1) error_proj_kernel<<<nb1, 256, 0, h->stream>>>(h->cm, h->M, h->p_pose, h->p_lm, h->p_uv, h->cam, lm4, h->errblk);
        CUDA_TRY(cudaGetLastError());
        prof_stop(h, 6);
        error_prior_kernel<<<nb2, 256, 0, h->stream>>>(h->cm, h->pc, N, h->P, h->L, h->rank == 0 ? 1 : 0, q, h->enc, h->has_enc, lm4, h->lm_prior, h->errblk + nb1, h->drift());
        CUDA_TRY(cudaGetLastError());
        error_finish_kernel<<<1, 256, 0, h->stream>>>(h->pc, h->rank == 0 ? 1 : 0, X, h->errblk, nb1 + nb2, h->status + 4);
        CUDA_TRY(cudaGetLastError());
        if (h->world > 1) { rc = allreduce(h, h->status + 4, 1); if (rc) return rc; }
        CUDA_TRY(cudaMemcpyAsync(h->h_status + 4, h->status + 4, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        *out = h->h_status[4];
        // the FK table now belongs to the trial values; a following linearise recomputes it
        return ASC_OK;
    }

    static int unpack(asc_handle* h, double* d_b, double* d_jq, double* d_jl) {
        unpack_blocks_kernel<N><<<cdiv(h->M, 256), 256, 0, h->stream>>>(h->M, h->p_orig, h->J, h->b_dbg, d_b, d_jq, d_jl);
        CUDA_TRY(cudaGetLastError());
        return ASC_OK;
    }
};

int dispatch_linearize(asc_handle* h, bool store_b, float* a, float* b) {
#define X(n) case n: return Engine<n>::linearize(h, store_b, a, b);
    switch (h->N) { ASC_DOF_LIST(X) default: return fail(ASC_ERR_INVALID, "no kernels compiled for %d DOF", h->N); }
#undef X
}
int dispatch_solve(asc_handle* h, double lambda, int* iters) {
#define X(n) case n: return Engine<n>::solve(h, lambda, iters);
    switch (h->N) { ASC_DOF_LIST(X) default: return fail(ASC_ERR_INVALID, "no kernels compiled for %d DOF", h->N); }
#undef X
}
int dispatch_marginal(asc_handle* h, double* cov36, int* iters) {
#define X(n) case n: return Engine<n>::marginal(h, cov36, iters);
    switch (h->N) { ASC_DOF_LIST(X) default: return fail(ASC_ERR_INVALID, "no kernels compiled for %d DOF", h->N); }
#undef X
}
int dispatch_error(asc_handle* h, const double* q, const double* lm4, const double* X_, double* out) {
#define X(n) case n: return Engine<n>::error(h, q, lm4, X_, out);
    switch (h->N) { ASC_DOF_LIST(X) default: return fail(ASC_ERR_INVALID, "no kernels compiled for %d DOF", h->N); }
#undef X
}
int dispatch_unpack(asc_handle* h, double* b, double* jq, double* jl) {
#define X(n) case n: return Engine<n>::unpack(h, b, jq, jl);
    switch (h->N) { ASC_DOF_LIST(X) default: return fail(ASC_ERR_INVALID, "no kernels compiled for %d DOF", h->N); }
#undef X
}
int dispatch_schur_c(asc_handle* h, double lambda, double* x, double* u4_in) {
    // pass C only, with a caller-provided landmark vector (Dogleg cross term)
#define X(n)                                                                                                                      \
    case n:                                                                                                                       \
        schur_pass_c_kernel<n><<<h->nCtaPose, kCtaThreads, 0, h->stream>>>(h->P, h->M, lambda, 0, h->pose_ptr, h->p_lm, h->J, u4_in, x, \
                                                                         h->Bpp, h->BpK, nullptr, h->ybuf, h->partC, nullptr, h->chain, h->lml, nullptr, nullptr, 0, \
                                                                         DriftTab{{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}}); \
        break;
    switch (h->N) { ASC_DOF_LIST(X) default: return fail(ASC_ERR_INVALID, "no kernels compiled for %d DOF", h->N); }
#undef X
    CUDA_TRY(cudaGetLastError());
    return ASC_OK;
}
bool dof_supported(int n) {
#define X(k) if (n == k) return true;
    ASC_DOF_LIST(X)
#undef X
    return false;
}

// ---- step scalars: [x.g_r, x.x, dl.gl, dl.dl] over the reduced vector and the landmark vector
__global__ void __launch_bounds__(256)
step_dots_kernel(int64_t nr, const double* __restrict__ x, const double* __restrict__ g, int L, const double* __restrict__ dl4,
                 const double* __restrict__ gl4, double* __restrict__ blk) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[4] = {0, 0, 0, 0};
    if (i < nr) { v[0] = x[i] * g[i]; v[1] = x[i] * x[i]; }
    if (i < (int64_t)L * 4) { v[2] = dl4[i] * gl4[i]; v[3] = dl4[i] * dl4[i]; }
    __shared__ double s[8][4];
    for (int c = 0; c < 4; ++c) {
        const double t = warp_sum(v[c]);
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][c] = t;
    }
    __syncthreads();
    if (threadIdx.x < kPartCols) {
        double t = 0.0;
        if (threadIdx.x < 4)
            for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
        blk[(size_t)blockIdx.x * kPartCols + threadIdx.x] = t;
    }
}

// per keyframe g_p^T B_pp g_p + 2 g_p^T B_pK g_K   (Dogleg, |A g|^2)
__global__ void __launch_bounds__(256)
dogleg_pose_quad_kernel(int P, int N, const double* __restrict__ Bpp, const double* __restrict__ BpK, const double* __restrict__ g, double* __restrict__ blk,
                        DriftTab dt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (p < P) {
        const double* gp = g + (size_t)p * N;
        const double* gK = g + (size_t)P * N;
        for (int r = 0; r < N; ++r) {
            double t = 0.0;
            for (int c = 0; c < N; ++c) t += Bpp[(size_t)p * N * N + r * N + c] * gp[c];
            double k = 0.0;
            for (int c = 0; c < 6; ++c) k += BpK[(size_t)p * N * 6 + r * 6 + c] * gK[c];
            v += gp[r] * (t + 2.0 * k);
            for (int k = 0; k < 2; ++k)
                if (dt.has[k] && dt.has[k][p]) v -= 2.0 * dt.w[k][r] * gp[r] * g[(size_t)(p - 1) * N + r];   // 2 g_{p-1}^T B_{p-1,p} g_p
        }
    }
    __shared__ double s[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) blk[blockIdx.x] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

int step_dots(asc_handle* h, const double* x, const double* dl4, double* out4_host) {
    const int64_t nr = (int64_t)h->P * h->N + 6;
    const int64_t n = std::max<int64_t>(nr, (int64_t)h->L * 4);
    const int nb = cdiv(n, 256);
    step_dots_kernel<<<nb, 256, 0, h->stream>>>(nr, x, h->gr, h->L, dl4, h->gl4, h->partA);
    CUDA_TRY(cudaGetLastError());
    reduce_rows_kernel<<<1, 256, 0, h->stream>>>(h->partA, nb, 4, kPartCols, nb, h->status + 8);
    CUDA_TRY(cudaGetLastError());
    // multi-GPU: the reduced-vector parts are replicated, the landmark parts are sharded
    if (h->world > 1) { int rc = allreduce(h, h->status + 10, 2); if (rc) return rc; }
    CUDA_TRY(cudaMemcpyAsync(h->h_status + 8, h->status + 8, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 4; ++i) out4_host[i] = h->h_status[8 + i];
    return ASC_OK;
}

int retract_to_try(asc_handle* h, double cn, double cu) {
    const int64_t nq = (int64_t)h->P * h->N;
    const int64_t n = std::max<int64_t>(nq, (int64_t)h->L * 4);
    const bool with_u = cu != 0.0;
    retract_kernel<<<cdiv(std::max<int64_t>(n, 1), 256), 256, 0, h->stream>>>(h->prm.pose3_chart, nq, h->L, h->q, h->lm4, h->X, cn, h->vx, h->dl4,
                                                                             h->vx + nq, cu, with_u ? h->gr : nullptr, with_u ? h->gl : nullptr,
                                                                             with_u ? h->gr + nq : nullptr, h->q_try, h->lm4_try, h->X_try);
    CUDA_TRY(cudaGetLastError());
    return ASC_OK;
}

void accept_try(asc_handle* h) {
    std::swap(h->q, h->q_try);
    std::swap(h->lm4, h->lm4_try);
    std::swap(h->X, h->X_try);
    h->linearized = false;
}

bool converged(const asc_params& p, double e0, double e1) {
    if (e1 <= p.error_tol) return true;
    const double abs_dec = e0 - e1, rel_dec = abs_dec / e0;
    return (p.relative_error_tol != 0 && rel_dec <= p.relative_error_tol) || abs_dec <= p.absolute_error_tol;
}

// [DEP] LevenbergMarquardtOptimizer::iterate
int lm_iterate(asc_handle* h, double* error, double* lambda, asc_iter_log* lg) {
    const asc_params& p = h->prm;
    float ms = 0;
    h->ms_prepare = h->ms_coarse = h->ms_pcg = 0.0;
    int rc = dispatch_linearize(h, false, nullptr, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev[3], h->stream));
    for (;;) {
        int its = 0;
        rc = dispatch_solve(h, *lambda, &its);
        lg->inner_steps++;
        lg->pcg_iterations += its;
        // an indeterminate system or a PCG that ran out of iterations both count as "not solved": lambda goes up (Appendix C.2)
        if (rc != ASC_OK && rc != ASC_ERR_INDETERMINATE && rc != ASC_ERR_NOT_CONVERGED) return rc;
        const bool solved = rc == ASC_OK;
        bool ok = false, stop = false;
        double new_error = 0.0;
        if (solved) {
            double d[4];
            rc = step_dots(h, h->vx, h->dl4, d);
            if (rc) return rc;
            // linear.error(0) - linear.error(delta) for an exact damped solve: 0.5 d.g + 0.5 lambda |d|^2
            const double lin_dec = 0.5 * (d[0] + d[2]) + 0.5 * *lambda * (d[1] + d[3]);
            if (lin_dec >= 0) {
                rc = retract_to_try(h, 1.0, 0.0);
                if (rc) return rc;
                rc = dispatch_error(h, h->q_try, h->lm4_try, h->X_try, &new_error);
                if (rc) return rc;
                const double cost_dec = *error - new_error;
                ok = lin_dec > 1e-15 ? (cost_dec / lin_dec) > p.min_model_fidelity : true;
                if (std::fabs(cost_dec) < p.relative_error_tol * *error) stop = true;
            }
        }
        if (ok) {
            accept_try(h);
            *error = new_error;
            *lambda = std::max(*lambda / p.lambda_factor, p.lambda_lower_bound);
            lg->accepted = 1;
            break;
        } else if (!stop) {
            *lambda *= p.lambda_factor;
            if (*lambda >= p.lambda_upper_bound) break;
        } else break;
    }
    CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
    CUDA_TRY(cudaEventSynchronize(h->ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]));
    lg->linearize_ms = ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[3], h->ev[2]));
    lg->solve_ms = ms;
    lg->prepare_ms = h->ms_prepare; lg->coarse_ms = h->ms_coarse; lg->pcg_ms = h->ms_pcg;
    return ASC_OK;
}

// |A g|^2 from the assembled blocks and the stored Jacobian planes (K5)
int dogleg_gHg(asc_handle* h, double* out) {
    const int N = h->N;
    const size_t PN = (size_t)h->P * N;
    int rc = dispatch_schur_c(h, 0.0, h->gr, h->gl4);   // yloc = -W g_l ; partC col 6 = sum g_p . yloc_p
    if (rc) return rc;
    reduce_rows_kernel<<<1, 256, 0, h->stream>>>(h->partC, h->nCtaPose, 7, kPcgCols, h->nCtaPose, h->status + 16);   // [16+6] = T1
    const int nbp = cdiv(h->P, 256), nbl = cdiv(h->L, 256);
    dogleg_pose_quad_kernel<<<nbp, 256, 0, h->stream>>>(h->P, N, h->Bpp, h->BpK, h->gr, h->errblk, h->drift());
    dogleg_landmark_quad_kernel<<<std::max(nbl, 1), 256, 0, h->stream>>>(h->L, h->C, h->gl, h->WK, h->gr + PN, h->errblk + nbp);
    CUDA_TRY(cudaGetLastError());
    sum_kernel<<<1, 256, 0, h->stream>>>(h->errblk, h->rank == 0 ? nbp : 0, h->status + 12);
    sum_kernel<<<1, 256, 0, h->stream>>>(h->errblk + nbp, std::max(nbl, 1), h->status + 13);
    CUDA_TRY(cudaGetLastError());
    if (h->world > 1) {
        // status[12] pose quad (rank 0 only), [13] landmark quad (sharded), [22] T1 (sharded)
        rc = allreduce(h, h->status + 12, 2); if (rc) return rc;
        rc = allreduce(h, h->status + 22, 1); if (rc) return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(h->h_status + 12, h->status + 12, 12 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    double hb[36 + 6];
    CUDA_TRY(cudaMemcpyAsync(hb, h->BKK, 36 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(hb + 36, h->gr + PN, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    double kk = 0.0;
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) kk += hb[36 + r] * hb[6 * r + c] * hb[36 + c];
    *out = h->h_status[12] + h->h_status[13] + kk - 2.0 * h->h_status[22];
    return ASC_OK;
}

// [DEP] DoglegOptimizer::iterate + DoglegOptimizerImpl::Iterate(ONE_STEP_PER_ITERATION)
int dogleg_iterate(asc_handle* h, double* error, double* Delta, asc_iter_log* lg) {
    float ms = 0;
    h->ms_prepare = h->ms_coarse = h->ms_pcg = 0.0;
    int rc = dispatch_linearize(h, false, nullptr, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev[3], h->stream));
    int its = 0;
    rc = dispatch_solve(h, 0.0, &its);   // Gauss-Newton point in (vx, dl4)
    lg->pcg_iterations += its;
    if (rc) return rc;
    double d[4], gg4[4], gHg = 0.0;
    rc = step_dots(h, h->vx, h->dl4, d);           // n.g, n.n
    if (rc) return rc;
    rc = step_dots(h, h->gr, h->gl4, gg4);         // g.g
    if (rc) return rc;
    rc = dogleg_gHg(h, &gHg);
    if (rc) return rc;
    const double ng = d[0] + d[2], nn = d[1] + d[3], gg = gg4[0] + gg4[2];
    const double step = gg / gHg;                  // steepest-descent point u = step * g
    const double uu = step * step * gg, un = step * ng;
    const double f0 = *error;
    double f_new = f0, cu = 0.0, cn = 1.0;
    bool stay = true, moved = true;
    while (stay) {
        lg->inner_steps++;
        const double D2 = *Delta * *Delta;
        if (D2 < uu) { cu = std::sqrt(D2 / uu); cn = 0.0; }
        else if (D2 < nn) {   // ComputeBlend
            const double a = uu - 2. * un + nn, b = 2. * (un - uu), c = uu - D2, sq = std::sqrt(b * b - 4 * a * c);
            const double tau1 = (-b + sq) / (2. * a), tau2 = (-b - sq) / (2. * a);
            const double tau = (0.0 <= tau1 && tau1 <= 1.0) ? tau1 : tau2;
            cu = 1. - tau; cn = tau;
        } else { cu = 0.0; cn = 1.0; }
        rc = retract_to_try(h, cn, cu * step);
        if (rc) return rc;
        rc = dispatch_error(h, h->q_try, h->lm4_try, h->X_try, &f_new);
        if (rc) return rc;
        // M(delta) = f0 - delta.g + 0.5 delta^T H delta with H n = g (undamped Gauss-Newton solve)
        const double a = cu * step;   // delta = a g + cn n
        const double dg = a * gg + cn * ng;
        const double dHd = a * a * gHg + 2.0 * a * cn * gg + cn * cn * ng;   // g^T H n = g^T g, n^T H n = n^T g
        const double M_new = f0 - dg + 0.5 * dHd;
        const double rho = (std::fabs(f0 - f_new) < 1e-15 || std::fabs(f0 - M_new) < 1e-15) ? 0.5 : (f0 - f_new) / (f0 - M_new);
        if (rho >= 0.75) {
            const double dn = std::sqrt(cu * cu * uu + 2 * cu * cn * un + cn * cn * nn);
            *Delta = std::max(*Delta, 3.0 * dn); stay = false;
        } else if (rho >= 0.25) stay = false;
        else if (rho >= 0.0) { if (*Delta > 1e-5) *Delta *= 0.5; stay = false; }
        else {
            // f increased (or the trial error is not finite: every comparison above is false for a NaN rho).  Retry with a smaller
            // region; at the minimum region DoglegOptimizerImpl::Iterate zeroes dx_d and keeps f_error: the error never increases.
            if (*Delta > 1e-5) { *Delta *= 0.5; stay = true; }
            else { stay = false; moved = false; }
        }
    }
    if (moved) { accept_try(h); *error = f_new; lg->accepted = 1; }
    else { *error = f0; lg->accepted = 0; }
    CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
    CUDA_TRY(cudaEventSynchronize(h->ev[2]));
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]));
    lg->linearize_ms = ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[3], h->ev[2]));
    lg->solve_ms = ms;
    lg->prepare_ms = h->ms_prepare; lg->coarse_ms = h->ms_coarse; lg->pcg_ms = h->ms_pcg;
    return ASC_OK;
}

// (Re)creates this rank's exchange buffer when it must grow, and maps every peer's buffer through CUDA IPC.  Collective:
// every rank calls it from asc_set_problem with the same P and N.  Falls back to the NCCL path (p2p_ok = false) on any
// failure, agreed between the ranks through one all-reduce.
int setup_peer_exchange(asc_handle* h) {
    h->p2p_ok = false;
    if (h->world <= 1 || h->world > kMaxPeers || !g_nccl.AllGather || getenv("ASC_NO_P2P")) return ASC_OK;
    const size_t PN = (size_t)h->P * h->N;
    const size_t stride = (PN + 8 + 2 * kPartCols + 31) / 32 * 32;
    const size_t need = 256 + 2 * stride * sizeof(double);
    double bad = 0.0;
    if (!h->xbuf || need > h->xbuf_bytes) {
        int rc = allreduce(h, h->status, 1);   // barrier: no peer is still reading the buffer that is about to go away
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int r = 0; r < h->world; ++r)
            if (h->peer_map[r]) { cudaIpcCloseMemHandle(h->peer_map[r]); h->peer_map[r] = nullptr; }
        if (h->xbuf) CUDA_TRY(cudaFree(h->xbuf));
        h->xbuf = nullptr;
        h->xbuf_bytes = need + need / 4;
        CUDA_TRY(cudaMalloc((void**)&h->xbuf, h->xbuf_bytes));
        CUDA_TRY(cudaMemsetAsync(h->xbuf, 0, h->xbuf_bytes, h->stream));
        cudaIpcMemHandle_t mine;
        if (cudaIpcGetMemHandle(&mine, h->xbuf) != cudaSuccess) { cudaGetLastError(); bad = 1.0; std::memset(&mine, 0, sizeof mine); }
        char* d_handles = nullptr;
        const size_t hs = sizeof(cudaIpcMemHandle_t);
        CUDA_TRY(cudaMalloc((void**)&d_handles, hs * h->world));
        CUDA_TRY(cudaMemcpyAsync(d_handles + hs * h->rank, &mine, hs, cudaMemcpyHostToDevice, h->stream));
        if (g_nccl.AllGather(d_handles + hs * h->rank, d_handles, hs, ncclChar, h->comm, h->stream) != ncclSuccess) { cudaFree(d_handles); return fail(ASC_ERR_COMM, "ncclAllGather failed"); }
        std::vector<cudaIpcMemHandle_t> all(h->world);
        CUDA_TRY(cudaMemcpyAsync(all.data(), d_handles, hs * h->world, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        cudaFree(d_handles);
        for (int r = 0; r < h->world && bad == 0.0; ++r) {
            if (r == h->rank) continue;
            if (cudaIpcOpenMemHandle(&h->peer_map[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); h->peer_map[r] = nullptr; bad = 1.0; }
        }
    }
    // agree on the outcome
    CUDA_TRY(cudaMemcpyAsync(h->status + 1, &bad, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int rc = allreduce(h, h->status + 1, 1);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(&bad, h->status + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (bad != 0.0) return ASC_OK;   // NCCL path
    for (int r = 0; r < h->world; ++r) {
        char* base = r == h->rank ? h->xbuf : (char*)h->peer_map[r];
        h->pcm.peer_flags[r] = (unsigned*)base;
        h->pcm.x[r] = (double*)(base + 256);
    }
    h->pcm.ctr = (unsigned*)h->xbuf + 32;   // second half of the 256-byte header: local sequence counter (peers write entries 0..7 only)
    h->pcm.fail = h->flags + 8;
    h->pcm.stride = (long long)stride;
    h->pcm.row_off = (long long)(PN + 8);
    h->pcm.rank = h->rank; h->pcm.world = h->world;
    h->p2p_ok = true;
    return ASC_OK;
}

int check_ready(asc_handle* h, bool need_values) {
    if (!h) return fail(ASC_ERR_INVALID, "null handle");
    if (!h->chain_set || !h->cam_set || !h->problem_set) return fail(ASC_ERR_INVALID, "set chain, camera and problem first");
    if (need_values && !h->values_set) return fail(ASC_ERR_INVALID, "set values first");
    CUDA_TRY(cudaSetDevice(h->dev));
    return ASC_OK;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

#include "default_params.inl"

const char* asc_last_error(void) { return g_last_error.c_str(); }
const char* asc_version(void) { return "arm_slam_calib_b200 0.1 (sm_100a)"; }

int asc_create(const asc_params* params, asc_handle** out) {
    if (!params || !out) return fail(ASC_ERR_INVALID, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(ASC_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    int dev = params->device;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(ASC_ERR_INVALID, "device %d out of range", dev);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(ASC_ERR_CUDA, "device %d is sm_%d%d; this library ships sm_100a code only", dev, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(dev));
    asc_handle* h = new asc_handle();
    h->prm = *params;
    h->dev = dev;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_c, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_q, cudaEventDisableTiming));
    for (auto& ev : h->ev) CUDA_TRY(cudaEventCreate(&ev));
    for (auto& ev : h->ev_stage) CUDA_TRY(cudaEventCreate(&ev));
    CUDA_TRY(cudaMallocHost((void**)&h->h_status, 64 * sizeof(double)));
    CUDA_TRY(cudaMallocHost((void**)&h->h_flags, 16 * sizeof(int)));
    h->t_base.resize(12);
    asc::xf_to12(asc::xf_identity(), h->t_base.data());
    *out = h;
    return ASC_OK;
}

int asc_set_params(asc_handle* h, const asc_params* params) {
    if (!h || !params) return fail(ASC_ERR_INVALID, "null argument");
    const int dev = h->prm.device;
    h->prm = *params;
    h->prm.device = dev;
    if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }   // tolerances are baked into the graph
    return ASC_OK;
}

int asc_destroy(asc_handle* h) {
    if (!h) return ASC_OK;
    cudaSetDevice(h->dev);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->xbuf) {
        if (h->world > 1 && h->comm && h->status) { allreduce(h, h->status, 1); cudaStreamSynchronize(h->stream); }   // peers are done with our buffer
        for (int r = 0; r < kMaxPeers; ++r) if (h->peer_map[r]) cudaIpcCloseMemHandle(h->peer_map[r]);
        cudaFree(h->xbuf);
    }
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->pcg_graph) cudaGraphExecDestroy(h->pcg_graph);
    free_device(h);
    if (h->arena) cudaFree(h->arena);
    for (int k = 0; k < 2; ++k) if (h->d_has_link[k]) { cudaFree(h->d_has_link[k]); cudaFree(h->d_link_meas[k]); cudaFree(h->d_link_w[k]); }
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->ev_stage) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->prof_ev) cudaEventDestroy(ev);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_c) cudaEventDestroy(h->ev_c);
    if (h->ev_q) cudaEventDestroy(h->ev_q);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return ASC_OK;
}

int asc_set_chain(asc_handle* h, int32_t n, const double* t_p2j, const double* axis, const double* t_c2j, const uint8_t* continuous,
                  const double* t_base) {
    if (!h || !t_p2j || !axis || !t_c2j) return fail(ASC_ERR_INVALID, "null argument");
    if (!dof_supported(n)) return fail(ASC_ERR_INVALID, "no kernels compiled for %d DOF", n);
    if (h->problem_set && n != h->N) return fail(ASC_ERR_INVALID, "chain DOF changed after the problem was set");
    h->N = n;
    h->t_p2j.assign(t_p2j, t_p2j + 12 * n);
    h->axis.assign(axis, axis + 3 * n);
    h->t_c2j.assign(t_c2j, t_c2j + 12 * n);
    h->continuous.assign(n, 0);
    if (continuous) h->continuous.assign(continuous, continuous + n);
    if (t_base) h->t_base.assign(t_base, t_base + 12);
    h->chain_set = true;
    return ASC_OK;
}

int asc_set_camera(asc_handle* h, double fx, double fy, double cx, double cy) {
    if (!h) return fail(ASC_ERR_INVALID, "null handle");
    h->cm.fx = fx; h->cm.fy = fy; h->cm.cx = cx; h->cm.cy = cy;
    h->cam_set = true;
    return ASC_OK;
}

int asc_set_problem(asc_handle* h, int32_t P, int32_t L, int64_t M, const int32_t* pose_idx, const int32_t* lm_idx, const double* uv,
                    const double* enc, const uint8_t* has_enc, const double* lm_prior, const double* x_prior) {
    if (!h || !h->chain_set || !h->cam_set) return fail(ASC_ERR_INVALID, "set chain and camera first");
    if (P <= 0 || L < 0 || M < 0 || M >= (int64_t)0x7fffffff) return fail(ASC_ERR_INVALID, "bad sizes");
    if ((M > 0 && (!pose_idx || !lm_idx || !uv)) || !enc || (L > 0 && !lm_prior) || !x_prior) return fail(ASC_ERR_INVALID, "null argument");
    for (int64_t m = 0; m < M; ++m)
        if (pose_idx[m] < 0 || pose_idx[m] >= P || lm_idx[m] < 0 || lm_idx[m] >= L) return fail(ASC_ERR_INVALID, "observation %lld has an out-of-range index", (long long)m);
    static const bool timing = getenv("ASC_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[asc] set_problem %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    CUDA_TRY(cudaSetDevice(h->dev));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    lap("validate");
    free_device(h);
    lap("free_device");
    const int N = h->N;
    h->P = P; h->L = L; h->M = M;
    h->values_set = false; h->linearized = false; h->lin_lambda = -1.0;
    h->link_set[0] = h->link_set[1] = false;   // link factors belong to the packed graph: set them again after asc_set_problem
    h->coarse_last_iters = 0;
    if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }
    // constants
    const asc_params& pr = h->prm;
    h->cm.inv_sigma_px = 1.0 / pr.sigma_px;
    h->cm.k2 = pr.cauchy_k * pr.cauchy_k;
    h->cm.inv_sigma_lm = 1.0 / pr.sigma_landmark;
    std::memcpy(h->pc.x_prior, x_prior, sizeof h->pc.x_prior);
    for (int i = 0; i < 6; ++i) h->pc.inv_sigma_x[i] = 1.0 / pr.sigma_extrinsic[i];
    for (int i = 0; i < ASC_MAX_DOF; ++i) h->pc.inv_sigma_enc[i] = 1.0 / pr.sigma_encoder[i];
    h->pc.continuous_mask = 0;
    for (int i = 0; i < N; ++i) if (h->continuous[i]) h->pc.continuous_mask |= 1 << i;
    h->pc.use_dead_band = pr.use_dead_band; h->pc.dead_band = pr.dead_band; h->pc.chart = pr.pose3_chart;

    // ---- the two observation orders (stable counting sorts: list order kept inside each segment)
    std::vector<int> pose_ptr, lm_ptr;
    // staging buffers persist across calls (no page faults when a graph of similar size is re-packed)
    // (thread_local objects resolve per OpenMP worker, so the parallel loops below go through plain references)
    static thread_local std::vector<int> tl_i[9];
    static thread_local std::vector<double> tl_d[2];
    std::vector<int>&p_lm = tl_i[0], &p_pose = tl_i[1], &p_orig = tl_i[2], &pos_of = tl_i[3], &l_pose = tl_i[4], &l_perm = tl_i[5], &run_lo = tl_i[6],
                    &run_hi = tl_i[7], &dest1 = tl_i[8];
    std::vector<double>&p_uv = tl_d[0], &l_uv = tl_d[1];
    // The staging vectors are page-locked (cudaHostRegister, redone only when a vector had to grow): the 115 MB of index /
    // measurement uploads at the end of this function then run as plain DMA (~5 ms) instead of staged pageable copies (~12 ms).
    struct PinReg { void* ptr = nullptr; size_t bytes = 0; };
    static thread_local PinReg reg_i[9], reg_d[2];
    auto resize_pinned = [](auto& v, size_t n, PinReg& r) {
        if (n > v.capacity() && r.ptr) { cudaHostUnregister(r.ptr); r.ptr = nullptr; r.bytes = 0; }
        v.resize(n);
        const size_t bytes = v.capacity() * sizeof(v[0]);
        if (v.data() && bytes && (r.ptr != (void*)v.data() || r.bytes != bytes)) {
            if (r.ptr) cudaHostUnregister(r.ptr);
            r.ptr = nullptr; r.bytes = 0;
            if (cudaHostRegister(v.data(), bytes, cudaHostRegisterDefault) == cudaSuccess) { r.ptr = v.data(); r.bytes = bytes; }
            else cudaGetLastError();   // stays pageable
        }
    };
    static const bool no_pin = getenv("ASC_NO_PIN") != nullptr;
    if (no_pin) {
        p_lm.resize(M); p_pose.resize(M); p_orig.resize(M); pos_of.resize(M); l_pose.resize(M); l_perm.resize(M); dest1.resize(M);
        p_uv.resize(2 * M); l_uv.resize(2 * M);
    } else {
        resize_pinned(p_lm, M, reg_i[0]); resize_pinned(p_pose, M, reg_i[1]); resize_pinned(p_orig, M, reg_i[2]); resize_pinned(pos_of, M, reg_i[3]);
        resize_pinned(l_pose, M, reg_i[4]); resize_pinned(l_perm, M, reg_i[5]); dest1.resize(M);
        resize_pinned(p_uv, 2 * M, reg_d[0]); resize_pinned(l_uv, 2 * M, reg_d[1]);
    }
    // pose-major: stable by list order
    parallel_bucket_sort(M, P, [&](int64_t m) { return pose_idx[m]; }, pose_ptr, dest1.data());
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        const int k = dest1[m];
        p_lm[k] = lm_idx[m]; p_pose[k] = pose_idx[m]; p_orig[k] = (int)m;
        p_uv[2 * k] = uv[2 * m]; p_uv[2 * k + 1] = uv[2 * m + 1];
    }
    // landmark-major: inside a landmark, observations ordered by keyframe (then by list order); pos_of = pose-major -> landmark-major
    parallel_bucket_sort(M, std::max(L, 1), [&](int64_t kp) { return p_lm[kp]; }, lm_ptr, pos_of.data());
#pragma omp parallel for schedule(static)
    for (int64_t kp = 0; kp < M; ++kp) {
        const int k = pos_of[kp], m = p_orig[kp];
        l_pose[k] = p_pose[kp]; l_perm[k] = (int)kp;
        l_uv[2 * k] = uv[2 * m]; l_uv[2 * k + 1] = uv[2 * m + 1];
    }
    lap("two observation orders");
    // ---- preconditioner clusters: runs of consecutive keyframes that share landmarks
    const int gmax = std::max(1, std::min(pr.pcg_cluster_max, 128 / N));
    h->gmax = gmax;
    h->use_clusters = gmax > 1;
    struct PairRec { long long key; int a, b; };
    std::vector<PairRec> pairs;
    std::vector<int> inc_lo, inc_hi, inc_lm, inc_cl, dest_ptr, pos_lo, pos_up, pair_a, pair_b, cl_inc_ptr, cl_inc;
    std::vector<MlHostLevel> hl;   // multilevel hierarchy of the coarse operator (host patterns)
    std::vector<int> cl_ptr, pose_cluster(P, 0);
    std::vector<long long> cl_off;
    if (h->use_clusters) {
        std::vector<double> cnt(2 * (size_t)P, 0.0);   // [shared with previous | own count], as doubles for the all-reduce
        std::vector<int> mark(std::max(L, 1), -2);
        for (int p = 0; p < P; ++p) {
            int shared = 0;
            for (int k = pose_ptr[p]; k < pose_ptr[p + 1]; ++k) {
                const int j = p_lm[k];
                if (mark[j] == p - 1) ++shared;
            }
            for (int k = pose_ptr[p]; k < pose_ptr[p + 1]; ++k) mark[p_lm[k]] = p;
            cnt[p] = shared; cnt[P + p] = pose_ptr[p + 1] - pose_ptr[p];
        }
        if (h->world > 1) {   // every rank must build the same clusters: sum the shard-local counts
            double* d = nullptr;
            CUDA_TRY(cudaMalloc((void**)&d, cnt.size() * sizeof(double)));
            CUDA_TRY(cudaMemcpyAsync(d, cnt.data(), cnt.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            int rc2 = allreduce(h, d, cnt.size());
            if (rc2) { cudaFree(d); return rc2; }
            CUDA_TRY(cudaMemcpyAsync(cnt.data(), d, cnt.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            cudaFree(d);
        }
        lap("covisibility counts + cluster exchange");
        int size = 0;
        for (int p = 0; p < P; ++p) {
            const double mn = p > 0 ? std::min(cnt[P + p], cnt[P + p - 1]) : 0.0;
            const bool joined = p > 0 && size < gmax && mn > 0 && cnt[p] >= pr.pcg_cluster_covis * mn;
            if (!joined) { cl_ptr.push_back(p); size = 0; }
            ++size;
            pose_cluster[p] = (int)cl_ptr.size() - 1;
        }
        cl_ptr.push_back(P);
        h->nC = (int)cl_ptr.size() - 1;
        cl_off.assign(h->nC + 1, 0);
        h->nmax = 0;
        for (int c = 0; c < h->nC; ++c) {
            const long long n = (long long)(cl_ptr[c + 1] - cl_ptr[c]) * N;
            cl_off[c + 1] = cl_off[c] + n * n;
            h->nmax = std::max(h->nmax, (int)n);
        }
        h->scTotal = (size_t)cl_off[h->nC];
        if (no_pin) { run_lo.resize(M); run_hi.resize(M); }
        else { resize_pinned(run_lo, M, reg_i[6]); resize_pinned(run_hi, M, reg_i[7]); }
        h->dup_obs = false;
        h->n1 = h->nC * N;
        h->nz = h->n1 + 6;
        h->use_coarse = pr.pcg_coarse_levels > 0 && h->nC > 1;
        {   // incidences = (landmark, cluster) runs of the landmark-major list; built per thread over landmark ranges, then concatenated
            const int T = std::max(1, std::min(omp_get_max_threads(), 16));
            struct Local { std::vector<int> lo, hi, lm, cl; std::vector<PairRec> pairs; bool dup = false; };
            std::vector<Local> loc(T);
#pragma omp parallel num_threads(T)
            {
                // slices follow the team size actually delivered (it may be smaller than T); unused Local entries stay empty
                const int t = omp_get_thread_num(), Tn = omp_get_num_threads();
                Local& Lc = loc[t];
                const int j0 = (int)((int64_t)L * t / Tn), j1 = (int)((int64_t)L * (t + 1) / Tn);
                for (int j = j0; j < j1; ++j) {
                    int a = lm_ptr[j];
                    const size_t first_inc = Lc.lo.size();
                    while (a < lm_ptr[j + 1]) {
                        int b = a + 1;
                        const int cid = pose_cluster[l_pose[a]];
                        while (b < lm_ptr[j + 1] && pose_cluster[l_pose[b]] == cid) ++b;
                        for (int kk = a; kk < b; ++kk) { run_lo[l_perm[kk]] = a; run_hi[l_perm[kk]] = b; }
                        Lc.lo.push_back(a); Lc.hi.push_back(b); Lc.lm.push_back(j); Lc.cl.push_back(cid);
                        for (int kk = a + 1; kk < b; ++kk) if (l_pose[kk] == l_pose[kk - 1]) Lc.dup = true;
                        a = b;
                    }
                    if (h->use_coarse)   // block updates of E = Z^T S Z: lower triangle of cluster pairs, in landmark order (local incidence ids)
                        for (size_t ia = first_inc; ia < Lc.lo.size(); ++ia)
                            for (size_t ib = first_inc; ib < Lc.lo.size(); ++ib)
                                if (Lc.cl[ia] >= Lc.cl[ib]) Lc.pairs.push_back({(long long)Lc.cl[ia] * h->nC + Lc.cl[ib], (int)ia, (int)ib});
                }
            }
            size_t ninc = 0, npair = 0;
            std::vector<size_t> inc_off(T + 1, 0), pair_off(T + 1, 0);
            for (int t = 0; t < T; ++t) { inc_off[t] = ninc; pair_off[t] = npair; ninc += loc[t].lo.size(); npair += loc[t].pairs.size(); h->dup_obs = h->dup_obs || loc[t].dup; }
            inc_lo.resize(ninc); inc_hi.resize(ninc); inc_lm.resize(ninc); inc_cl.resize(ninc); pairs.resize(npair);
#pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int t = 0; t < T; ++t) {
                const Local& Lc = loc[t];
                std::copy(Lc.lo.begin(), Lc.lo.end(), inc_lo.begin() + inc_off[t]); std::copy(Lc.hi.begin(), Lc.hi.end(), inc_hi.begin() + inc_off[t]);
                std::copy(Lc.lm.begin(), Lc.lm.end(), inc_lm.begin() + inc_off[t]); std::copy(Lc.cl.begin(), Lc.cl.end(), inc_cl.begin() + inc_off[t]);
                for (size_t i = 0; i < Lc.pairs.size(); ++i) {
                    PairRec r = Lc.pairs[i];
                    r.a += (int)inc_off[t]; r.b += (int)inc_off[t];
                    pairs[pair_off[t] + i] = r;
                }
            }
        }
        lap("runs + incidences + pairs");
        std::vector<long long> dest_key;   // unique (ca, cb) keys of the lower triangle, ascending; after the union: of ALL ranks
        std::vector<double> dest_w;        // shared-landmark pair count per key (summed over ranks)
        if (h->use_coarse) {
            {   // stable LSD radix sort of the pairs by (cluster a, cluster b): two parallel counting passes
                const int nCl = h->nC;
                std::vector<PairRec> tmp(pairs.size());
                std::vector<int> bptr, dst(pairs.size());
                for (int pass = 0; pass < 2; ++pass) {
                    parallel_bucket_sort((int64_t)pairs.size(), nCl, [&](int64_t i) { return pass == 0 ? (int)(pairs[i].key % nCl) : (int)(pairs[i].key / nCl); }, bptr, dst.data());
#pragma omp parallel for schedule(static)
                    for (int64_t i = 0; i < (int64_t)pairs.size(); ++i) tmp[dst[i]] = pairs[i];
                    pairs.swap(tmp);
                }
            }
            pair_a.resize(pairs.size()); pair_b.resize(pairs.size());
            std::vector<int> loc_ptr;   // pair range of each LOCAL key
            for (size_t i = 0; i < pairs.size(); ++i) {
                if (i == 0 || pairs[i].key != pairs[i - 1].key) { loc_ptr.push_back((int)i); dest_key.push_back(pairs[i].key); }
                pair_a[i] = pairs[i].a; pair_b[i] = pairs[i].b;
            }
            loc_ptr.push_back((int)pairs.size());
            std::vector<long long> loc_key = dest_key;
            dest_w.resize(dest_key.size());
            for (size_t d = 0; d < dest_key.size(); ++d) dest_w[d] = loc_ptr[d + 1] - loc_ptr[d];
            // every cluster owns a diagonal block even without observations on this rank (rank 0 adds B_pp + lambda I to it)
            {
                std::vector<long long> k2; std::vector<double> w2;
                k2.reserve(dest_key.size() + h->nC); w2.reserve(dest_key.size() + h->nC);
                size_t q = 0;
                for (int c = 0; c < h->nC; ++c) {
                    const long long dk = (long long)c * h->nC + c;
                    while (q < dest_key.size() && dest_key[q] < dk) { k2.push_back(dest_key[q]); w2.push_back(dest_w[q]); ++q; }
                    if (q < dest_key.size() && dest_key[q] == dk) { k2.push_back(dk); w2.push_back(dest_w[q]); ++q; }
                    else { k2.push_back(dk); w2.push_back(0.0); }
                }
                while (q < dest_key.size()) { k2.push_back(dest_key[q]); w2.push_back(dest_w[q]); ++q; }
                dest_key.swap(k2); dest_w.swap(w2);
            }
            if (h->world > 1) {   // union of the block patterns of all shards (every rank must hold the same hierarchy)
                int rcu = union_block_keys(h, dest_key, dest_w);
                if (rcu) return rcu;
            }
            // pair ranges against the (global) key list: empty for blocks this rank has no landmark in (the local keys are a
            // subsequence of the global ones, and the pairs are sorted by key)
            dest_ptr.assign(dest_key.size() + 1, 0);
            {
                size_t q = 0;
                int pos = 0;
                for (size_t d = 0; d < dest_key.size(); ++d) {
                    dest_ptr[d] = pos;
                    if (q < loc_key.size() && loc_key[q] == dest_key[d]) { pos = loc_ptr[q + 1]; ++q; }
                }
                dest_ptr[dest_key.size()] = pos;
            }
            h->nDest = (int)dest_key.size();
        }
        h->nInc = (int)inc_lo.size();
        cl_inc_ptr.assign(h->nC + 1, 0);
        for (int c : inc_cl) cl_inc_ptr[c + 1]++;
        for (int c = 0; c < h->nC; ++c) cl_inc_ptr[c + 1] += cl_inc_ptr[c];
        cl_inc.assign(inc_cl.size(), 0);
        {
            std::vector<int> f(cl_inc_ptr.begin(), cl_inc_ptr.end() - 1);
            for (size_t i = 0; i < inc_cl.size(); ++i) cl_inc[f[inc_cl[i]]++] = (int)i;
        }
        if (h->use_coarse) {   // ---- multilevel hierarchy (coarse_ml.cuh)
            const bool fixed_gs = pr.pcg_ml_agg < 0;   // negative: exactly |pcg_ml_agg| nodes per aggregate, no enlargement (deep hierarchies for tests)
            const int gs = std::max(2, std::abs(pr.pcg_ml_agg)), top_dim = std::min(kTopMaxDim, std::max(6 * N, pr.pcg_ml_top_dim)), max_levels = std::max(1, pr.pcg_coarse_levels);
            hl.clear();
            hl.emplace_back();
            MlHostLevel& L0 = hl[0];
            L0.n = h->nC;
            {   // BSR of level 1 from the lower-triangle keys
                std::vector<int> cntr(h->nC + 1, 0);
                for (long long k : dest_key) { const int a = (int)(k / h->nC), b = (int)(k % h->nC); cntr[a + 1]++; if (a != b) cntr[b + 1]++; }
                for (int c = 0; c < h->nC; ++c) cntr[c + 1] += cntr[c];
                L0.row_ptr = cntr;
                L0.col.assign(cntr[h->nC], 0); L0.w.assign(cntr[h->nC], 0.0); L0.diag_pos.assign(h->nC, -1);
                pos_lo.assign(dest_key.size(), 0); pos_up.assign(dest_key.size(), 0);
                // keys ascend by (a, b) with b <= a.  Row r holds first its lower part (b < r, from the keys (r, b), already in
                // ascending b), then the diagonal, then its upper part (columns a > r from the keys (a, r), ascending a).
                std::vector<int> nlow(h->nC, 0), fill_lo(h->nC, 0), fill_up(h->nC, 0);
                for (long long k : dest_key) { const int a = (int)(k / h->nC), b = (int)(k % h->nC); if (a != b) nlow[a]++; }
                for (size_t d = 0; d < dest_key.size(); ++d) {
                    const int a = (int)(dest_key[d] / h->nC), b = (int)(dest_key[d] % h->nC);
                    if (a == b) { const int pos = cntr[a] + nlow[a]; L0.col[pos] = a; L0.w[pos] = dest_w[d]; L0.diag_pos[a] = pos; pos_lo[d] = pos; pos_up[d] = pos; }
                    else {
                        const int plo = cntr[a] + fill_lo[a]++;
                        const int pup = cntr[b] + nlow[b] + 1 + fill_up[b]++;
                        L0.col[plo] = b; L0.w[plo] = dest_w[d]; L0.col[pup] = a; L0.w[pup] = dest_w[d];
                        pos_lo[d] = plo; pos_up[d] = pup;
                    }
                }
            }
            while ((int)hl.size() <= max_levels && hl.back().n * N > top_dim) {
                MlHostLevel& L = hl.back();
                int na = 0;
                // Aggregates as large as it takes to reach the dense top level in ONE step when <= 64 nodes per aggregate can do it: the
                // smoothed level-0 sweeps carry the convergence (prototype: 106 -> 111 -> 114 PCG iterations for aggregates of 8 / 16 /
                // 32..64), while every extra level adds ~9 latency-bound phases to the V-cycle of each PCG iteration.
                const long long need = (14LL * L.n * N + 10LL * top_dim - 1) / (10LL * top_dim);
                const int gs_l = fixed_gs ? gs : (int)std::min<long long>(64, std::max<long long>(gs, need));
                ml_aggregate(L, gs_l, L.agg, na);
                if (na >= L.n) break;   // no coarsening possible (no couplings): stop here
                L.n_next = na;
                hl.emplace_back();
                ml_coarsen_pattern(hl[hl.size() - 2], hl.back());
            }
            h->ntop = hl.back().n * N;
            if (h->ntop > kTopMaxDim) {   // the graph did not coarsen far enough (e.g. no couplings between clusters): cluster-Jacobi alone
                h->use_coarse = false;
                hl.clear();
            }
        }
    } else { h->nC = 0; h->nmax = 0; h->scTotal = 0; h->use_coarse = false; h->n1 = 0; h->nz = 0; }
    lap("pair sort + coarse hierarchy");

    const size_t PN = (size_t)P * N, Mz = (size_t)M;
    h->nCtaPose = cdiv(P, kWarpsPerCta);
    h->nBlkLm = std::max(1, cdiv(L, 128));
    h->nBlkLmK2 = std::max(1, cdiv(L, kLmPerBlock));
    h->nBlkB = std::max(1, cdiv(L, kPassBLmPerBlock));
    int rc = ASC_OK;
    std::vector<ArenaReq> reqs;
#define ALLOC(ptr, count) reqs.push_back(ArenaReq{(void**)&h->ptr, sizeof(*h->ptr) * std::max<size_t>((size_t)(count), 1)})
    ALLOC(pose_ptr, P + 1); ALLOC(p_lm, Mz); ALLOC(p_pose, Mz); ALLOC(p_orig, Mz); ALLOC(p_uv, Mz);
    ALLOC(lm_ptr, L + 1); ALLOC(l_pose, Mz); ALLOC(l_perm, Mz); ALLOC(l_uv, Mz); ALLOC(p_inv, Mz);
    ALLOC(enc, PN); ALLOC(lm_prior, (size_t)L * 3);
    if (has_enc) { ALLOC(has_enc, P); } else { h->has_enc = nullptr; }
    ALLOC(q, PN); ALLOC(q_try, PN); ALLOC(lm4, (size_t)L * 4); ALLOC(lm4_try, (size_t)L * 4); ALLOC(X, 12); ALLOC(X_try, 12);
    ALLOC(q_save, PN); ALLOC(lm4_save, (size_t)L * 4); ALLOC(X_save, 12);
    ALLOC(cam, (size_t)P * 12); ALLOC(chain, PN * 6); ALLOC(lml, (size_t)L * 4);
    ALLOC(J, (size_t)(N + 3) * Mz); ALLOC(b_dbg, 1);
    ALLOC(poseblk, PN * N + PN * 6 + 36 + PN + 6 + 8);
    ALLOC(C, (size_t)L * 6); ALLOC(gl, (size_t)L * 3); ALLOC(WK, (size_t)L * 18); ALLOC(Cinv, (size_t)L * 6);
    ALLOC(cg4, (size_t)L * 4); ALLOC(u4, (size_t)L * 4); ALLOC(dl4, (size_t)L * 4); ALLOC(gl4, (size_t)L * 4); ALLOC(v4, Mz * 4);
    h->regionA = std::max(PN * N, h->scTotal);
    ALLOC(precblk, h->regionA + PN + kPartCols);
    if (h->use_clusters) {
        ALLOC(cl_ptr, h->nC + 1); ALLOC(cl_off, h->nC + 1); ALLOC(pose_cluster, P); ALLOC(p_run_lo, Mz); ALLOC(p_run_hi, Mz);
        ALLOC(Wm, Mz * 3 * N); ALLOC(Mc, h->scTotal);
        ALLOC(inc_lo, h->nInc); ALLOC(inc_hi, h->nInc); ALLOC(inc_lm, h->nInc); ALLOC(cl_inc_ptr, h->nC + 1); ALLOC(cl_inc, h->nInc);
    }
    h->ml.clear();
    h->ml_out = nullptr;
    if (h->use_coarse) {
        const size_t ni = h->nInc, nd = h->nDest, np = pair_a.size(), nz = h->nz, NN = (size_t)N * N;
        ALLOC(dest_ptr, nd + 1); ALLOC(pos_lo, nd); ALLOC(pos_up, nd);
        ALLOC(pair_a, np); ALLOC(pair_b, np);
        ALLOC(Vinc, ni * 3 * N); ALLOC(Tinc, ni * 3 * N); ALLOC(rc, nz); ALLOC(zc, nz); ALLOC(ryc, nz);
        ALLOC(pose_agg, P);
        h->ml.resize(hl.size());   // fixed size from here on: the arena requests keep pointers into it
        h->ml_degree = std::max(1, std::min(pr.pcg_ml_degree, 12));
        {
            static const double frac_of_degree[13] = {4, 4, 4, 8, 10, 15, 20, 25, 30, 35, 40, 45, 50};   // Chebyshev interval [hi / frac, hi]
            h->ml_frac = frac_of_degree[h->ml_degree];
        }
        ALLOC(mlblk, (size_t)hl[0].col.size() * NN + (size_t)h->n1 * 6);
        size_t max_grid = 1;
        for (size_t l = 0; l < hl.size(); ++l) {
            const MlHostLevel& H = hl[l];
            asc_handle::MlLevel& D = h->ml[l];
            D.n = H.n; D.nnzb = (int)H.col.size(); D.n_next = H.n_next;
            const bool top = l + 1 == hl.size();
            reqs.push_back(ArenaReq{(void**)&D.row_ptr, sizeof(int) * (size_t)(H.n + 1)});
            reqs.push_back(ArenaReq{(void**)&D.col, sizeof(int) * std::max<size_t>(H.col.size(), 1)});
            reqs.push_back(ArenaReq{(void**)&D.diag_pos, sizeof(int) * (size_t)H.n});
            reqs.push_back(ArenaReq{(void**)&D.bpos, sizeof(int2) * std::max<size_t>(H.col.size(), 1)});
            if (l > 0) reqs.push_back(ArenaReq{(void**)&D.Eb, sizeof(double) * H.col.size() * NN});
            reqs.push_back(ArenaReq{(void**)&D.r, sizeof(double) * (size_t)H.n * N});
            if (!top) {
                reqs.push_back(ArenaReq{(void**)&D.agg, sizeof(int) * (size_t)H.n});
                reqs.push_back(ArenaReq{(void**)&D.mptr, sizeof(int) * (size_t)(H.n_next + 1)});
                reqs.push_back(ArenaReq{(void**)&D.midx, sizeof(int) * (size_t)H.n});
                reqs.push_back(ArenaReq{(void**)&D.cptr, sizeof(int) * H.cptr.size()});
                reqs.push_back(ArenaReq{(void**)&D.cidx, sizeof(int) * std::max<size_t>(H.cidx.size(), 1)});
                reqs.push_back(ArenaReq{(void**)&D.Dinv, sizeof(double) * (size_t)H.n * NN});
                reqs.push_back(ArenaReq{(void**)&D.lmax, sizeof(double) * 32});   // [lmax | pad | c1_k, c2_k per Chebyshev step]
                reqs.push_back(ArenaReq{(void**)&D.za, sizeof(double) * (size_t)H.n * N});
                reqs.push_back(ArenaReq{(void**)&D.zb, sizeof(double) * (size_t)H.n * N});
                reqs.push_back(ArenaReq{(void**)&D.d, sizeof(double) * (size_t)H.n * N});
                reqs.push_back(ArenaReq{(void**)&D.rs, sizeof(double) * (size_t)H.n * N});
                max_grid = std::max<size_t>(max_grid, (size_t)cdiv(H.n, 4));
            }
        }
        ALLOC(Yb, (size_t)h->n1 * 6); ALLOC(Ski, 36); ALLOC(mlcol, h->n1);
        ALLOC(Atop, (size_t)h->ntop * h->ntop); ALLOC(ztop, h->ntop); ALLOC(mlpart, 2 * max_grid + 2); ALLOC(mlbar, 4); ALLOC(mlpiv, 2 * (size_t)h->ntop);
    }
    ALLOC(Minv, PN * N); ALLOC(MinvK, 36); ALLOC(rhs, PN + 8);
    ALLOC(vx, PN + 8); ALLOC(vr, PN + 8); ALLOC(vz, PN + 8); ALLOC(vp, PN + 8);
    ALLOC(ybuf, PN + 8 + 2 * kPartCols);
    const size_t nStep = (size_t)cdiv(std::max<int64_t>((int64_t)PN + 6, (int64_t)L * 4), 256);
    ALLOC(partA, std::max<size_t>(std::max<size_t>((size_t)P + 1024, nStep), (size_t)h->nC) * kPartCols); ALLOC(partC, (size_t)h->nCtaPose * kPartCols);
    ALLOC(partB, (size_t)h->nBlkB * kPartCols); ALLOC(partL, (size_t)h->nBlkLm * kPartCols); ALLOC(partR, 64 * kPartCols);
    h->nErrBlk = cdiv(M, 256) + cdiv((int64_t)P + L, 256) + h->nBlkLm + 8;
    ALLOC(errblk, h->nErrBlk); ALLOC(lmerr, h->nBlkLmK2);
    ALLOC(scal, 16); ALLOC(status, 64); ALLOC(flags, 16); ALLOC(cheir, 1); ALLOC(SKK, 36); ALLOC(stage3, (size_t)L * 3);
#undef ALLOC
    rc = commit_arena(h, reqs);
    if (rc) return rc;
    h->Bpp = h->poseblk; h->BpK = h->Bpp + PN * N; h->BKK = h->BpK + PN * 6; h->gr = h->BKK + 36; h->err_lin = h->gr + PN + 6;
    h->Dloc = h->precblk; h->rloc = h->Dloc + h->regionA; h->krow = h->rloc + PN;
    rc = setup_peer_exchange(h);
    if (rc) return rc;
    if (h->use_coarse) {
        h->ml[0].Eb = h->mlblk;
        h->EK = h->mlblk + hl[0].col.size() * (size_t)N * N;
        CUDA_TRY(cudaFuncSetAttribute(ml_top_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * kTopMaxDim * (kTopThreads / 32 + 1))));
    }
    lap("cudaMalloc + memset");
#define UP(dst, src, bytes) CUDA_TRY(cudaMemcpyAsync(h->dst, src, (bytes), cudaMemcpyHostToDevice, h->stream))
    UP(pose_ptr, pose_ptr.data(), sizeof(int) * (P + 1)); UP(lm_ptr, lm_ptr.data(), sizeof(int) * (L + 1));
    if (M) {
        UP(p_lm, p_lm.data(), sizeof(int) * Mz); UP(p_pose, p_pose.data(), sizeof(int) * Mz); UP(p_orig, p_orig.data(), sizeof(int) * Mz);
        UP(p_uv, p_uv.data(), sizeof(double) * 2 * Mz); UP(l_pose, l_pose.data(), sizeof(int) * Mz); UP(l_perm, l_perm.data(), sizeof(int) * Mz);
        UP(l_uv, l_uv.data(), sizeof(double) * 2 * Mz); UP(p_inv, pos_of.data(), sizeof(int) * Mz);
    }
    if (h->use_clusters) {
        UP(cl_ptr, cl_ptr.data(), sizeof(int) * (h->nC + 1)); UP(cl_off, cl_off.data(), sizeof(long long) * (h->nC + 1));
        UP(pose_cluster, pose_cluster.data(), sizeof(int) * P);
        if (M) { UP(p_run_lo, run_lo.data(), sizeof(int) * Mz); UP(p_run_hi, run_hi.data(), sizeof(int) * Mz); }
    }
    if (h->use_clusters) {
        UP(cl_inc_ptr, cl_inc_ptr.data(), sizeof(int) * (h->nC + 1));
        if (h->nInc) {
            UP(inc_lo, inc_lo.data(), sizeof(int) * h->nInc); UP(inc_hi, inc_hi.data(), sizeof(int) * h->nInc); UP(inc_lm, inc_lm.data(), sizeof(int) * h->nInc);
            UP(cl_inc, cl_inc.data(), sizeof(int) * h->nInc);
        }
    }
    std::vector<std::vector<int2>> bpos_stage(hl.size());   // alive until the stream synchronises below
    if (h->use_coarse) {
        UP(pose_agg, pose_cluster.data(), sizeof(int) * P);
        UP(dest_ptr, dest_ptr.data(), sizeof(int) * (h->nDest + 1)); UP(pos_lo, pos_lo.data(), sizeof(int) * h->nDest); UP(pos_up, pos_up.data(), sizeof(int) * h->nDest);
        if (!pair_a.empty()) { UP(pair_a, pair_a.data(), sizeof(int) * pair_a.size()); UP(pair_b, pair_b.data(), sizeof(int) * pair_b.size()); }
        for (size_t l = 0; l < hl.size(); ++l) {
            const MlHostLevel& H = hl[l];
            asc_handle::MlLevel& D = h->ml[l];
            CUDA_TRY(cudaMemcpyAsync(D.row_ptr, H.row_ptr.data(), sizeof(int) * (H.n + 1), cudaMemcpyHostToDevice, h->stream));
            if (!H.col.empty()) CUDA_TRY(cudaMemcpyAsync(D.col, H.col.data(), sizeof(int) * H.col.size(), cudaMemcpyHostToDevice, h->stream));
            CUDA_TRY(cudaMemcpyAsync(D.diag_pos, H.diag_pos.data(), sizeof(int) * H.n, cudaMemcpyHostToDevice, h->stream));
            {   // element positions of the transposed block rows
                bpos_stage[l].resize(H.col.size());
                for (int i = 0; i < H.n; ++i)
                    for (int k = H.row_ptr[i]; k < H.row_ptr[i + 1]; ++k) bpos_stage[l][k] = make_int2(H.row_ptr[i] * N * N + (k - H.row_ptr[i]), H.row_ptr[i + 1] - H.row_ptr[i]);
                if (!H.col.empty()) CUDA_TRY(cudaMemcpyAsync(D.bpos, bpos_stage[l].data(), sizeof(int2) * H.col.size(), cudaMemcpyHostToDevice, h->stream));
            }
            if (l + 1 < hl.size()) {
                CUDA_TRY(cudaMemcpyAsync(D.agg, H.agg.data(), sizeof(int) * H.n, cudaMemcpyHostToDevice, h->stream));
                CUDA_TRY(cudaMemcpyAsync(D.mptr, H.mptr.data(), sizeof(int) * (H.n_next + 1), cudaMemcpyHostToDevice, h->stream));
                CUDA_TRY(cudaMemcpyAsync(D.midx, H.midx.data(), sizeof(int) * H.n, cudaMemcpyHostToDevice, h->stream));
                CUDA_TRY(cudaMemcpyAsync(D.cptr, H.cptr.data(), sizeof(int) * H.cptr.size(), cudaMemcpyHostToDevice, h->stream));
                if (!H.cidx.empty()) CUDA_TRY(cudaMemcpyAsync(D.cidx, H.cidx.data(), sizeof(int) * H.cidx.size(), cudaMemcpyHostToDevice, h->stream));
            }
        }
    }
    UP(enc, enc, sizeof(double) * PN);
    if (L) UP(lm_prior, lm_prior, sizeof(double) * 3 * L);
    if (has_enc) UP(has_enc, has_enc, P);
#undef UP
    CUDA_TRY(cudaStreamSynchronize(h->stream));   // host staging vectors go out of scope
    lap("uploads");
    h->problem_set = true;
    return ASC_OK;
}

static int set_links(asc_handle* h, int kind, const uint8_t* has, const double* meas, const double* sigma) {
    if (!h || !h->problem_set) return fail(ASC_ERR_INVALID, "set the problem first");
    CUDA_TRY(cudaSetDevice(h->dev));
    h->link_set[kind] = false;
    h->linearized = false;
    if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }   // the link table is a kernel argument of the captured loop
    if (!has) return ASC_OK;
    if (!meas || !sigma) return fail(ASC_ERR_INVALID, "null argument");
    const int P = h->P, N = h->N;
    for (int i = 0; i < N; ++i) if (!(sigma[i] > 0)) return fail(ASC_ERR_INVALID, "link sigma must be positive");
    std::vector<uint8_t> hd((size_t)P + 1, 0);
    for (int p = 1; p < P; ++p) hd[p] = has[p] ? 1 : 0;   // a link needs a predecessor; [P] stays 0 (kernels read p + 1)
    std::vector<double> w(N);
    for (int i = 0; i < N; ++i) { h->pc.inv_sigma_link[kind][i] = 1.0 / sigma[i]; w[i] = h->pc.inv_sigma_link[kind][i] * h->pc.inv_sigma_link[kind][i]; }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->d_has_link[kind]) { cudaFree(h->d_has_link[kind]); cudaFree(h->d_link_meas[kind]); cudaFree(h->d_link_w[kind]); h->d_has_link[kind] = nullptr; }
    CUDA_TRY(cudaMalloc((void**)&h->d_has_link[kind], (size_t)P + 1));
    CUDA_TRY(cudaMalloc((void**)&h->d_link_meas[kind], sizeof(double) * (size_t)P * N));
    CUDA_TRY(cudaMalloc((void**)&h->d_link_w[kind], sizeof(double) * N));
    CUDA_TRY(cudaMemcpyAsync(h->d_has_link[kind], hd.data(), (size_t)P + 1, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_link_meas[kind], meas, sizeof(double) * (size_t)P * N, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_link_w[kind], w.data(), sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->link_set[kind] = true;
    return ASC_OK;
}

int asc_set_drift(asc_handle* h, const uint8_t* has_drift, const double* drift_meas, const double* sigma_drift) {
    return set_links(h, 0, has_drift, drift_meas, sigma_drift);
}

int asc_set_velocity(asc_handle* h, const uint8_t* has_velocity, const double* velocity_meas, const double* sigma_velocity) {
    return set_links(h, 1, has_velocity, velocity_meas, sigma_velocity);
}

int asc_set_values(asc_handle* h, const double* q, const double* l, const double* x) {
    if (!h || !h->problem_set) return fail(ASC_ERR_INVALID, "set the problem first");
    if (!q || (h->L > 0 && !l) || !x) return fail(ASC_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->dev));
    CUDA_TRY(cudaMemcpyAsync(h->q, q, sizeof(double) * h->P * h->N, cudaMemcpyHostToDevice, h->stream));
    if (h->L) {   // one contiguous copy + a pad kernel (a pitched 24->32 B copy from pageable memory takes ~13 ms for 90k rows)
        CUDA_TRY(cudaMemcpyAsync(h->stage3, l, sizeof(double) * 3 * h->L, cudaMemcpyHostToDevice, h->stream));
        pad3to4_kernel<<<cdiv((int64_t)h->L * 4, 256), 256, 0, h->stream>>>(h->L, h->stage3, h->lm4);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(h->X, x, sizeof(double) * 12, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->values_set = true;
    h->linearized = false;
    return ASC_OK;
}

// Incremental mode: the drivers call OptimizeStep every one or two frames on a GROWING graph with the previous
// estimate as the initial values (sim_genrobot.cpp:97-104, sim.cpp; ArmSlamCalib::SimulationStep adds the new configuration, its
// encoder factor and the new projection factors, ArmSlamCalib.cpp:307-327).  asc_append grows the packed problem in place: the
// existing observations (original order), encoder readings, landmark priors and the CURRENT ESTIMATE are taken from the device, the
// new keyframes / landmarks / observations are appended, and the graph is re-packed -- so a following asc_optimize starts from the
// previous solution (warm start) and the caller never re-sends what the library already holds.
int asc_append(asc_handle* h, int32_t n_new_poses, int32_t n_new_landmarks, int64_t n_new_obs, const int32_t* pose_idx, const int32_t* lm_idx,
               const double* uv, const double* enc_new, const uint8_t* has_enc_new, const double* lm_prior_new, const double* q_new, const double* l_new) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (n_new_poses < 0 || n_new_landmarks < 0 || n_new_obs < 0) return fail(ASC_ERR_INVALID, "bad sizes");
    if ((n_new_obs > 0 && (!pose_idx || !lm_idx || !uv)) || (n_new_poses > 0 && (!enc_new || !q_new)) || (n_new_landmarks > 0 && (!lm_prior_new || !l_new)))
        return fail(ASC_ERR_INVALID, "null argument");
    const int N = h->N;
    const size_t P0 = h->P, L0 = h->L, M0 = (size_t)h->M, P1 = P0 + n_new_poses, L1 = L0 + n_new_landmarks, M1 = M0 + (size_t)n_new_obs;
    // what the device holds, back in the caller's order
    std::vector<int> d_lm(M0), d_pose(M0), d_orig(M0);
    std::vector<double> d_uv(2 * M0), enc(P1 * N), prior(3 * L1), q(P1 * N), l(3 * L1), x(12);
    std::vector<uint8_t> has(P1, 1);
    bool any_has = has_enc_new != nullptr || h->has_enc != nullptr;
    if (M0) {
        CUDA_TRY(cudaMemcpyAsync(d_lm.data(), h->p_lm, sizeof(int) * M0, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemcpyAsync(d_pose.data(), h->p_pose, sizeof(int) * M0, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemcpyAsync(d_orig.data(), h->p_orig, sizeof(int) * M0, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemcpyAsync(d_uv.data(), h->p_uv, sizeof(double) * 2 * M0, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(enc.data(), h->enc, sizeof(double) * P0 * N, cudaMemcpyDeviceToHost, h->stream));
    if (L0) CUDA_TRY(cudaMemcpyAsync(prior.data(), h->lm_prior, sizeof(double) * 3 * L0, cudaMemcpyDeviceToHost, h->stream));
    if (h->has_enc) CUDA_TRY(cudaMemcpyAsync(has.data(), h->has_enc, P0, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    rc = asc_get_values(h, q.data(), L0 ? l.data() : nullptr, x.data());
    if (rc) return rc;
    std::vector<int32_t> pi(M1), li(M1);
    std::vector<double> uvs(2 * M1);
    for (size_t k = 0; k < M0; ++k) {
        const size_t m = (size_t)d_orig[k];
        pi[m] = d_pose[k]; li[m] = d_lm[k]; uvs[2 * m] = d_uv[2 * k]; uvs[2 * m + 1] = d_uv[2 * k + 1];
    }
    for (size_t m = 0; m < (size_t)n_new_obs; ++m) { pi[M0 + m] = pose_idx[m]; li[M0 + m] = lm_idx[m]; uvs[2 * (M0 + m)] = uv[2 * m]; uvs[2 * (M0 + m) + 1] = uv[2 * m + 1]; }
    if (n_new_poses) {
        std::memcpy(enc.data() + P0 * N, enc_new, sizeof(double) * (size_t)n_new_poses * N);
        std::memcpy(q.data() + P0 * N, q_new, sizeof(double) * (size_t)n_new_poses * N);
        for (int p = 0; p < n_new_poses; ++p) has[P0 + p] = has_enc_new ? has_enc_new[p] : 1;
    }
    if (n_new_landmarks) {
        std::memcpy(prior.data() + 3 * L0, lm_prior_new, sizeof(double) * 3 * (size_t)n_new_landmarks);
        std::memcpy(l.data() + 3 * L0, l_new, sizeof(double) * 3 * (size_t)n_new_landmarks);
    }
    double x_prior[12];
    std::memcpy(x_prior, h->pc.x_prior, sizeof x_prior);
    rc = asc_set_problem(h, (int32_t)P1, (int32_t)L1, (int64_t)M1, pi.data(), li.data(), uvs.data(), enc.data(), any_has ? has.data() : nullptr, prior.data(), x_prior);
    if (rc) return rc;
    return asc_set_values(h, q.data(), l.data(), x.data());
}

int asc_get_values(asc_handle* h, double* q, double* l, double* x) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (q) CUDA_TRY(cudaMemcpyAsync(q, h->q, sizeof(double) * h->P * h->N, cudaMemcpyDeviceToHost, h->stream));
    if (l && h->L) {
        unpad4to3_kernel<<<cdiv((int64_t)h->L * 3, 256), 256, 0, h->stream>>>(h->L, h->lm4, h->stage3);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(l, h->stage3, sizeof(double) * 3 * h->L, cudaMemcpyDeviceToHost, h->stream));
    }
    if (x) CUDA_TRY(cudaMemcpyAsync(x, h->X, sizeof(double) * 12, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ASC_OK;
}

int asc_error(asc_handle* h, double* error) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!error) return fail(ASC_ERR_INVALID, "null argument");
    h->linearized = false;   // the FK table is overwritten
    return dispatch_error(h, h->q, h->lm4, h->X, error);
}

int asc_linearize(asc_handle* h, asc_linear_stats* stats) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    float a = 0, b = 0;
    rc = dispatch_linearize(h, false, &a, &b);
    if (rc) return rc;
    if (stats) {
        unsigned long long ch = 0;
        CUDA_TRY(cudaMemcpyAsync(&ch, h->cheir, sizeof ch, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->h_status, h->err_lin, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        stats->error = h->h_status[0];
        stats->linearize_ms = a; stats->assemble_ms = b; stats->cheirality = (int64_t)ch;
    }
    return ASC_OK;
}

int asc_optimize(asc_handle* h, int32_t mode, asc_iter_log* log, int32_t log_capacity, int32_t* n_iterations, double* final_error) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (mode != ASC_BATCH_LM && mode != ASC_BATCH_DOGLEG) return fail(ASC_ERR_INVALID, "unknown optimisation mode %d", mode);
    const asc_params& p = h->prm;
    double error = 0.0, lambda = p.lambda_initial, Delta = p.delta_initial;
    rc = dispatch_error(h, h->q, h->lm4, h->X, &error);   // NonlinearOptimizer state at construction
    if (rc) return rc;
    int iterations = 0;
    if (!(error <= p.error_tol) && iterations < p.max_iterations) {   // defaultOptimize, SURVEY Appendix C.1
        double e_old;
        do {
            e_old = error;
            asc_iter_log lg;
            std::memset(&lg, 0, sizeof lg);
            h->outer_iteration = iterations;
            rc = mode == ASC_BATCH_LM ? lm_iterate(h, &error, &lambda, &lg) : dogleg_iterate(h, &error, &Delta, &lg);
            if (rc) break;
            ++iterations;
            lg.iteration = iterations; lg.error = error; lg.lambda_or_delta = mode == ASC_BATCH_LM ? lambda : Delta;
            if (log && iterations <= log_capacity) log[iterations - 1] = lg;
        } while (iterations < p.max_iterations && !converged(p, e_old, error));
    }
    if (n_iterations) *n_iterations = iterations;
    if (final_error) *final_error = error;
    return rc;
}

int asc_indeterminate_key(asc_handle* h, char* kind, int64_t* index) {
    if (!h || !kind || !index) return fail(ASC_ERR_INVALID, "null argument");
    *kind = h->fail_kind; *index = h->fail_index;
    return ASC_OK;
}

int asc_get_projection_blocks(asc_handle* h, double* b, double* jq, double* jl) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    const size_t M = (size_t)h->M, N = h->N;
    double2* bdbg = nullptr;
    double *d_b = nullptr, *d_jq = nullptr, *d_jl = nullptr;
    auto release = [&]() { cudaFree(bdbg); cudaFree(d_b); cudaFree(d_jq); cudaFree(d_jl); };
    auto checked = [&](cudaError_t e, const char* what) -> int {
        if (e == cudaSuccess) return ASC_OK;
        release();
        return fail(ASC_ERR_CUDA, "asc_get_projection_blocks: %s: %s", what, cudaGetErrorString(e));
    };
    if ((rc = checked(cudaMalloc((void**)&bdbg, std::max<size_t>(M, 1) * sizeof(double2)), "cudaMalloc"))) return rc;
    if ((rc = checked(cudaMalloc((void**)&d_b, std::max<size_t>(M, 1) * 2 * sizeof(double)), "cudaMalloc"))) return rc;
    if ((rc = checked(cudaMalloc((void**)&d_jq, std::max<size_t>(M, 1) * 2 * N * sizeof(double)), "cudaMalloc"))) return rc;
    if ((rc = checked(cudaMalloc((void**)&d_jl, std::max<size_t>(M, 1) * 6 * sizeof(double)), "cudaMalloc"))) return rc;
    double2* saved = h->b_dbg;
    h->b_dbg = bdbg;
    rc = dispatch_linearize(h, true, nullptr, nullptr);
    if (!rc && M) rc = dispatch_unpack(h, d_b, d_jq, d_jl);
    h->b_dbg = saved;
    if (rc) {
        cudaStreamSynchronize(h->stream);
        release();
        return rc;
    }
    if (b && M && (rc = checked(cudaMemcpyAsync(b, d_b, M * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "copy of b"))) return rc;
    if (jq && M && (rc = checked(cudaMemcpyAsync(jq, d_jq, M * 2 * N * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "copy of J_q"))) return rc;
    if (jl && M && (rc = checked(cudaMemcpyAsync(jl, d_jl, M * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "copy of J_l"))) return rc;
    if ((rc = checked(cudaStreamSynchronize(h->stream), "synchronize"))) return rc;
    release();
    return ASC_OK;
}

int asc_get_landmark_blocks(asc_handle* h, double* c, double* gl, double* wk) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!h->linearized) return fail(ASC_ERR_INVALID, "call asc_linearize first");
    const size_t L = h->L;
    std::vector<double> c6(L * 6);
    CUDA_TRY(cudaMemcpyAsync(c6.data(), h->C, L * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (gl) CUDA_TRY(cudaMemcpyAsync(gl, h->gl, L * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (wk) CUDA_TRY(cudaMemcpyAsync(wk, h->WK, L * 18 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (c)
        for (size_t j = 0; j < L; ++j) {
            const double* s = &c6[6 * j];
            double* o = c + 9 * j;
            o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[1]; o[4] = s[3]; o[5] = s[4]; o[6] = s[2]; o[7] = s[4]; o[8] = s[5];
        }
    return ASC_OK;
}

int asc_get_pose_blocks(asc_handle* h, double* bpp, double* bpk, double* gp, double* bkk, double* gk) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!h->linearized) return fail(ASC_ERR_INVALID, "call asc_linearize first");
    const size_t PN = (size_t)h->P * h->N;
    if (bpp) CUDA_TRY(cudaMemcpyAsync(bpp, h->Bpp, PN * h->N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (bpk) CUDA_TRY(cudaMemcpyAsync(bpk, h->BpK, PN * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (gp) CUDA_TRY(cudaMemcpyAsync(gp, h->gr, PN * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (bkk) CUDA_TRY(cudaMemcpyAsync(bkk, h->BKK, 36 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (gk) CUDA_TRY(cudaMemcpyAsync(gk, h->gr + PN, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ASC_OK;
}

int asc_solve_damped(asc_handle* h, double lambda, double* dq, double* dl, double* dx, int32_t* pcg_iterations) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!h->linearized) { rc = dispatch_linearize(h, false, nullptr, nullptr); if (rc) return rc; }
    int its = 0;
    rc = dispatch_solve(h, lambda, &its);
    if (pcg_iterations) *pcg_iterations = its;
    if (rc) return rc;
    const size_t PN = (size_t)h->P * h->N;
    if (dq) CUDA_TRY(cudaMemcpyAsync(dq, h->vx, PN * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (dx) CUDA_TRY(cudaMemcpyAsync(dx, h->vx + PN, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (dl && h->L) {
        unpad4to3_kernel<<<cdiv((int64_t)h->L * 3, 256), 256, 0, h->stream>>>(h->L, h->dl4, h->stage3);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(dl, h->stage3, sizeof(double) * 3 * h->L, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ASC_OK;
}

int asc_extrinsic_marginal(asc_handle* h, double* cov36, int32_t* pcg_iterations) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!cov36) return fail(ASC_ERR_INVALID, "null argument");
    if (!h->linearized) { rc = dispatch_linearize(h, false, nullptr, nullptr); if (rc) return rc; }
    int its = 0;
    rc = dispatch_marginal(h, cov36, &its);
    if (pcg_iterations) *pcg_iterations = its;
    return rc;
}

// ---- measurement support
int asc_save_values(asc_handle* h) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->q_save, h->q, sizeof(double) * h->P * h->N, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->lm4_save, h->lm4, sizeof(double) * h->L * 4, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->X_save, h->X, sizeof(double) * 12, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return ASC_OK;
}

int asc_restore_values(asc_handle* h) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->q, h->q_save, sizeof(double) * h->P * h->N, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->lm4, h->lm4_save, sizeof(double) * h->L * 4, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->X, h->X_save, sizeof(double) * 12, cudaMemcpyDeviceToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->linearized = false;
    return ASC_OK;
}

int64_t asc_launch_count(asc_handle* h) { return h ? h->launches : 0; }

int asc_solver_info(asc_handle* h, int32_t* n_clusters, int32_t* coarse_dim, int64_t* cluster_block_doubles) {
    if (!h || !h->problem_set) return fail(ASC_ERR_INVALID, "set the problem first");
    if (n_clusters) *n_clusters = h->nC;
    if (coarse_dim) *coarse_dim = h->use_coarse ? h->nz : 0;
    if (cluster_block_doubles) *cluster_block_doubles = (int64_t)h->scTotal;
    return ASC_OK;
}

int asc_debug_ml_trace(asc_handle* h, unsigned long long* out, int32_t capacity) {
    if (!h || !h->ml_trace) return fail(ASC_ERR_INVALID, "no trace (set ASC_ML_TRACE=1)");
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(out, h->ml_trace, sizeof(unsigned long long) * std::min<size_t>((size_t)capacity, 64 * 148 * 2), cudaMemcpyDeviceToHost));
    return ASC_OK;
}

int asc_coarse_info(asc_handle* h, int32_t* n_levels, int32_t* level_nodes, int64_t* level_blocks, int32_t capacity, int32_t* degree) {
    if (!h || !h->problem_set) return fail(ASC_ERR_INVALID, "set the problem first");
    const int nl = h->use_coarse ? (int)h->ml.size() : 0;
    if (n_levels) *n_levels = nl;
    for (int l = 0; l < nl && l < capacity; ++l) {
        if (level_nodes) level_nodes[l] = h->ml[l].n;
        if (level_blocks) level_blocks[l] = h->ml[l].nnzb;
    }
    if (degree) *degree = h->ml_degree;
    return ASC_OK;
}

// CUDA-event stopwatch on the stream every kernel of this handle is launched on
int asc_timer_start(asc_handle* h) {
    if (!h) return fail(ASC_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->dev));
    if (!h->tm[0]) { CUDA_TRY(cudaEventCreate(&h->tm[0])); CUDA_TRY(cudaEventCreate(&h->tm[1])); }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaEventRecord(h->tm[0], h->stream));
    return ASC_OK;
}

int asc_timer_stop(asc_handle* h, double* elapsed_ms) {
    if (!h || !h->tm[0] || !elapsed_ms) return fail(ASC_ERR_INVALID, "timer not started");
    CUDA_TRY(cudaSetDevice(h->dev));
    CUDA_TRY(cudaEventRecord(h->tm[1], h->stream));
    CUDA_TRY(cudaEventSynchronize(h->tm[1]));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->tm[0], h->tm[1]));
    *elapsed_ms = ms;
    return ASC_OK;
}

int asc_profile_begin(asc_handle* h, int32_t kernel_id, int32_t max_samples) {
    if (!h || kernel_id < 0 || kernel_id > 9 || max_samples < 0) return fail(ASC_ERR_INVALID, "bad profile request");
    CUDA_TRY(cudaSetDevice(h->dev));
    while ((int)h->prof_ev.size() < 2 * max_samples) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        h->prof_ev.push_back(e);
    }
    h->prof_id = kernel_id;
    h->prof_n = 0;
    return ASC_OK;
}

int asc_profile_read(asc_handle* h, int32_t* n_samples, double* mean_ms, double* min_ms) {
    if (!h) return fail(ASC_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(h->dev));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    double sum = 0.0, mn = 1e300;
    for (int i = 0; i < h->prof_n; ++i) {
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
        sum += ms; mn = std::min(mn, (double)ms);
    }
    if (n_samples) *n_samples = h->prof_n;
    if (mean_ms) *mean_ms = h->prof_n ? sum / h->prof_n : 0.0;
    if (min_ms) *min_ms = h->prof_n ? mn : 0.0;
    h->prof_id = 0;
    return ASC_OK;
}

// ---- multi-GPU
int asc_comm_id_size(void) { return (int)sizeof(ncclUniqueId); }

int asc_comm_create_id(void* id_bytes) {
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return fail(ASC_ERR_COMM, "ncclGetUniqueId failed");
    std::memcpy(id_bytes, &id, sizeof id);
    return ASC_OK;
}

int asc_comm_init(asc_handle* h, int32_t rank, int32_t world, const void* id_bytes) {
    if (!h || !id_bytes || rank < 0 || rank >= world) return fail(ASC_ERR_INVALID, "bad communicator arguments");
    if (world == 1) { h->rank = 0; h->world = 1; return ASC_OK; }
    int rc = load_nccl();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(h->dev));
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, id, rank);
    if (r != ncclSuccess) return fail(ASC_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    h->rank = rank; h->world = world;
    return ASC_OK;
}

int asc_shard_by_landmark(int32_t L, int64_t M, const int32_t* lm_idx, int32_t world, int32_t* lm_begin) {
    if (L < 0 || M < 0 || world <= 0 || !lm_begin || (M > 0 && !lm_idx)) return fail(ASC_ERR_INVALID, "bad arguments");
    std::vector<int64_t> cnt(L + 1, 0);
    for (int64_t m = 0; m < M; ++m) {
        if (lm_idx[m] < 0 || lm_idx[m] >= L) return fail(ASC_ERR_INVALID, "landmark index out of range");
        cnt[lm_idx[m] + 1]++;
    }
    for (int j = 0; j < L; ++j) cnt[j + 1] += cnt[j] + 1;   // +1: weigh the landmark itself so empty ones spread too
    const int64_t total = cnt[L];
    lm_begin[0] = 0;
    int j = 0;
    for (int r = 1; r < world; ++r) {
        const int64_t target = total * r / world;
        while (j < L && cnt[j] < target) ++j;
        lm_begin[r] = j;
    }
    lm_begin[world] = L;
    return ASC_OK;
}

}  // extern "C"
