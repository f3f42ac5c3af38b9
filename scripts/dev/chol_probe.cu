// Dev probe: recursive Cholesky (UPPER, column-major) with a single-CTA shared-memory leaf vs cusolverDnSpotrf, n = 6006 fp32.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cublas_v2.h>
#define CK(x) do { auto e = (x); if (e != 0) { printf("fail %s = %d line %d\n", #x, (int)e, __LINE__); exit(1);} } while (0)
__global__ void make_spd(float* A, int n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * n) return;
    int r = i / n, c = i % n;
    unsigned h = (unsigned)(min(r, c) * 2654435761u + max(r, c) * 40503u);
    float v = ((h >> 8) & 0xffff) / 65536.0f - 0.5f;
    A[i] = (r == c) ? (float)n * 0.3f : v;
}
constexpr int kLeaf = 96;
// U^T U = A for one m x m block (m <= kLeaf), upper triangle of column-major A (ld); in place.  One CTA, 256 threads.
__global__ void __launch_bounds__(256) chol_leaf_kernel(float* A, int m, int ld, int* info) {
    __shared__ float s[kLeaf][kLeaf + 1];   // s[r][c], r <= c used
    const int t = threadIdx.x;
    for (int i = t; i < m * m; i += 256) { const int c = i / m, r = i % m; s[r][c] = (r <= c) ? A[(size_t)c * ld + r] : 0.f; }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
        const float piv = s[k][k];
        if (!(piv > 0.f)) { if (t == 0) atomicMax(info, k + 1); return; }
        const float d = sqrtf(piv), id = 1.f / d;
        __syncthreads();
        // row k of U: U[k][c] = A[k][c] / d, c > k
        for (int c = k + t; c < m; c += 256) s[k][c] = (c == k) ? d : s[k][c] * id;
        __syncthreads();
        // trailing update A[r][c] -= U[k][r] U[k][c], k < r <= c
        const int w = m - k - 1;
        for (int i = t; i < w * w; i += 256) {
            const int r = k + 1 + i / w, c = k + 1 + i % w;
            if (r <= c) s[r][c] -= s[k][r] * s[k][c];
        }
        __syncthreads();
    }
    for (int i = t; i < m * m; i += 256) { const int c = i / m, r = i % m; if (r <= c) A[(size_t)c * ld + r] = s[r][c]; }
}
static void chol_rec(cublasHandle_t b, float* A, int m, int ld, int* info) {
    if (m <= kLeaf) { chol_leaf_kernel<<<1, 256>>>(A, m, ld, info); return; }
    const float one = 1.f, mone = -1.f;
    const int m1 = (m / 2 + 31) / 32 * 32, m2 = m - m1;
    float *A12 = A + (size_t)m1 * ld, *A22 = A + (size_t)m1 * ld + m1;
    chol_rec(b, A, m1, ld, info);
    CK(cublasStrsm(b, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, m1, m2, &one, A, ld, A12, ld));
    CK(cublasSsyrk(b, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, m2, m1, &mone, A12, ld, &one, A22, ld));
    chol_rec(b, A22, m2, ld, info);
}
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 6006;
    size_t nn = (size_t)n * n;
    float *A, *B, *E0; int* info;
    CK(cudaMalloc(&A, nn * 4)); CK(cudaMalloc(&B, nn * 4)); CK(cudaMalloc(&E0, nn * 4)); CK(cudaMalloc(&info, 8)); cudaMemset(info, 0, 8);
    cusolverDnHandle_t s; cublasHandle_t b; CK(cusolverDnCreate(&s)); CK(cublasCreate(&b));
    int lw = 0; CK(cusolverDnSpotrf_bufferSize(s, CUBLAS_FILL_MODE_UPPER, n, A, n, &lw));
    float* work; CK(cudaMalloc(&work, (size_t)lw * 4 + 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto Tm = [&](const char* name, auto f) { for (int rep = 0; rep < 3; ++rep) { cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("%-34s %8.3f ms\n", name, ms); } };
    int blocks = (int)((nn + 255) / 256);
    make_spd<<<blocks, 256>>>(E0, n);
    Tm("copy only", [&] { cudaMemcpyAsync(A, E0, nn * 4, cudaMemcpyDeviceToDevice); });
    Tm("copy + cusolver potrf", [&] { cudaMemcpyAsync(A, E0, nn * 4, cudaMemcpyDeviceToDevice); CK(cusolverDnSpotrf(s, CUBLAS_FILL_MODE_UPPER, n, A, n, work, lw, info + 1)); });
    Tm("copy + recursive chol", [&] { cudaMemcpyAsync(B, E0, nn * 4, cudaMemcpyDeviceToDevice); chol_rec(b, B, n, n, info); });
    std::vector<float> ha(nn), hb(nn); int hinfo[2];
    cudaMemcpy(ha.data(), A, nn * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), B, nn * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hinfo, info, 8, cudaMemcpyDeviceToHost);
    double d = 0, sc = 0;
    for (int c = 0; c < n; ++c) for (int r = 0; r <= c; ++r) { d = fmax(d, fabs((double)ha[(size_t)c * n + r] - hb[(size_t)c * n + r])); sc = fmax(sc, fabs(ha[(size_t)c * n + r])); }
    printf("max |U_rec - U_cusolver| = %.3e (scale %.3e) info %d %d\n", d, sc, hinfo[0], hinfo[1]);
    return 0;
}
