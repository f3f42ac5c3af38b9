// Dev probe: E^-1 = U^-1 U^-T with U^-1 by block recursion (cuBLAS strsm leaves + sgemm) and a blocked syrk, vs potrs(I).
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
__global__ void ident_ld(float* A, int m, int ld) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < (size_t)m * m) { int c = i / m, r = i % m; A[(size_t)c * ld + r] = (r == c) ? 1.f : 0.f; } }
__global__ void zero_lower(float* A, int m, int ld) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < (size_t)m * m) { int c = i / m, r = i % m; if (r > c) A[(size_t)c * ld + r] = 0.f; } }
__global__ void mirror_upper_colmajor(float* A, int n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < (size_t)n * n) { int c = i / n, r = i % n; if (r > c) A[(size_t)c * n + r] = A[(size_t)r * n + c]; } }

// X (col-major, ld) <- inverse of the upper-triangular U (col-major, ld); blocks of X below the diagonal are left untouched
static void inv_upper(cublasHandle_t b, const float* U, float* X, int m, int ld, float* T, int leaf) {
    const float one = 1.f, zero = 0.f, mone = -1.f;
    if (m <= leaf) {
        ident_ld<<<(int)(((size_t)m * m + 255) / 256), 256>>>(X, m, ld);
        CK(cublasStrsm(b, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, m, &one, U, ld, X, ld));
        return;
    }
    const int m1 = (m / 2 + 31) / 32 * 32, m2 = m - m1;
    const float* U12 = U + (size_t)m1 * ld;
    inv_upper(b, U, X, m1, ld, T, leaf);
    inv_upper(b, U + (size_t)m1 * ld + m1, X + (size_t)m1 * ld + m1, m2, ld, T, leaf);
    // T = U12 * X22 (m1 x m2); X12 = -X11 * T.  X22 is upper triangular: trmm would halve the flops but runs at half the rate
    CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_N, m1, m2, m2, &one, U12, ld, X + (size_t)m1 * ld + m1, ld, &zero, T, m1));
    CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_N, m1, m2, m1, &mone, X, ld, T, m1, &zero, X + (size_t)m1 * ld, ld));
}
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 6006, leaf = argc > 2 ? atoi(argv[2]) : 1600;
    size_t nn = (size_t)n * n;
    float *A, *B, *C, *X, *T, *E0; int* info;
    CK(cudaMalloc(&A, nn * 4)); CK(cudaMalloc(&B, nn * 4)); CK(cudaMalloc(&C, nn * 4)); CK(cudaMalloc(&X, nn * 4)); CK(cudaMalloc(&T, nn * 4)); CK(cudaMalloc(&E0, nn * 4)); CK(cudaMalloc(&info, 4));
    cusolverDnHandle_t s; cublasHandle_t b; CK(cusolverDnCreate(&s)); CK(cublasCreate(&b));
    int lw = 0; CK(cusolverDnSpotrf_bufferSize(s, CUBLAS_FILL_MODE_UPPER, n, A, n, &lw));
    float* work; CK(cudaMalloc(&work, (size_t)lw * 4 + 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto Tm = [&](const char* name, auto f) { for (int rep = 0; rep < 2; ++rep) { cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("%-34s %8.3f ms\n", name, ms); } };
    int blocks = (int)((nn + 255) / 256);
    make_spd<<<blocks, 256>>>(E0, n);
    cudaMemcpy(A, E0, nn * 4, cudaMemcpyDeviceToDevice);
    CK(cusolverDnSpotrf(s, CUBLAS_FILL_MODE_UPPER, n, A, n, work, lw, info));
    const float one = 1.f, zero = 0.f;
    Tm("potrs(I)", [&] { ident_ld<<<blocks, 256>>>(B, n, n); CK(cusolverDnSpotrs(s, CUBLAS_FILL_MODE_UPPER, n, n, A, n, B, n, info)); });
    Tm("recursive U^-1", [&] { inv_upper(b, A, X, n, n, T, leaf); });
    Tm("zero_lower + ssyrk + mirror", [&] { zero_lower<<<blocks, 256>>>(X, n, n); CK(cublasSsyrk(b, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, n, n, &one, X, n, &zero, C, n)); mirror_upper_colmajor<<<blocks, 256>>>(C, n); });
    // check: || E0 * C - I ||_max and vs potrs
    CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, E0, n, C, n, &zero, T, n));
    std::vector<float> h(nn), hb(nn), hc(nn);
    cudaMemcpy(h.data(), T, nn * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), B, nn * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc.data(), C, nn * 4, cudaMemcpyDeviceToHost);
    double worst = 0, dmax = 0, bmax = 0;
    for (size_t i = 0; i < nn; ++i) { double v = h[i] - ((i / n == i % n) ? 1.0 : 0.0); worst = fmax(worst, fabs(v)); dmax = fmax(dmax, fabs((double)hb[i] - hc[i])); bmax = fmax(bmax, fabs(hb[i])); }
    printf("max |E*Einv - I| = %.3e   max |Einv - potrs| = %.3e (scale %.3e)\n", worst, dmax, bmax);
    return 0;
}
