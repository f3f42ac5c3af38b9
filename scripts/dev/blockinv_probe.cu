// Dev probe: E^-1 by 2x2 block (Schur-complement) inversion with Cholesky-based leaves vs potrf + recursive inverse, n = 6006 fp32.
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
__global__ void mirror_upper_block(float* A, int m, int ld) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < (size_t)m * m) { int c = i / m, r = i % m; if (r > c) A[(size_t)c * ld + r] = A[(size_t)r * ld + c]; } }
static cublasHandle_t b; static cusolverDnHandle_t s; static float* work; static int lw;
static void inv_upper(const float* U, float* X, int m, int ld, float* T) {
    const float one = 1.f, zero = 0.f, mone = -1.f;
    if (m <= 1024) { ident_ld<<<(int)(((size_t)m * m + 255) / 256), 256>>>(X, m, ld); CK(cublasStrsm(b, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, m, &one, U, ld, X, ld)); return; }
    const int m1 = (m / 2 + 31) / 32 * 32, m2 = m - m1;
    inv_upper(U, X, m1, ld, T); inv_upper(U + (size_t)m1 * ld + m1, X + (size_t)m1 * ld + m1, m2, ld, T);
    CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_N, m1, m2, m2, &one, U + (size_t)m1 * ld, ld, X + (size_t)m1 * ld + m1, ld, &zero, T, m1));
    CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_N, m1, m2, m1, &mone, X, ld, T, m1, &zero, X + (size_t)m1 * ld, ld));
}
// C (upper, ld) = X X^T for upper-triangular X (ld), blocked 2x2; X's strictly-lower blocks must be zero
static void tri_syrk(const float* X, float* C, int m, int ld) {
    const float one = 1.f, zero = 0.f;
    const int m1 = (m / 2 + 31) / 32 * 32, m2 = m - m1;
    const float *X12 = X + (size_t)m1 * ld, *X22 = X + (size_t)m1 * ld + m1;
    CK(cublasSsyrk(b, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, m1, m1, &one, X, ld, &zero, C, ld));
    CK(cublasSsyrk(b, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, m1, m2, &one, X12, ld, &one, C, ld));
    CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_T, m1, m2, m2, &one, X12, ld, X22, ld, &zero, C + (size_t)m1 * ld, ld));
    CK(cublasSsyrk(b, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, m2, m2, &one, X22, ld, &zero, C + (size_t)m1 * ld + m1, ld));
}
// full symmetric inverse of the SPD block M (upper valid, ld) into C (ld); X, T scratch; M is overwritten by its factor
static void spd_inverse_block(float* M, float* C, int m, int ld, float* X, float* T, int* info) {
    CK(cusolverDnSpotrf(s, CUBLAS_FILL_MODE_UPPER, m, M, ld, work, lw, info));
    // zero the scratch block (lower part must be zero for tri_syrk)
    CK(cudaMemset2DAsync(X, (size_t)ld * 4, 0, (size_t)m * 4, m));
    inv_upper(M, X, m, ld, T);
    tri_syrk(X, C, m, ld);
    mirror_upper_block<<<(int)(((size_t)m * m + 255) / 256), 256>>>(C, m, ld);
}
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 6006;
    size_t nn = (size_t)n * n;
    float *A, *C, *X, *T, *W, *E0, *Cref; int* info;
    CK(cudaMalloc(&A, nn * 4)); CK(cudaMalloc(&C, nn * 4)); CK(cudaMalloc(&Cref, nn * 4)); CK(cudaMalloc(&X, nn * 4)); CK(cudaMalloc(&T, nn * 4)); CK(cudaMalloc(&W, nn * 4)); CK(cudaMalloc(&E0, nn * 4)); CK(cudaMalloc(&info, 16)); cudaMemset(info, 0, 16);
    CK(cusolverDnCreate(&s)); CK(cublasCreate(&b));
    CK(cusolverDnSpotrf_bufferSize(s, CUBLAS_FILL_MODE_UPPER, n, A, n, &lw)); CK(cudaMalloc(&work, (size_t)lw * 4 + 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto Tm = [&](const char* name, auto f) { for (int rep = 0; rep < 3; ++rep) { cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("%-40s %8.3f ms\n", name, ms); } };
    int blocks = (int)((nn + 255) / 256);
    make_spd<<<blocks, 256>>>(E0, n);
    const float one = 1.f, zero = 0.f, mone = -1.f;
    Tm("copy + potrf + recursive inverse (now)", [&] {
        cudaMemcpyAsync(A, E0, nn * 4, cudaMemcpyDeviceToDevice);
        spd_inverse_block(A, Cref, n, n, X, T, info);
    });
    Tm("copy + 2x2 block inversion", [&] {
        cudaMemcpyAsync(A, E0, nn * 4, cudaMemcpyDeviceToDevice);
        const int m1 = (n / 2 + 31) / 32 * 32, m2 = n - m1;
        float *M12 = A + (size_t)m1 * n, *M22 = A + (size_t)m1 * n + m1;
        float *C12 = C + (size_t)m1 * n, *C22 = C + (size_t)m1 * n + m1;
        spd_inverse_block(A, C, m1, n, X, T, info + 1);                                                     // C11 <- A11^-1 (full)
        CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_N, m1, m2, m1, &one, C, n, M12, n, &zero, W, m1));          // W = A11^-1 M12
        CK(cublasSgemm(b, CUBLAS_OP_T, CUBLAS_OP_N, m2, m2, m1, &mone, M12, n, W, m1, &one, M22, n));        // S = M22 - M12^T W
        spd_inverse_block(M22, C22, m2, n, X + (size_t)m1 * n + m1, T, info + 2);                            // C22 <- S^-1 (full)
        CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_N, m1, m2, m2, &mone, W, m1, C22, n, &zero, C12, n));       // C12 = -W S^-1
        CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_T, m1, m1, m2, &mone, C12, n, W, m1, &one, C, n));          // C11 = A11^-1 - C12 W^T
    });
    std::vector<float> ha(nn), hb(nn); int hinfo[4];
    cudaMemcpy(ha.data(), Cref, nn * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), C, nn * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hinfo, info, 16, cudaMemcpyDeviceToHost);
    double d = 0, sc = 0;
    for (int c = 0; c < n; ++c) for (int r = 0; r <= c; ++r) { d = fmax(d, fabs((double)ha[(size_t)c * n + r] - hb[(size_t)c * n + r])); sc = fmax(sc, fabs(ha[(size_t)c * n + r])); }
    printf("max |Einv_block - Einv_ref| (upper) = %.3e (scale %.3e) info %d %d %d\n", d, sc, hinfo[0], hinfo[1], hinfo[2]);
    return 0;
}
