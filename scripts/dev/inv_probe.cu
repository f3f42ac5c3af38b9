// Dev probe: timings of ways to get E^-1 (fp32, n = 6006) out of cuSOLVER / cuBLAS on this GPU.
#include <cstdio>
#include <cstdlib>
#include <vector>
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
__global__ void ident(float* A, int n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < (size_t)n * n) A[i] = (i / n == i % n) ? 1.f : 0.f; }
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 6006;
    size_t nn = (size_t)n * n;
    float *A, *B, *C; int* info;
    CK(cudaMalloc(&A, nn * 4)); CK(cudaMalloc(&B, nn * 4)); CK(cudaMalloc(&C, nn * 4)); CK(cudaMalloc(&info, 4));
    cusolverDnHandle_t s; cublasHandle_t b; CK(cusolverDnCreate(&s)); CK(cublasCreate(&b));
    int lw = 0, lw2 = 0; CK(cusolverDnSpotrf_bufferSize(s, CUBLAS_FILL_MODE_UPPER, n, A, n, &lw)); CK(cusolverDnSpotri_bufferSize(s, CUBLAS_FILL_MODE_UPPER, n, A, n, &lw2));
    if (lw2 > lw) lw = lw2;
    float* work; CK(cudaMalloc(&work, (size_t)lw * 4 + 1024));
    size_t wd = 0, wh = 0;
    CK(cusolverDnXtrtri_bufferSize(s, CUBLAS_FILL_MODE_UPPER, CUBLAS_DIAG_NON_UNIT, n, CUDA_R_32F, A, n, &wd, &wh));
    void* dwork; CK(cudaMalloc(&dwork, wd + 1024)); std::vector<char> hwork(wh + 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto T = [&](const char* name, auto f) { for (int rep = 0; rep < 2; ++rep) { cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) printf("%-34s %8.3f ms\n", name, ms); } };
    int blocks = (int)((nn + 255) / 256);
    for (int pass = 0; pass < 1; ++pass) {
        make_spd<<<blocks, 256>>>(A, n);
        T("potrf", [&] { make_spd<<<blocks, 256>>>(A, n); CK(cusolverDnSpotrf(s, CUBLAS_FILL_MODE_UPPER, n, A, n, work, lw, info)); });
        T("potrs(I)", [&] { ident<<<blocks, 256>>>(B, n); CK(cusolverDnSpotrs(s, CUBLAS_FILL_MODE_UPPER, n, n, A, n, B, n, info)); });
        T("potri", [&] { cudaMemcpy(C, A, nn * 4, cudaMemcpyDeviceToDevice); CK(cusolverDnSpotri(s, CUBLAS_FILL_MODE_UPPER, n, C, n, work, lw, info)); });
        T("Xtrtri", [&] { cudaMemcpy(C, A, nn * 4, cudaMemcpyDeviceToDevice); CK(cusolverDnXtrtri(s, CUBLAS_FILL_MODE_UPPER, CUBLAS_DIAG_NON_UNIT, n, CUDA_R_32F, C, n, dwork, wd, hwork.data(), wh, info)); });
        const float one = 1.f, zero = 0.f;
        T("strsm(U, I) one side", [&] { ident<<<blocks, 256>>>(B, n); CK(cublasStrsm(b, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, A, n, B, n)); });
        T("ssyrk (n^3, dense)", [&] { CK(cublasSsyrk(b, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, n, n, &one, C, n, &zero, B, n)); });
        T("sgemm n^3x2", [&] { CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, C, n, C, n, &zero, B, n)); });
        T("strmm (tri x dense)", [&] { CK(cublasStrmm(b, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, n, &one, C, n, A, n, B, n)); });
        CK(cublasSetMathMode(b, CUBLAS_TF32_TENSOR_OP_MATH));
        T("sgemm n^3x2 TF32", [&] { CK(cublasSgemm(b, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, C, n, C, n, &zero, B, n)); });
        CK(cublasSetMathMode(b, CUBLAS_DEFAULT_MATH));
        T("gemmEx 32F_EMULATED_16BFX9", [&] { auto st = cublasGemmEx(b, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, C, CUDA_R_32F, n, C, CUDA_R_32F, n, &zero, B, CUDA_R_32F, n, CUBLAS_COMPUTE_32F_EMULATED_16BFX9, CUBLAS_GEMM_DEFAULT); if (st) printf("  (gemmEx emulated status %d)\n", (int)st); });
    }
    return 0;
}
