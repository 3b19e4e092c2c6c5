// PTDF construction on the GPU (setup step of ADMM(...), /root/reference/src/helpers/ptdf.jl:1-41):
//   A[l,n] = +1 at line.from, -1 at line.to; B = diag(susceptance); Bl = B*A; Bn = A'*B*A;
//   PTDF = Bl * B_inv with B_inv = inv(Bn) on the non-slack nodes, zero row/column at the slack.
// Bn without the slack is symmetric positive definite for a connected grid, so instead of the reference's dense `inv`
// the reduced system is Cholesky-factorised once and PTDF[:, keep]' = Bn_kk^-1 * Bl[:, keep]' is obtained by two
// triangular solves with the L line columns as right-hand sides (same result to rounding, a third of the flops, no
// explicit inverse).  Assembly and the final transposition are kernels of this file; the factorisation / solves are
// cuSOLVER's dense potrf / potrs (a plain library call for a one-off setup step), loaded at run time so that
// libdopf.so itself does not depend on libcusolver.
#include "../../include/dopf.h"

#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <dlfcn.h>

#include <cstdio>
#include <string>

namespace {

thread_local std::string g_ptdf_error;

struct Solver {
    void *lib = nullptr;
    cusolverStatus_t (*create)(cusolverDnHandle_t *) = nullptr;
    cusolverStatus_t (*destroy)(cusolverDnHandle_t) = nullptr;
    cusolverStatus_t (*set_stream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*potrf_buf)(cusolverDnHandle_t, cublasFillMode_t, int, double *, int, int *) = nullptr;
    cusolverStatus_t (*potrf)(cusolverDnHandle_t, cublasFillMode_t, int, double *, int, double *, int, int *) = nullptr;
    cusolverStatus_t (*potrs)(cusolverDnHandle_t, cublasFillMode_t, int, int, const double *, int, double *, int, int *) = nullptr;
    bool load()
    {
        if (lib) return true;
        const char *names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so"};
        for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
        if (!lib) return false;
#define SYM(field, name) field = (decltype(field))dlsym(lib, name); if (!field) return false
        SYM(create, "cusolverDnCreate"); SYM(destroy, "cusolverDnDestroy"); SYM(set_stream, "cusolverDnSetStream");
        SYM(potrf_buf, "cusolverDnDpotrf_bufferSize"); SYM(potrf, "cusolverDnDpotrf"); SYM(potrs, "cusolverDnDpotrs");
#undef SYM
        return true;
    }
};
Solver g_solver;

// reduced index of node n (the slack is removed); -1 for the slack itself
__device__ __forceinline__ int ridx(int n, int slack) { return n == slack ? -1 : (n > slack ? n - 1 : n); }

// Bn_kk (column-major, M x M, M = N-1) += b * a a'  and  R[:, l] = b * a  (a = incidence row of line l without the slack)
__global__ void k_ptdf_assemble(int L, int M, int slack, const int *from, const int *to, const double *b, double *Bn, double *R)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int i = ridx(from[l], slack), j = ridx(to[l], slack);
    const double s = b[l];
    if (i >= 0) { atomicAdd(Bn + (size_t)i * M + i, s); R[(size_t)l * M + i] += s; }
    if (j >= 0) { atomicAdd(Bn + (size_t)j * M + j, s); R[(size_t)l * M + j] -= s; }
    if (i >= 0 && j >= 0) { atomicAdd(Bn + (size_t)i * M + j, -s); atomicAdd(Bn + (size_t)j * M + i, -s); }
}

// out[l][n] (row-major L x N) = X[ridx(n)][l] (X column-major M x L), zero at the slack; 32x32 tiles through shared memory
__global__ void k_ptdf_transpose(int L, int N, int M, int slack, const double *X, double *out)
{
    __shared__ double tile[32][33];
    const int l0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    {   // read: rows of X (node index) contiguous
        const int n = n0 + threadIdx.x, l = l0 + threadIdx.y;
        for (int k = 0; k < 32; k += 8) {
            const int ll = l + k;
            double x = 0.0;
            if (n < N && ll < L) { const int r = ridx(n, slack); if (r >= 0) x = X[(size_t)ll * M + r]; }
            tile[threadIdx.y + k][threadIdx.x] = x;      // tile[l][n]
        }
    }
    __syncthreads();
    {
        const int n = n0 + threadIdx.x;
        for (int k = 0; k < 32; k += 8) {
            const int ll = l0 + threadIdx.y + k;
            if (n < N && ll < L) out[(size_t)ll * N + n] = tile[threadIdx.y + k][threadIdx.x];
        }
    }
}

}  // namespace

extern "C" const char *dopf_ptdf_last_error(void) { return g_ptdf_error.c_str(); }

extern "C" int dopf_calculate_ptdf(int32_t N, int32_t L, const int32_t *line_from, const int32_t *line_to, const double *susceptance,
                                   int32_t slack, int32_t device, double *out)
{
#define FAIL(code, ...) do { char buf_[400]; snprintf(buf_, sizeof buf_, __VA_ARGS__); g_ptdf_error = buf_; cleanup(); return code; } while (0)
#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) FAIL(DOPF_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)
    int *d_from = nullptr, *d_to = nullptr, *d_info = nullptr;
    double *d_b = nullptr, *d_Bn = nullptr, *d_R = nullptr, *d_out = nullptr, *d_work = nullptr;
    cusolverDnHandle_t hs = nullptr;
    cudaStream_t st = nullptr;
    auto cleanup = [&]() {
        if (hs) g_solver.destroy(hs);
        cudaFree(d_from); cudaFree(d_to); cudaFree(d_info); cudaFree(d_b); cudaFree(d_Bn); cudaFree(d_R); cudaFree(d_out); cudaFree(d_work);
        if (st) cudaStreamDestroy(st);
    };
    if (N < 2 || L < 1 || !line_from || !line_to || !susceptance || !out || slack < 0 || slack >= N) FAIL(DOPF_E_ARG, "dopf_calculate_ptdf: invalid argument");
    for (int l = 0; l < L; ++l)
        if (line_from[l] < 0 || line_from[l] >= N || line_to[l] < 0 || line_to[l] >= N || line_from[l] == line_to[l]) FAIL(DOPF_E_ARG, "dopf_calculate_ptdf: line %d has invalid end nodes", l);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(DOPF_E_CUDA, "no CUDA device available; libdopf has no CPU path");
    if (device >= 0) CKP(cudaSetDevice(device));
    if (!g_solver.load()) FAIL(DOPF_E_UNSUPPORTED, "libcusolver could not be loaded (%s)", dlerror() ? dlerror() : "symbol missing");
    const int M = N - 1;
    CKP(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CKP(cudaMalloc(&d_from, sizeof(int) * L)); CKP(cudaMalloc(&d_to, sizeof(int) * L)); CKP(cudaMalloc(&d_b, sizeof(double) * L));
    CKP(cudaMalloc(&d_info, sizeof(int)));
    CKP(cudaMalloc(&d_Bn, sizeof(double) * (size_t)M * M)); CKP(cudaMalloc(&d_R, sizeof(double) * (size_t)M * L)); CKP(cudaMalloc(&d_out, sizeof(double) * (size_t)L * N));
    CKP(cudaMemcpyAsync(d_from, line_from, sizeof(int) * L, cudaMemcpyHostToDevice, st));
    CKP(cudaMemcpyAsync(d_to, line_to, sizeof(int) * L, cudaMemcpyHostToDevice, st));
    CKP(cudaMemcpyAsync(d_b, susceptance, sizeof(double) * L, cudaMemcpyHostToDevice, st));
    CKP(cudaMemsetAsync(d_Bn, 0, sizeof(double) * (size_t)M * M, st)); CKP(cudaMemsetAsync(d_R, 0, sizeof(double) * (size_t)M * L, st));
    k_ptdf_assemble<<<(L + 127) / 128, 128, 0, st>>>(L, M, slack, d_from, d_to, d_b, d_Bn, d_R);
    CKP(cudaGetLastError());
    if (g_solver.create(&hs) != CUSOLVER_STATUS_SUCCESS) { hs = nullptr; FAIL(DOPF_E_CUDA, "cusolverDnCreate failed"); }
    g_solver.set_stream(hs, st);
    int lwork = 0;
    if (g_solver.potrf_buf(hs, CUBLAS_FILL_MODE_LOWER, M, d_Bn, M, &lwork) != CUSOLVER_STATUS_SUCCESS) FAIL(DOPF_E_CUDA, "cusolverDnDpotrf_bufferSize failed");
    CKP(cudaMalloc(&d_work, sizeof(double) * (size_t)(lwork > 0 ? lwork : 1)));
    int info = 0;
    if (g_solver.potrf(hs, CUBLAS_FILL_MODE_LOWER, M, d_Bn, M, d_work, lwork, d_info) != CUSOLVER_STATUS_SUCCESS) FAIL(DOPF_E_CUDA, "cusolverDnDpotrf failed");
    CKP(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st)); CKP(cudaStreamSynchronize(st));
    if (info != 0) FAIL(DOPF_E_ARG, "the reduced susceptance matrix is not positive definite (leading minor %d): the grid is not connected", info);
    if (g_solver.potrs(hs, CUBLAS_FILL_MODE_LOWER, M, L, d_Bn, M, d_R, M, d_info) != CUSOLVER_STATUS_SUCCESS) FAIL(DOPF_E_CUDA, "cusolverDnDpotrs failed");
    k_ptdf_transpose<<<dim3((N + 31) / 32, (L + 31) / 32), dim3(32, 8), 0, st>>>(L, N, M, slack, d_R, d_out);
    CKP(cudaGetLastError());
    CKP(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)L * N, cudaMemcpyDeviceToHost, st));
    CKP(cudaStreamSynchronize(st));
    cleanup();
    return DOPF_OK;
#undef CKP
#undef FAIL
}
