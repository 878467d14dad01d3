// Held-out scoring -- the step right after the hot path in the reference CLI:
//     measure_cv_performance(V, X_test)              prmf/__init__.py:768-798, called at script/prmf_runner.py:1074-1079
// For every held-out sample x:  u* = argmin_{u >= 0} ||x - V u||_2 ,  error = ||x - V u*|| / ||x||.
// The reference loops over the samples in Python calling scipy.optimize.nnls (its own comment: "TODO multi-sample
// version of nnls?").  Here the whole batch is four launches, all fp64, every reduction in a fixed order:
//   cv_gram_kernel   G = V^T V                       (k x k, per-block partials folded in order)
//   cv_xv_kernel     b_i = V^T x_i , ||x_i||^2       (warp per sample, one streaming pass over X_test)
//   cv_nnls_kernel   active-set NNLS on (G, b_i)     (Lawson & Hanson 1995, the method scipy.optimize.nnls implements,
//                                                     stated on the normal equations: entering variable = largest
//                                                     dual, step back to the boundary when a passive coefficient
//                                                     turns negative, tol = 10 max(n, k) eps; thread per sample;
//                                                     V has full column rank, so the minimiser is unique and equals
//                                                     scipy's to rounding -- tests: u rel 1e-8, error rel 1e-9)
//   cv_resid_kernel  ||x_i - V u_i||                 (explicit second pass, like scipy's final np.linalg.norm)
#include "../../include/prmf_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <cfloat>
#include <cstdio>
#include <string>

namespace {

thread_local std::string g_cv_error;

int cv_fail(const char* what, cudaError_t e) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    g_cv_error = buf;
    return PRMF_ERR_CUDA;
}

constexpr int kCvMaxK = 128;      // the engine's largest k (V^T V must fit in shared memory: 128 KB)
constexpr int kGramBlocks = 64;

__device__ __forceinline__ double cv_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// partial Grams over row slices: part[b][a][c] = sum_{j in slice b} V[j][a] V[j][c]
__global__ void __launch_bounds__(256) cv_gram_kernel(const double* __restrict__ V, int64_t n, int k, double* __restrict__ part) {
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t j0 = blockIdx.x * per, j1 = min(n, j0 + per);
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
        const int a = e / k, c = e - a * k;
        double s = 0.0;
        for (int64_t j = j0; j < j1; ++j) s = fma(V[j * k + a], V[j * k + c], s);
        part[(int64_t)blockIdx.x * k * k + e] = s;
    }
}

__global__ void __launch_bounds__(256) cv_gram_sum_kernel(const double* __restrict__ part, int blocks, int kk2, double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= kk2) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += part[(int64_t)b * kk2 + e];
    G[e] = s;
}

// b[i][0:k] = V^T x_i and xx[i] = ||x_i||^2 ; one warp per sample, factor tiles of 16 accumulators
__global__ void __launch_bounds__(256) cv_xv_kernel(const double* __restrict__ X, int64_t ld, int64_t mt, int64_t n,
                                                    const double* __restrict__ V, int k, double* __restrict__ B,
                                                    double* __restrict__ xx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= mt) return;
    const double* xr = X + row * ld;
    for (int c0 = 0; c0 < k; c0 += 16) {
        double acc[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = 0.0;
        double nn = 0.0;
        for (int64_t j = lane; j < n; j += 32) {
            const double x = xr[j];
            const double* vr = V + j * k + c0;
            nn = fma(x, x, nn);
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (c0 + q < k) acc[q] = fma(x, vr[q], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const double s = cv_warp_sum(acc[q]);
            if (lane == 0 && c0 + q < k) B[row * k + c0 + q] = s;
        }
        if (c0 == 0) {
            nn = cv_warp_sum(nn);
            if (lane == 0) xx[row] = nn;
        }
    }
}

// Solve G[P,P] s_P = b_P by Cholesky (G symmetric positive definite on the passive set); idx lists P.
// L: p x p lower factor (row pitch k) in this sample's global workspace.  Returns false when a pivot is not positive.
__device__ bool cv_solve_passive(const double* __restrict__ sG, int k, const double* __restrict__ b,
                                 const int* __restrict__ idx, int p, double* __restrict__ L, double* __restrict__ y) {
    for (int i = 0; i < p; ++i) {
        for (int j = 0; j <= i; ++j) {
            double s = sG[idx[i] * k + idx[j]];
            for (int t = 0; t < j; ++t) s -= L[i * k + t] * L[j * k + t];
            if (i == j) {
                // not positive, or lost to cancellation: column i is (numerically) dependent on the ones before it
                if (!(s > 1e-13 * sG[idx[i] * k + idx[i]])) return false;
                L[i * k + i] = sqrt(s);
            } else {
                L[i * k + j] = s / L[j * k + j];
            }
        }
    }
    for (int i = 0; i < p; ++i) {                       // forward: L y = b_P
        double s = b[idx[i]];
        for (int t = 0; t < i; ++t) s -= L[i * k + t] * y[t];
        y[i] = s / L[i * k + i];
    }
    for (int i = p - 1; i >= 0; --i) {                  // backward: L^T s = y
        double s = y[i];
        for (int t = i + 1; t < p; ++t) s -= L[t * k + i] * y[t];
        y[i] = s / L[i * k + i];
    }
    return true;
}

// One thread per sample.  All per-sample state lives in a global workspace (k*k + 5k doubles and 2k ints per
// sample) rather than in thread-local arrays: local-memory frames are reserved for every thread the device can
// hold, which would cost gigabytes at k = 128.  status[i]: 1 converged, -1 iteration limit, -2 singular block.
__global__ void __launch_bounds__(64) cv_nnls_kernel(const double* __restrict__ G, const double* __restrict__ B, int64_t mt,
                                                     int k, double tol, int maxiter, double* __restrict__ U,
                                                     int* __restrict__ status, double* __restrict__ work_d,
                                                     int* __restrict__ work_i) {
    extern __shared__ double sG[];
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) sG[e] = G[e];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mt) return;
    double* wd = work_d + i * ((int64_t)k * k + 5 * k);
    double* L = wd;
    double* b = L + (int64_t)k * k;
    double* x = b + k;
    double* s = x + k;
    double* w = s + k;
    double* y = w + k;
    int* idx = work_i + i * (2 * (int64_t)k);
    int* P = idx + k;
    for (int c = 0; c < k; ++c) { b[c] = B[i * k + c]; x[c] = 0.0; s[c] = 0.0; w[c] = b[c]; P[c] = 0; }
    int iter = 0, st = 1;
    while (true) {
        // entering variable: largest dual outside the passive set (first maximum wins)
        int enter = -1, np_ = 0;
        bool any = false;
        double best = 0.0;
        // P[c]: 0 at its bound, 1 passive, 2 rejected in this round (its column depends on the passive ones)
        for (int c = 0; c < k; ++c) {
            if (P[c] == 1) { ++np_; continue; }
            if (P[c] == 0 && w[c] > tol) any = true;
        }
        if (np_ == k || !any) break;
        for (int c = 0; c < k; ++c) {
            if (P[c] != 0) continue;
            if (enter < 0 || w[c] > best) { best = w[c]; enter = c; }
        }
        P[enter] = 1;
        int p = 0;
        for (int c = 0; c < k; ++c) { s[c] = 0.0; if (P[c] == 1) idx[p++] = c; }
        if (!cv_solve_passive(sG, k, b, idx, p, L, y)) {
            // Lawson & Hanson's independence test: a candidate whose column is (nearly) a combination of the passive
            // columns cannot lower the residual -- leave it at zero and try the next largest dual
            P[enter] = 2;
            continue;
        }
        for (int c = 0; c < k; ++c)
            if (P[c] == 2) P[c] = 0;
        for (int q = 0; q < p; ++q) s[idx[q]] = y[q];
        while (iter < maxiter) {
            double smin = DBL_MAX;
            for (int q = 0; q < p; ++q) smin = fmin(smin, s[idx[q]]);
            if (!(smin < 0.0)) break;
            ++iter;
            double alpha = DBL_MAX;
            for (int c = 0; c < k; ++c)
                if (P[c] == 1 && s[c] < 0.0) alpha = fmin(alpha, x[c] / (x[c] - s[c]));
            for (int c = 0; c < k; ++c) { x[c] *= (1.0 - alpha); x[c] += alpha * s[c]; }
            for (int c = 0; c < k; ++c)
                if (P[c] == 1 && x[c] <= tol) P[c] = 0;
            p = 0;
            for (int c = 0; c < k; ++c) { if (P[c] == 1) idx[p++] = c; }
            if (p > 0 && !cv_solve_passive(sG, k, b, idx, p, L, y)) { st = -2; break; }
            for (int c = 0; c < k; ++c) s[c] = 0.0;
            for (int q = 0; q < p; ++q) s[idx[q]] = y[q];
        }
        if (st < 0) break;
        for (int c = 0; c < k; ++c) x[c] = s[c];
        for (int c = 0; c < k; ++c) {
            double g = 0.0;
            for (int l = 0; l < k; ++l) g = fma(sG[c * k + l], x[l], g);
            w[c] = b[c] - g;
        }
        if (iter == maxiter) { st = -1; break; }
    }
    for (int c = 0; c < k; ++c) U[i * k + c] = x[c];
    status[i] = st;
}

// rnorm[i] = ||x_i - V u_i||_2 ; one warp per sample
__global__ void __launch_bounds__(256) cv_resid_kernel(const double* __restrict__ X, int64_t ld, int64_t mt, int64_t n,
                                                       const double* __restrict__ V, int k, const double* __restrict__ U,
                                                       double* __restrict__ rnorm) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= mt) return;
    const double* xr = X + row * ld;
    const double* ur = U + row * k;
    double acc = 0.0;
    for (int64_t j = lane; j < n; j += 32) {
        double s = 0.0;
        for (int c = 0; c < k; ++c) s = fma(V[j * k + c], ur[c], s);
        const double d = xr[j] - s;
        acc = fma(d, d, acc);
    }
    acc = cv_warp_sum(acc);
    if (lane == 0) rnorm[row] = sqrt(acc);
}

}  // namespace

extern "C" {

const char* prmf_cv_last_error(void) { return g_cv_error.c_str(); }

int prmf_nnls_rows(int device, const double* V_host, int64_t n, int k, const double* X_host, int64_t mt, int64_t ld,
                   double* U_out, double* rnorm_out, double* xnorm_out, int32_t* status_out) {
    g_cv_error.clear();
    if (!V_host || (!X_host && mt > 0) || n <= 0 || k <= 0 || mt < 0 || ld < n) {
        g_cv_error = "prmf_nnls_rows: bad arguments";
        return PRMF_ERR_ARG;
    }
    if (k > kCvMaxK) {
        g_cv_error = "prmf_nnls_rows: k > 128 is not supported";
        return PRMF_ERR_ARG;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cv_fail("cudaSetDevice", e);
    if (mt == 0) return PRMF_OK;
    double *dV = nullptr, *dX = nullptr, *dG = nullptr, *dGp = nullptr, *dB = nullptr, *dxx = nullptr, *dU = nullptr, *dr = nullptr;
    int* dst = nullptr;
    double* dwd = nullptr;
    int* dwi = nullptr;
    cudaStream_t st = nullptr;
    const int kk2 = k * k;
    int rc = PRMF_OK;
#define CVCU(call) do { e = (call); if (e != cudaSuccess) { rc = cv_fail(#call, e); goto done; } } while (0)
    CVCU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CVCU(cudaMalloc((void**)&dV, sizeof(double) * n * k));
    CVCU(cudaMalloc((void**)&dX, sizeof(double) * mt * n));
    CVCU(cudaMalloc((void**)&dG, sizeof(double) * kk2));
    CVCU(cudaMalloc((void**)&dGp, sizeof(double) * kGramBlocks * kk2));
    CVCU(cudaMalloc((void**)&dB, sizeof(double) * mt * k));
    CVCU(cudaMalloc((void**)&dxx, sizeof(double) * mt));
    CVCU(cudaMalloc((void**)&dU, sizeof(double) * mt * k));
    CVCU(cudaMalloc((void**)&dr, sizeof(double) * mt));
    CVCU(cudaMalloc((void**)&dst, sizeof(int) * mt));
    CVCU(cudaMalloc((void**)&dwd, sizeof(double) * mt * ((size_t)kk2 + 5 * k)));
    CVCU(cudaMalloc((void**)&dwi, sizeof(int) * mt * 2 * k));
    CVCU(cudaMemcpyAsync(dV, V_host, sizeof(double) * n * k, cudaMemcpyHostToDevice, st));
    CVCU(cudaMemcpy2DAsync(dX, n * sizeof(double), X_host, ld * sizeof(double), n * sizeof(double), mt,
                           cudaMemcpyHostToDevice, st));
    {
        cv_gram_kernel<<<kGramBlocks, 256, 0, st>>>(dV, n, k, dGp);
        cv_gram_sum_kernel<<<(kk2 + 255) / 256, 256, 0, st>>>(dGp, kGramBlocks, kk2, dG);
        const unsigned wblocks = (unsigned)((mt * 32 + 255) / 256);
        cv_xv_kernel<<<wblocks, 256, 0, st>>>(dX, n, mt, n, dV, k, dB, dxx);
        const double tol = 10.0 * (double)(n > k ? n : k) * DBL_EPSILON;       // scipy: 10 * max(m, n) * np.spacing(1.)
        if (sizeof(double) * kk2 > 48 * 1024)
            CVCU(cudaFuncSetAttribute(cv_nnls_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * kk2)));
        cv_nnls_kernel<<<(unsigned)((mt + 63) / 64), 64, sizeof(double) * kk2, st>>>(dG, dB, mt, k, tol, 3 * k, dU, dst, dwd,
                                                                                      dwi);
        cv_resid_kernel<<<wblocks, 256, 0, st>>>(dX, n, mt, n, dV, k, dU, dr);
        CVCU(cudaGetLastError());
    }
    if (U_out) CVCU(cudaMemcpyAsync(U_out, dU, sizeof(double) * mt * k, cudaMemcpyDeviceToHost, st));
    if (rnorm_out) CVCU(cudaMemcpyAsync(rnorm_out, dr, sizeof(double) * mt, cudaMemcpyDeviceToHost, st));
    if (xnorm_out) CVCU(cudaMemcpyAsync(xnorm_out, dxx, sizeof(double) * mt, cudaMemcpyDeviceToHost, st));
    if (status_out) CVCU(cudaMemcpyAsync(status_out, dst, sizeof(int) * mt, cudaMemcpyDeviceToHost, st));
    CVCU(cudaStreamSynchronize(st));
#undef CVCU
done:
    cudaFree(dV); cudaFree(dX); cudaFree(dG); cudaFree(dGp); cudaFree(dB); cudaFree(dxx); cudaFree(dU); cudaFree(dr);
    cudaFree(dst); cudaFree(dwd); cudaFree(dwi);
    if (st) cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
