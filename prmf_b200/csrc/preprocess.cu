// GPU quantile normalisation: the step immediately before the PRMF hot path in the reference CLI
// (`X = quantile_transform(X)`, script/prmf_runner.py:1019-1020; sklearn.preprocessing.quantile_transform with its
// defaults: n_quantiles = min(1000, m), uniform output, axis 0).  97 s on the host at 37 032 x 6 750; here:
// transpose -> per-gene segmented sort (CUB) -> 1000 order-statistic interpolations per gene -> two-sided
// interpolation of every entry, all with numpy's exact formulas (no FMA contraction) so results agree with
// sklearn to the last bits.
//
//   quantiles_[q][j] = np.nanpercentile(X[:, j], references*100)[q]      (numpy `_lerp`, method "linear")
//   out[i][j] = 0.5 * (np.interp(x, Q, R) - np.interp(-x, -Q[::-1], -R[::-1]));  x == Q[0] -> 0, x == Q[-1] -> 1
#include "../../include/prmf_b200.h"

#include <cub/device/device_segmented_sort.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <string>

namespace {

thread_local std::string g_pre_error;

int pfail(int code, const char* what, cudaError_t e) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    g_pre_error = buf;
    return code;
}

#define PCU(call)                                                   \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) { rc = pfail(PRMF_ERR_CUDA, #call, e_); goto done; } \
    } while (0)

// Xt[j][r] = X[rows ? rows[r] : r][j]   (gather + transpose; ms rows kept)
__global__ void __launch_bounds__(256)
gather_transpose_kernel(const double* __restrict__ X, int64_t ld, int64_t n, const int64_t* __restrict__ rows, int64_t ms,
                        double* __restrict__ Xt) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = ty; q < 32; q += 8) {
        const int64_t r = r0 + q, j = j0 + tx;
        double v = 0.0;
        if (r < ms && j < n) v = X[(rows ? rows[r] : r) * ld + j];
        tile[q][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = ty; q < 32; q += 8) {
        const int64_t j = j0 + q, r = r0 + tx;
        if (j < n && r < ms) Xt[j * ms + r] = tile[tx][q];
    }
}

__global__ void __launch_bounds__(256)
count_nan_kernel(const double* __restrict__ X, int64_t ld, int64_t m, int64_t n, unsigned long long* __restrict__ count) {
    unsigned long long c = 0;
    for (int64_t row = blockIdx.x; row < m; row += gridDim.x)
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) c += isnan(X[row * ld + j]) ? 1 : 0;
    if (c) atomicAdd(count, c);
}

// Qt[j][q] = lerp(sorted[j][lo[q]], sorted[j][hi[q]], g[q])   -- numpy _lerp, evaluated without contraction
__global__ void __launch_bounds__(256)
quantiles_kernel(const double* __restrict__ sorted, int64_t ms, int64_t n, int nq, const int64_t* __restrict__ lo,
                 const int64_t* __restrict__ hi, const double* __restrict__ g, double* __restrict__ Qt) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * nq) return;
    const int64_t j = idx / nq;
    const int q = (int)(idx - j * nq);
    const double a = sorted[j * ms + lo[q]], b = sorted[j * ms + hi[q]];
    const double t = g[q];
    const double diff = __dsub_rn(b, a);
    double r = __dadd_rn(a, __dmul_rn(diff, t));
    if (t >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
    Qt[idx] = r;
}

// One block per gene: the gene's nq quantiles (and the interpolation slopes, as np.interp precomputes them) sit
// in shared memory; threads stream the gene's m values (transposed layout, coalesced) and write the result.
__global__ void __launch_bounds__(256)
transform_kernel(const double* __restrict__ Xt, int64_t m, int64_t ldt, int64_t n, const double* __restrict__ Qt, int nq,
                 const double* __restrict__ refs, double* __restrict__ Ot) {
    extern __shared__ double sm[];
    double* sQ = sm;                 // nq
    double* sR = sm + nq;            // nq
    double* sSf = sm + 2 * nq;       // nq-1 forward slopes   (R[i+1]-R[i])/(Q[i+1]-Q[i])
    double* sSr = sm + 3 * nq;       // nq-1 reverse slopes   ((-R[i-1])-(-R[i]))/((-Q[i-1])-(-Q[i])), stored at i
    const int64_t j = blockIdx.x;
    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
        sQ[q] = Qt[j * nq + q];
        sR[q] = refs[q];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nq - 1; q += blockDim.x) {
        sSf[q] = __ddiv_rn(__dsub_rn(sR[q + 1], sR[q]), __dsub_rn(sQ[q + 1], sQ[q]));
        sSr[q + 1] = __ddiv_rn(__dsub_rn(-sR[q], -sR[q + 1]), __dsub_rn(-sQ[q], -sQ[q + 1]));
    }
    __syncthreads();
    const double qlo = sQ[0], qhi = sQ[nq - 1];
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
        const double x = Xt[j * ldt + i];
        // forward: np.interp(x, Q, R): jf = last index with Q[jf] <= x
        double fwd;
        if (x > qhi) fwd = sR[nq - 1];
        else if (x < qlo) fwd = sR[0];
        else {
            int lo2 = 0, hi2 = nq;                      // invariant: Q[lo2] <= x, (hi2 == nq or Q[hi2] > x)
            while (hi2 - lo2 > 1) {
                const int mid = (lo2 + hi2) >> 1;
                if (sQ[mid] <= x) lo2 = mid; else hi2 = mid;
            }
            const int jf = lo2;
            if (jf == nq - 1 || sQ[jf] == x) fwd = sR[jf];
            else {
                const double sl = sSf[jf];
                fwd = __dadd_rn(__dmul_rn(sl, __dsub_rn(x, sQ[jf])), sR[jf]);
                if (isnan(fwd)) {
                    fwd = __dadd_rn(__dmul_rn(sl, __dsub_rn(x, sQ[jf + 1])), sR[jf + 1]);
                    if (isnan(fwd) && sR[jf] == sR[jf + 1]) fwd = sR[jf];
                }
            }
        }
        // reverse: np.interp(-x, -Q[::-1], -R[::-1]); with xp'[j'] = -Q[nq-1-j'] the bracket xp'[j'] <= -x is
        // i = first index with Q[i] >= x ... expressed on the reversed arrays exactly as numpy evaluates it
        double rev;
        const double nx = -x;
        if (nx > -qlo) rev = -sR[0];                    // -x > xp'[last] = -Q[0]
        else if (nx < -qhi) rev = -sR[nq - 1];          // -x < xp'[0]   = -Q[nq-1]
        else {
            // j' = last index with xp'[j'] <= -x  <=>  i = nq-1-j' = first index with Q[i] >= x ... (Q[i] >= x)
            int lo2 = -1, hi2 = nq - 1;                 // invariant: (lo2 == -1 or Q[lo2] < x), Q[hi2] >= x
            while (hi2 - lo2 > 1) {
                const int mid = (lo2 + hi2) >> 1;
                if (sQ[mid] >= x) hi2 = mid; else lo2 = mid;
            }
            const int ir = hi2;                         // xp'[j'] = -Q[ir], fp'[j'] = -R[ir]; j' == last <=> ir == 0
            if (ir == 0 || -sQ[ir] == nx) rev = -sR[ir];
            else {
                const double sl = sSr[ir];              // ((-R[ir-1]) - (-R[ir])) / ((-Q[ir-1]) - (-Q[ir]))
                rev = __dadd_rn(__dmul_rn(sl, __dsub_rn(nx, -sQ[ir])), -sR[ir]);
                if (isnan(rev)) {
                    rev = __dadd_rn(__dmul_rn(sl, __dsub_rn(nx, -sQ[ir - 1])), -sR[ir - 1]);
                    if (isnan(rev) && sR[ir] == sR[ir - 1]) rev = -sR[ir];
                }
            }
        }
        double out = __dmul_rn(0.5, __dsub_rn(fwd, rev));
        if (x == qhi) out = 1.0;                        // upper bound first, then lower (sklearn's order)
        if (x == qlo) out = 0.0;
        Ot[j * ldt + i] = out;
    }
}

// out[i][j] = Ot[j][i]
__global__ void __launch_bounds__(256)
transpose_back_kernel(const double* __restrict__ Ot, int64_t ldt, int64_t m, int64_t n, double* __restrict__ out, int64_t ld_out) {
    __shared__ double tile[32][33];
    const int64_t j0 = (int64_t)blockIdx.y * 32, i0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int q = ty; q < 32; q += 8) {
        const int64_t j = j0 + q, i = i0 + tx;
        tile[q][tx] = (j < n && i < m) ? Ot[j * ldt + i] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = ty; q < 32; q += 8) {
        const int64_t i = i0 + q, j = j0 + tx;
        if (i < m && j < n) out[i * ld_out + j] = tile[tx][q];
    }
}

}  // namespace

extern "C" {

const char* prmf_preprocess_last_error(void) { return g_pre_error.c_str(); }

int prmf_quantile_transform(int device, void* stream_, const double* X_dev, int64_t m, int64_t n, int64_t ld,
                            const int64_t* rows_dev, int64_t ms, int nq, const double* refs_host,
                            const int64_t* lo_host, const int64_t* hi_host, const double* g_host, double* out_dev,
                            int64_t ld_out, double* quantiles_out_host) {
    int rc = PRMF_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    double *Xt = nullptr, *sorted = nullptr, *Qt = nullptr, *refs = nullptr, *g = nullptr, *Ot = nullptr;
    int64_t *lo = nullptr, *hi = nullptr, *offs = nullptr;
    void* temp = nullptr;
    unsigned long long* nan_count = nullptr;
    size_t temp_bytes = 0;
    unsigned long long nans = 0;
    if (!X_dev || !out_dev || m <= 0 || n <= 0 || ld < n || ld_out < n || ms <= 0 || ms > m || nq < 1 || nq > ms ||
        !refs_host || !lo_host || !hi_host || !g_host) {
        g_pre_error = "prmf_quantile_transform: bad arguments";
        return PRMF_ERR_ARG;
    }
    if ((size_t)n * (size_t)ms > 2147483647ull) {
        g_pre_error = "prmf_quantile_transform: more than 2^31-1 values to sort (subsample the rows)";
        return PRMF_ERR_ARG;
    }
    {
        cudaError_t e0 = cudaSetDevice(device);
        if (e0 != cudaSuccess) return pfail(PRMF_ERR_CUDA, "cudaSetDevice", e0);
    }
    PCU(cudaMalloc(&nan_count, sizeof(unsigned long long)));
    PCU(cudaMemsetAsync(nan_count, 0, sizeof(unsigned long long), stream));
    count_nan_kernel<<<592, 256, 0, stream>>>(X_dev, ld, m, n, nan_count);
    PCU(cudaMemcpyAsync(&nans, nan_count, sizeof nans, cudaMemcpyDeviceToHost, stream));
    PCU(cudaStreamSynchronize(stream));
    if (nans) {
        g_pre_error = "prmf_quantile_transform: X contains NaN (not supported on the GPU path)";
        rc = PRMF_ERR_ARG;
        goto done;
    }
    PCU(cudaMalloc(&Xt, sizeof(double) * n * ms));
    PCU(cudaMalloc(&sorted, sizeof(double) * n * ms));
    PCU(cudaMalloc(&Qt, sizeof(double) * n * nq));
    PCU(cudaMalloc(&refs, sizeof(double) * nq));
    PCU(cudaMalloc(&g, sizeof(double) * nq));
    PCU(cudaMalloc(&lo, sizeof(int64_t) * nq));
    PCU(cudaMalloc(&hi, sizeof(int64_t) * nq));
    PCU(cudaMalloc(&offs, sizeof(int64_t) * (n + 1)));
    PCU(cudaMemcpyAsync(refs, refs_host, sizeof(double) * nq, cudaMemcpyHostToDevice, stream));
    PCU(cudaMemcpyAsync(g, g_host, sizeof(double) * nq, cudaMemcpyHostToDevice, stream));
    PCU(cudaMemcpyAsync(lo, lo_host, sizeof(int64_t) * nq, cudaMemcpyHostToDevice, stream));
    PCU(cudaMemcpyAsync(hi, hi_host, sizeof(int64_t) * nq, cudaMemcpyHostToDevice, stream));
    {
        std::string offs_host((size_t)(n + 1) * sizeof(int64_t), '\0');
        int64_t* oh = reinterpret_cast<int64_t*>(&offs_host[0]);
        for (int64_t j = 0; j <= n; ++j) oh[j] = j * ms;
        PCU(cudaMemcpyAsync(offs, oh, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, stream));
        PCU(cudaStreamSynchronize(stream));
    }
    {
        dim3 tg((unsigned)((n + 31) / 32), (unsigned)((ms + 31) / 32));
        gather_transpose_kernel<<<tg, 256, 0, stream>>>(X_dev, ld, n, rows_dev, ms, Xt);
    }
    PCU(cub::DeviceSegmentedSort::SortKeys(nullptr, temp_bytes, Xt, sorted, (int)(n * ms), (int)n, offs, offs + 1, stream));
    PCU(cudaMalloc(&temp, temp_bytes ? temp_bytes : 8));
    PCU(cub::DeviceSegmentedSort::SortKeys(temp, temp_bytes, Xt, sorted, (int)(n * ms), (int)n, offs, offs + 1, stream));
    quantiles_kernel<<<(unsigned)((n * nq + 255) / 256), 256, 0, stream>>>(sorted, ms, n, nq, lo, hi, g, Qt);
    PCU(cudaGetLastError());
    if (quantiles_out_host)
        PCU(cudaMemcpyAsync(quantiles_out_host, Qt, sizeof(double) * n * nq, cudaMemcpyDeviceToHost, stream));
    // transform all m rows: reuse `sorted` as the transposed output when no subsample was taken
    cudaFree(Xt); Xt = nullptr;
    cudaFree(sorted); sorted = nullptr;
    PCU(cudaMalloc(&Xt, sizeof(double) * n * m));
    PCU(cudaMalloc(&Ot, sizeof(double) * n * m));
    {
        dim3 tg((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32));
        gather_transpose_kernel<<<tg, 256, 0, stream>>>(X_dev, ld, n, nullptr, m, Xt);
        const size_t smem = sizeof(double) * 4 * (size_t)nq;
        PCU(cudaFuncSetAttribute(transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        transform_kernel<<<(unsigned)n, 256, smem, stream>>>(Xt, m, m, n, Qt, nq, refs, Ot);
        dim3 tb((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
        transpose_back_kernel<<<tb, 256, 0, stream>>>(Ot, m, m, n, out_dev, ld_out);
    }
    PCU(cudaGetLastError());
    PCU(cudaStreamSynchronize(stream));
done:
    cudaFree(Xt); cudaFree(sorted); cudaFree(Qt); cudaFree(refs); cudaFree(g); cudaFree(lo); cudaFree(hi); cudaFree(offs);
    cudaFree(temp); cudaFree(Ot); cudaFree(nan_count);
    return rc;
}

}  // extern "C"
