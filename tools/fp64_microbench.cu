// Microbenchmark: FP64 FMA (DFMA) vs FP64 tensor (DMMA m8n8k4) throughput on sm_100a, alone and mixed.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_microbench tools/fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>   // 0 = DFMA only, 1 = DMMA only, 2 = mixed (1 DMMA : 4 DFMA-instr)
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double f[16], c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { f[i] = threadIdx.x + i; c[i] = i; }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < (MODE == 2 ? 4 : 8); ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_thread_iter) {
    double* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(double));
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, 1000, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)148 * 4 * 256 * iters * fma_per_thread_iter;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s (2*FMA)  err=%s\n", name, ms, 2 * fmas / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    // per thread per iter: DFMA mode 16 FMAs; DMMA mode: 8 mma * 256 FMA / 32 lanes = 64; mixed: 16 + 4*8 = 48
    run<0>("DFMA only", 16);
    run<1>("DMMA m8n8k4 only", 64);
    run<2>("mixed 16 DFMA + 4 DMMA", 48);
    run<0>("DFMA only (again)", 16);
    return 0;
}
