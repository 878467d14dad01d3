// PRMF hot-path kernels for sm_100a (B200).  All arithmetic is IEEE fp64; every reduction has a fixed
// order so results are bitwise reproducible from run to run.
//
// Reference lines are in /root/reference/script/prmf_runner.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prmf {

constexpr double kEps = 1.1920928955078125e-07;   // np.finfo(np.float32).eps, prmf_runner.py:23
constexpr int kObjStride = 8;

__device__ __forceinline__ double2 ld_stream(const double* p) {
    // streaming 128-bit load, do not allocate in L1 (X is read once per pass)
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reduction (fixed tree); result valid in thread 0.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (warp == 0) {
        double t = lane < nw ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[0] = t;
    }
    __syncthreads();
    return scratch[0];
}

// ----------------------------------------------------------------------------------------------------
// ||X||_F^2  (np.linalg.norm(X), :640) : per-block partials, summed in order by sum_partials_kernel
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const double* __restrict__ X, int64_t ldx, int64_t m,
                                                    int n2, double* __restrict__ part) {
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int64_t row = blockIdx.x; row < m; row += gridDim.x) {
        const double* xr = X + row * ldx;
        double a = 0.0;
        for (int j = threadIdx.x * 2; j < n2; j += blockDim.x * 2) {
            double2 x = ld_stream(xr + j);
            a = fma(x.x, x.x, a);
            a = fma(x.y, x.y, a);
        }
        acc += a;
    }
    double t = block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}

__global__ void sum_partials_kernel(const double* __restrict__ part, int count, double* __restrict__ out) {
    __shared__ double scratch[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) a += part[i];
    double t = block_sum(a, scratch);
    if (threadIdx.x == 0) out[0] = t;
}

// ----------------------------------------------------------------------------------------------------
// Pass 1 over X:  A[:, k0:k0+KT] = X . V[:, k0:k0+KT]          (U_up_num, :420)
// A warp owns RW consecutive rows; lanes stride along the genes (coalesced 128-bit streaming loads of X,
// coalesced loads of the factor-major copy Vt which stays L1/L2 resident); one shuffle reduction per row.
// ----------------------------------------------------------------------------------------------------
template <int KT, int RW>
__global__ void __launch_bounds__(256, 2)
xv_kernel(const double* __restrict__ X, int64_t ldx, int64_t m, int n2, const double* __restrict__ Vt,
          int64_t ldvt, int k0, int k, double* __restrict__ A) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double* vt = Vt + (int64_t)k0 * ldvt;
    for (int64_t row0 = warp * RW; row0 < m; row0 += nwarps * RW) {
        const double* xr[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            int64_t row = row0 + r < m ? row0 + r : m - 1;     // tail rows recompute the last row
            xr[r] = X + row * ldx;
        }
        double acc[RW][KT];
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int c = 0; c < KT; ++c) acc[r][c] = 0.0;
#pragma unroll 1
        for (int j = lane * 2; j < n2; j += 64) {
            double2 x[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r) x[r] = ld_stream(xr[r] + j);
#pragma unroll
            for (int c = 0; c < KT; ++c) {
                const double2 v = *reinterpret_cast<const double2*>(vt + (int64_t)c * ldvt + j);
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    acc[r][c] = fma(x[r].x, v.x, acc[r][c]);
                    acc[r][c] = fma(x[r].y, v.y, acc[r][c]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int c = 0; c < KT; ++c) acc[r][c] = warp_sum(acc[r][c]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < RW; ++r)
                if (row0 + r < m) {
#pragma unroll
                    for (int c = 0; c < KT; ++c) A[(row0 + r) * k + k0 + c] = acc[r][c];
                }
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// U update (:421-422) + per-block partials of U^T U (:425).  sum(U^2) (:359) is its trace.
//   den = U.Gv + U ; U <- U * (A / den  if den != 0 else 1)
// Persistent blocks loop over row tiles; the tile of new U rows is staged in shared memory and every
// thread accumulates its (a,b) pairs of the Gram matrix in registers across tiles.
// ----------------------------------------------------------------------------------------------------
constexpr int kMaxPairsPerThread = 64;   // k <= 128 with 256 threads

template <int NQ>
__global__ void __launch_bounds__(256)
u_update_kernel(double* __restrict__ U, const double* __restrict__ A, const double* __restrict__ Gv,
                int64_t m, int k, int rows_per_tile, double* __restrict__ Gu_part) {
    extern __shared__ double sm[];
    double* sGv = sm;                  // k*k
    double* sU = sm + k * k;           // rows_per_tile * k   (new U rows)
    const int kk2 = k * k;
    for (int i = threadIdx.x; i < kk2; i += blockDim.x) sGv[i] = Gv[i];
    double gacc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) gacc[q] = 0.0;
    __syncthreads();
    const int64_t ntiles = (m + rows_per_tile - 1) / rows_per_tile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * rows_per_tile;
        const int rows = (int)min((int64_t)rows_per_tile, m - r0);
        for (int e = threadIdx.x; e < rows * k; e += blockDim.x) {
            const int r = e / k, c = e - r * k;
            const double* urow = U + (r0 + r) * k;
            double den = 0.0;
            for (int l = 0; l < k; ++l) den = fma(urow[l], sGv[l * k + c], den);
            const double u = urow[c];
            den += u;
            const double a = A[(r0 + r) * k + c];
            const double f = (den != 0.0) ? a / den : 1.0;
            sU[e] = u * f;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < rows * k; e += blockDim.x) U[r0 * k + e] = sU[e];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int pidx = threadIdx.x + q * 256;
            if (pidx < kk2) {
                const int a = pidx / k, b = pidx - a * k;
                double s = gacc[q];
                for (int r = 0; r < rows; ++r) s = fma(sU[r * k + a], sU[r * k + b], s);
                gacc[q] = s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int pidx = threadIdx.x + q * 256;
        if (pidx < kk2) Gu_part[(int64_t)blockIdx.x * kk2 + pidx] = gacc[q];
    }
}

// ----------------------------------------------------------------------------------------------------
// Pass 2 over X:  Bpart[chunk][:, k0:k0+KT] = X[rows of chunk]^T . U[rows of chunk, k0:k0+KT]   (:424)
// A thread owns 4 consecutive genes (two 128-bit streaming loads per row) and KT accumulators for each;
// the U row is the same for the whole block and comes from shared memory (broadcast).
// ----------------------------------------------------------------------------------------------------
constexpr int kXtuRowsPerStage = 32;

template <int KT>
__global__ void __launch_bounds__(256, 2)
xtu_kernel(const double* __restrict__ X, int64_t ldx, int64_t m, int n, const double* __restrict__ U,
           int k, int k0, int panel_w, int64_t rows_per_chunk, double* __restrict__ Bpart) {
    __shared__ double sU[kXtuRowsPerStage][KT];
    const int panel = blockIdx.x;
    const int64_t chunk = blockIdx.y;
    const int jl = threadIdx.x * 4;
    const int j0 = panel * panel_w + jl;
    const bool active = (jl < panel_w) && (j0 < (int)ldx);
    const int64_t rbeg = chunk * rows_per_chunk;
    const int64_t rend = min(m, rbeg + rows_per_chunk);
    double acc[4][KT];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int c = 0; c < KT; ++c) acc[g][c] = 0.0;
    const double* xp = X + j0;
    for (int64_t r0 = rbeg; r0 < rend; r0 += kXtuRowsPerStage) {
        const int rows = (int)min((int64_t)kXtuRowsPerStage, rend - r0);
        __syncthreads();
        for (int e = threadIdx.x; e < rows * KT; e += blockDim.x) {
            const int r = e / KT, c = e - r * KT;
            sU[r][c] = U[(r0 + r) * k + k0 + c];
        }
        __syncthreads();
        if (active) {
            int r = 0;
            for (; r + 4 <= rows; r += 4) {
                double2 xa[4], xb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double* p = xp + (r0 + r + q) * ldx;
                    xa[q] = ld_stream(p);
                    xb[q] = ld_stream(p + 2);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int c = 0; c < KT; ++c) {
                        const double u = sU[r + q][c];
                        acc[0][c] = fma(xa[q].x, u, acc[0][c]);
                        acc[1][c] = fma(xa[q].y, u, acc[1][c]);
                        acc[2][c] = fma(xb[q].x, u, acc[2][c]);
                        acc[3][c] = fma(xb[q].y, u, acc[3][c]);
                    }
            }
            for (; r < rows; ++r) {
                const double* p = xp + (r0 + r) * ldx;
                const double2 xa = ld_stream(p), xb = ld_stream(p + 2);
#pragma unroll
                for (int c = 0; c < KT; ++c) {
                    const double u = sU[r][c];
                    acc[0][c] = fma(xa.x, u, acc[0][c]);
                    acc[1][c] = fma(xa.y, u, acc[1][c]);
                    acc[2][c] = fma(xb.x, u, acc[2][c]);
                    acc[3][c] = fma(xb.y, u, acc[3][c]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int j = j0 + g;
            if (j < n) {
                double* out = Bpart + ((int64_t)chunk * n + j) * k + k0;
#pragma unroll
                for (int c = 0; c < KT; ++c) out[c] = acc[g][c];
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// Sum the per-chunk X^T U partials and per-block U^T U partials in a fixed order into the packed
// buffer that is all-reduced over ranks:  red = [ B (n*k) | Gu (k*k) | sum(U^2) | pad ]
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reduce_pack_kernel(const double* __restrict__ Bpart, int chunks, int64_t nk, const double* __restrict__ Gu_part,
                   int gu_blocks, int k, double* __restrict__ red) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int kk2 = k * k;
    if (idx < nk) {
        double s = 0.0;
        for (int c = 0; c < chunks; ++c) s += Bpart[(int64_t)c * nk + idx];
        red[idx] = s;
    } else if (idx < nk + kk2) {
        const int e = (int)(idx - nk);
        double s = 0.0;
        for (int b = 0; b < gu_blocks; ++b) s += Gu_part[(int64_t)b * kk2 + e];
        red[idx] = s;
    }
}

// sum(U^2) = trace(U^T U); runs after the all-reduce so every rank derives it from identical data.
__global__ void fro_from_gram_kernel(double* __restrict__ red, int64_t nk, int k) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int a = 0; a < k; ++a) s += red[nk + a * k + a];
        red[nk + k * k] = s;
    }
}

// ----------------------------------------------------------------------------------------------------
// Packed pathway tables (device view)
// ----------------------------------------------------------------------------------------------------
struct Pathways {
    int P;
    const int64_t* path_ptr;     // P+1
    const int32_t* support_idx;  // S
    const int64_t* row_ptr;      // S+1
    const int32_t* col_local;    // E
    const double* w;             // E
    const double* deg;           // S   column sums of W                         (:680)
    const double* ldiag;         // S   diag(L) = deg - self-loop weight          (:683)
    const double* isd;           // S   diag(L)^-1/2, 0 where diag(L) == 0        (:58-62)
};

// pos[j*k + c] = packed row of gene j in the active pathway of factor c, or -1
__global__ void build_pos_kernel(Pathways pw, const int32_t* __restrict__ active, int k, int32_t* __restrict__ pos) {
    const int c = blockIdx.y;
    const int p = active[c];
    const int64_t beg = pw.path_ptr[p], end = pw.path_ptr[p + 1];
    for (int64_t r = beg + blockIdx.x * blockDim.x + threadIdx.x; r < end; r += (int64_t)gridDim.x * blockDim.x)
        pos[(int64_t)pw.support_idx[r] * k + c] = (int32_t)r;
}

// ----------------------------------------------------------------------------------------------------
// V update (:425-444) + per-block partials of V_new^T V_new and sum(V_new * B).
//   C = V.Gu ; num = B + (gamma*W v + delta*(v+1)^-2 on the support) ; den = C + gamma*deg*v
//   den < eps -> eps ; V <- V*num/den ; V < eps -> eps
// red = [B | Gu | fro] (after the all-reduce).  gd = {gamma, delta} on the device.
// ----------------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(256)
v_update_kernel(double* __restrict__ V, double* __restrict__ Vt, int64_t ldvt, const double* __restrict__ red,
                int n, int k, Pathways pw, const int32_t* __restrict__ active, const int32_t* __restrict__ pos,
                const double* __restrict__ gd, int rows_per_tile, double* __restrict__ Gv_part,
                double* __restrict__ VB_part) {
    extern __shared__ double sm[];
    double* sGu = sm;                   // k*k
    double* sV = sm + k * k;            // rows_per_tile*k (new V rows)
    __shared__ double scratch[32];
    const int kk2 = k * k;
    const int64_t nk = (int64_t)n * k;
    const double* B = red;
    for (int i = threadIdx.x; i < kk2; i += blockDim.x) sGu[i] = red[nk + i];
    const double gamma = gd[0], delta = gd[1];
    double gacc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) gacc[q] = 0.0;
    double vb = 0.0;
    __syncthreads();
    const int ntiles = (n + rows_per_tile - 1) / rows_per_tile;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int j0 = tile * rows_per_tile;
        const int rows = min(rows_per_tile, n - j0);
        for (int e = threadIdx.x; e < rows * k; e += blockDim.x) {
            const int r = e / k, c = e - r * k;
            const int j = j0 + r;
            const double* vrow = V + (int64_t)j * k;
            double cden = 0.0;
            for (int l = 0; l < k; ++l) cden = fma(vrow[l], sGu[l * k + c], cden);   // V.Gu   (:425)
            const double v = vrow[c];
            const double b = B[(int64_t)j * k + c];
            double num = b, den = cden;
            const int32_t pr = pos[(int64_t)j * k + c];
            if (pr >= 0) {
                const int64_t base = pw.path_ptr[active[c]];
                double wv = 0.0;
                for (int64_t e2 = pw.row_ptr[pr]; e2 < pw.row_ptr[pr + 1]; ++e2)
                    wv = fma(pw.w[e2], V[(int64_t)pw.support_idx[base + pw.col_local[e2]] * k + c], wv);
                const double vp1 = v + 1.0;
                const double man = gamma * wv;                                          // :434
                const double ign = delta * (1.0 / (vp1 * vp1));                         // :438
                num = b + (man + ign);                                                  // :440
                den = cden + gamma * (pw.deg[pr] * v);                                  // :435,:441
            }
            if (den < kEps) den = kEps;                                                 // :442
            double vn = v * (num / den);                                                // :443
            if (vn < kEps) vn = kEps;                                                   // :444
            sV[e] = vn;
            vb = fma(vn, b, vb);
        }
        __syncthreads();
        // Pathway neighbours of a gene may live in another block's tile, so the gene-major V must stay
        // intact until every block is done: new values go to the factor-major copy Vt only, and
        // vt_to_v_kernel refreshes V afterwards.
        for (int e = threadIdx.x; e < rows * k; e += blockDim.x) {
            const int r = e / k, c = e - r * k;
            Vt[(int64_t)c * ldvt + j0 + r] = sV[e];
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int pidx = threadIdx.x + q * 256;
            if (pidx < kk2) {
                const int a = pidx / k, b2 = pidx - a * k;
                double s = gacc[q];
                for (int r = 0; r < rows; ++r) s = fma(sV[r * k + a], sV[r * k + b2], s);
                gacc[q] = s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int pidx = threadIdx.x + q * 256;
        if (pidx < kk2) Gv_part[(int64_t)blockIdx.x * kk2 + pidx] = gacc[q];
    }
    const double t = block_sum(vb, scratch);
    if (threadIdx.x == 0) VB_part[blockIdx.x] = t;
}

// V (gene-major, n x k) <- Vt (factor-major, k x ldvt): second half of the V update.  The update kernel
// reads old V (own row and pathway neighbours in other tiles) and writes only Vt, so no block can see a
// half-updated V.
__global__ void __launch_bounds__(256)
vt_to_v_kernel(const double* __restrict__ Vt, int64_t ldvt, int n, int k, double* __restrict__ V) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (int64_t)n * k) {
        const int j = (int)(idx / k), c = (int)(idx - (int64_t)j * k);
        V[idx] = Vt[(int64_t)c * ldvt + j];
    }
}

__global__ void __launch_bounds__(256)
v_to_vt_kernel(const double* __restrict__ V, int n, int k, double* __restrict__ Vt, int64_t ldvt) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (int64_t)n * k) {
        const int j = (int)(idx / k), c = (int)(idx - (int64_t)j * k);
        Vt[(int64_t)c * ldvt + j] = V[idx];
    }
}

// Gram of V from scratch (after prmf_set_UV): same tiling as the update kernel so partial layout matches.
template <int NQ>
__global__ void __launch_bounds__(256)
gram_rows_kernel(const double* __restrict__ M, int64_t rows_total, int k, int rows_per_tile,
                 double* __restrict__ G_part) {
    extern __shared__ double sm[];
    double* sT = sm;
    const int kk2 = k * k;
    double gacc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) gacc[q] = 0.0;
    const int64_t ntiles = (rows_total + rows_per_tile - 1) / rows_per_tile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * rows_per_tile;
        const int rows = (int)min((int64_t)rows_per_tile, rows_total - r0);
        __syncthreads();
        for (int e = threadIdx.x; e < rows * k; e += blockDim.x) sT[e] = M[r0 * k + e];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int pidx = threadIdx.x + q * 256;
            if (pidx < kk2) {
                const int a = pidx / k, b = pidx - a * k;
                double s = gacc[q];
                for (int r = 0; r < rows; ++r) s = fma(sT[r * k + a], sT[r * k + b], s);
                gacc[q] = s;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int pidx = threadIdx.x + q * 256;
        if (pidx < kk2) G_part[(int64_t)blockIdx.x * kk2 + pidx] = gacc[q];
    }
}

// G[e] = sum_b G_part[b][e]  (fixed order)
__global__ void __launch_bounds__(256)
sum_gram_parts_kernel(const double* __restrict__ G_part, int blocks, int kk2, double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < kk2) {
        double s = 0.0;
        for (int b = 0; b < blocks; ++b) s += G_part[(int64_t)b * kk2 + e];
        G[e] = s;
    }
}

// ----------------------------------------------------------------------------------------------------
// Objective of one inner step (:336-372) without a pass over X, and the tradeoff feedback (:542-548).
//   recon^2 = ||X||^2 - 2 sum(V_new*B) + sum(Gu*Gv_new)        (B = X^T U_new, Gu = U_new^T U_new)
//   manifold = sum_k vhat_k^T Lhat_{p_k} vhat_k ; ignore = sum_k sum_{i in supp} 1/(vhat_i + 1)
// One block.  Also publishes Gv_new for the next step's U update.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
objective_kernel(const double* __restrict__ V, int n, int k, const double* __restrict__ red,
                 const double* __restrict__ Gv_part, const double* __restrict__ VB_part, int vblocks,
                 const double* __restrict__ normX_sq, Pathways pw, const int32_t* __restrict__ active,
                 double* __restrict__ Gv, double* __restrict__ gd, double tradeoff,
                 double* __restrict__ obj_out, int* __restrict__ step_counter, int obj_capacity) {
    extern __shared__ double sm[];
    double* sGv = sm;        // k*k
    __shared__ double scratch[32];
    const int kk2 = k * k;
    const int64_t nk = (int64_t)n * k;
    double gg = 0.0;
    for (int e = threadIdx.x; e < kk2; e += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < vblocks; ++b) s += Gv_part[(int64_t)b * kk2 + e];
        sGv[e] = s;
        Gv[e] = s;
        gg = fma(s, red[nk + e], gg);
    }
    double vb = 0.0;
    for (int b = threadIdx.x; b < vblocks; b += blockDim.x) vb += VB_part[b];
    const double GG = block_sum(gg, scratch);
    const double VB = block_sum(vb, scratch);
    __syncthreads();
    // manifold / ignore: (factor, support row) pairs strided over the block
    double man = 0.0, ign = 0.0;
    for (int c = 0; c < k; ++c) {
        const int p = active[c];
        const int64_t beg = pw.path_ptr[p], end = pw.path_ptr[p + 1];
        const double nrm = sqrt(sGv[c * k + c]);
        for (int64_t r = beg + threadIdx.x; r < end; r += blockDim.x) {
            const double vr = V[(int64_t)pw.support_idx[r] * k + c] / nrm;                 // :345
            const double ir = pw.isd[r];
            double y = (ir * (pw.ldiag[r] * ir)) * vr;                                      // diagonal of Lhat
            for (int64_t e2 = pw.row_ptr[r]; e2 < pw.row_ptr[r + 1]; ++e2) {
                const int cl = pw.col_local[e2];
                if (beg + cl == r) continue;                                                // self loop is in ldiag
                const double vc = V[(int64_t)pw.support_idx[beg + cl] * k + c] / nrm;
                y = fma(ir * (-pw.w[e2] * pw.isd[beg + cl]), vc, y);
            }
            man = fma(y, vr, man);                                                          // :350
            ign += 1.0 / (vr + 1.0);                                                        // :352
        }
    }
    const double MAN = block_sum(man, scratch);
    const double IGN = block_sum(ign, scratch);
    if (threadIdx.x == 0) {
        const double gamma = gd[0], delta = gd[1];
        double r2 = normX_sq[0] - 2.0 * VB + GG;
        const double recon = sqrt(r2 > 0.0 ? r2 : 0.0);
        const double fro = red[nk + kk2];
        const double obj = recon + gamma * MAN + delta * IGN + fro;                         // :362
        const int s = *step_counter;
        if (s < obj_capacity) {
            double* o = obj_out + (int64_t)s * kObjStride;
            o[0] = recon; o[1] = MAN; o[2] = IGN; o[3] = fro; o[4] = obj; o[5] = gamma; o[6] = delta; o[7] = r2;
        }
        *step_counter = s + 1;
        if (tradeoff >= 0.0) {                                                              // :542-548
            const double den = tradeoff * MAN;
            const double g2 = (den == 0.0) ? 1.0 : ((1.0 - tradeoff) * recon) / den;
            gd[0] = g2; gd[1] = g2;
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// Factor x pathway tables (restrict :129-194 / score :115-127, force_distinct_lapls :232, find_mins :49)
// One block per pathway; a warp per factor (strided); lanes over support rows; fixed-order shuffles.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scores_kernel(const double* __restrict__ V, int k, const double* __restrict__ Gv, Pathways pw,
              double* __restrict__ mass, double* __restrict__ quad_norm, double* __restrict__ quad_raw) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int p = blockIdx.x; p < pw.P; p += gridDim.x) {
        const int64_t beg = pw.path_ptr[p], end = pw.path_ptr[p + 1];
        for (int c = warp; c < k; c += nw) {
            const double nrm = sqrt(Gv[c * k + c]);
            double ms = 0.0, qn = 0.0, qr = 0.0;
            for (int64_t r = beg + lane; r < end; r += 32) {
                const double v = V[(int64_t)pw.support_idx[r] * k + c];
                const double vu = v / nrm;
                const double ir = pw.isd[r];
                double yn = (ir * (pw.ldiag[r] * ir)) * vu;
                double yr = pw.ldiag[r] * v;
                for (int64_t e2 = pw.row_ptr[r]; e2 < pw.row_ptr[r + 1]; ++e2) {
                    const int cl = pw.col_local[e2];
                    if (beg + cl == r) continue;
                    const double vc = V[(int64_t)pw.support_idx[beg + cl] * k + c];
                    yn = fma(ir * (-pw.w[e2] * pw.isd[beg + cl]), vc / nrm, yn);
                    yr = fma(-pw.w[e2], vc, yr);
                }
                ms = fma(vu, vu, ms);
                qn = fma(yn, vu, qn);
                qr = fma(yr, v, qr);
            }
            ms = warp_sum(ms); qn = warp_sum(qn); qr = warp_sum(qr);
            if (lane == 0) {
                mass[(int64_t)c * pw.P + p] = ms;
                quad_norm[(int64_t)c * pw.P + p] = qn;
                quad_raw[(int64_t)c * pw.P + p] = qr;
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// Exact residual ||X - U V^T||_F^2 (verification only; one extra pass over X).  Warp per row.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
residual_kernel(const double* __restrict__ X, int64_t ldx, int64_t m, int n, const double* __restrict__ U,
                const double* __restrict__ V, int k, double* __restrict__ part) {
    __shared__ double scratch[32];
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    for (int64_t row = warp; row < m; row += nwarps) {
        const double* xr = X + row * ldx;
        const double* ur = U + row * k;
        for (int j = lane; j < n; j += 32) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s = fma(ur[l], V[(int64_t)j * k + l], s);
            const double d = xr[j] - s;
            acc = fma(d, d, acc);
        }
    }
    const double t = block_sum(acc, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
}

}  // namespace prmf
