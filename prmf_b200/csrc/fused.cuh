// Single-pass fused X kernel (k <= 10): X is read from HBM ONCE per inner step.
//
//   for every block of 8 samples:   A = X_blk . V          (needs the whole gene axis)
//                                   U_blk <- U_blk * A / (U_blk . Gv + U_blk)
//                                   B += X_blk^T . U_blk    (needs the new U of the block)
//
// The two products are separated by a reduction over ALL genes of a row, so a row block must stay on chip while
// its k dot products are completed.  A row is 54 KB at n = 6750: the gene axis is split into `panels` column
// panels (<= 512 genes) handled by different CTAs, the sample axis into `groups` row groups; grid = panels x
// groups CTAs, all co-resident (<= 1 per SM).  The CTAs of a group walk the same row blocks and exchange, per
// block, their 8 x k partial dot products through a small L2-resident ring with release/acquire flags; every CTA
// then forms the same U_blk (bitwise: fixed panel order) and continues with the outer-product accumulation from
// the tile that is still in its shared memory.
//
// Inside a CTA (13 warps):
//   warp 0        producer: TMA bulk copies of 8 row pieces + the 8 old U rows per stage into a ring
//   warps 1..4    exchange: sum the consumer warps' partials, publish, wait for the other panels, U update
//   warps 5..12   consumers: phase A on FP64 tensor cores (DMMA m8n8k4: 8 samples x 4 genes x 8 factors per
//                 instruction, so the reduction over genes happens in the accumulator fragment, no shuffles;
//                 factors 8,9 by DFMA + one transposed butterfly), phase B by DFMA (thread owns 2 genes x k)
// Phase B of block s-LAG is interleaved with phase A of block s so the exchange latency is off the critical path.
#pragma once
#include "kernels.cuh"

namespace prmf {

constexpr int kFConsWarps = 8;
constexpr int kFExchWarps = 4;
constexpr int kFThreads = (1 + kFExchWarps + kFConsWarps) * 32;   // 416
constexpr int kFRS = 8;            // samples per stage (the DMMA M dimension)
constexpr int kFKP = 10;           // factor pitch of the per-stage 8 x k blocks
constexpr int kFExSlots = 16;      // exchange ring depth per row group (>= 2 x smem ring)
constexpr int kFPaSlots = kFExchWarps;   // partial-sum buffers: one per exchange warp, so nobody waits on a barrier
                                         // that can run two phases ahead
constexpr int kFLag = 2;           // phase B runs this many stages behind phase A
constexpr int kFMaxPanels = 32;
constexpr int kFMaxKSteps = 16;    // 4-gene DMMA steps per consumer warp (panel <= 512 genes)

struct FusedParams {
    const double* X;  int64_t ldx;  int64_t m;  int n;  int k;
    const double* Uold;            // (m + 16) x k
    double* Unew;                  // (m + 16) x k, written by the panel-0 CTAs
    const double* V;               // n x k
    const double* Gv;              // k x k
    int panels, panel_w, groups;   // panel_w % 4 == 0, <= 512
    int64_t rows_per_group;        // % 8 == 0
    int stages;                    // shared-memory ring depth (> kFLag + 1)
    uint32_t pitch;                // bytes between tile rows in shared memory (== 32 mod 128)
    double* Bpart;                 // [groups][n][k]
    double* Gu_part;               // [groups * kFExchWarps][k*k]
    double* Ex;                    // [groups][kFExSlots][panels][kFRS * kFKP]
    unsigned long long* Flags;     // [groups][kFExSlots][panels]
    unsigned long long epoch;      // flags of this launch are epoch + stage + 1
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_cg_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Transposed butterfly: every lane holds NV partial sums; afterwards the lane pair (l, l^1) holds the warp-wide sum
// of value index  bit4*NV/2 + bit3*NV/4 + ...  (fixed order, NV adds + NV shuffles instead of 5 NV).
template <int NV>
__device__ __forceinline__ double butterfly_reduce(double (&v)[NV], int lane) {
    static_assert(NV == 8 || NV == 16, "butterfly_reduce: 8 or 16 values");
    int n = NV;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        if (n > 1) {
            const int half = n >> 1;
            const bool up = lane & m;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < half) {
                    const double send = up ? v[i] : v[i + half];
                    const double keep = up ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                }
            }
            n = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
        }
    }
    return v[0];
}
// value index held by `lane` after butterfly_reduce<NV>
template <int NV>
__device__ __forceinline__ int butterfly_index(int lane) {
    int idx = 0, n = NV;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        if (n > 1) { n >>= 1; if (lane & m) idx += n; }
    }
    return idx;
}

template <int K>
__global__ void __launch_bounds__(kFThreads, 1)
fused_xvu_kernel(const FusedParams p) {
    constexpr int NX = K > 8 ? K - 8 : 0;                 // factors beyond the 8 DMMA columns (0, 1 or 2)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int panel = blockIdx.x, group = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c0 = (int64_t)panel * p.panel_w;
    const int width = (int)min((int64_t)p.panel_w, p.ldx - c0);         // columns present in X (multiple of 4)
    const uint32_t x_stage_bytes = (uint32_t)kFRS * p.pitch;
    const uint32_t u_bytes = (uint32_t)kFRS * K * 8u;
    const uint32_t stage_bytes = x_stage_bytes + ((u_bytes + 127u) & ~127u);
    const int S = p.stages;
    unsigned char* sm_stage = smem_raw;
    double* sPA = reinterpret_cast<double*>(smem_raw + (size_t)S * stage_bytes);   // [kFPaSlots][kFConsWarps][80]
    double* sUn = sPA + kFPaSlots * kFConsWarps * kFRS * kFKP;                                // [S][80]
    double* sGv = sUn + S * kFRS * kFKP;                                              // K*K
    uint64_t* bars = reinterpret_cast<uint64_t*>(sGv + ((K * K + 1) & ~1));
    uint64_t* full_bar = bars;               // [S]  TMA landed
    uint64_t* empty_bar = bars + S;          // [S]  phase B done (8 consumer warps)
    uint64_t* un_bar = bars + 2 * S;         // [S]  U_new of the stage ready (1 exchange warp)
    uint64_t* pa_full = bars + 3 * S;        // [kFPaSlots]  partials written (8 consumer warps)
    uint64_t* pa_empty = pa_full + kFPaSlots;   // [kFPaSlots]  partials consumed (1 exchange warp)

    const int64_t rbeg = (int64_t)group * p.rows_per_group;
    const int64_t rend = min(p.m, rbeg + p.rows_per_group);
    const int ns = rend > rbeg ? (int)((rend - rbeg + kFRS - 1) / kFRS) : 0;

    // zero the whole dynamic region once: tile rows / columns that are never loaded must stay finite
    {
        const size_t words = ((size_t)S * stage_bytes + sizeof(double) * (kFPaSlots * kFConsWarps * kFRS * kFKP + S * kFRS * kFKP)) / 8;
        double* z = reinterpret_cast<double*>(smem_raw);
        for (size_t i = threadIdx.x; i < words; i += blockDim.x) z[i] = 0.0;
    }
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) sGv[i] = p.Gv[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kFConsWarps);
            mbar_init(&un_bar[s], 1);
        }
        for (int s = 0; s < kFPaSlots; ++s) {
            mbar_init(&pa_full[s], kFConsWarps);
            mbar_init(&pa_empty[s], 1);
        }
        fence_mbar_init();
    }
    // make the generic-proxy zero fill visible to the async proxy (TMA writes follow)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // =================================== producer ===================================
    if (warp == 0) {
        const uint64_t pol_x = l2_policy_evict_first();
        const uint64_t pol_u = l2_policy_evict_last();
        const uint32_t row_bytes = (uint32_t)width * 8u;
        for (int s = 0; s < ns; ++s) {
            const int slot = s % S;
            const uint32_t ph = (uint32_t)(s / S) & 1u;
            if (lane == 0) mbar_wait(&empty_bar[slot], ph ^ 1u);
            __syncwarp();
            const int64_t r0 = rbeg + (int64_t)s * kFRS;
            const int rows = (int)min((int64_t)kFRS, rend - r0);
            unsigned char* st = sm_stage + (size_t)slot * stage_bytes;
            if (lane == 0) {
                mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)rows * row_bytes + u_bytes);
                bulk_g2s(st + x_stage_bytes, p.Uold + r0 * K, u_bytes, &full_bar[slot], pol_u);
            }
            __syncwarp();
            if (lane < rows) bulk_g2s(st + (size_t)lane * p.pitch, p.X + (r0 + lane) * p.ldx + c0, row_bytes, &full_bar[slot], pol_x);
        }
        return;
    }

    // =================================== exchange warps ===================================
    if (warp <= kFExchWarps) {
        const int ew = warp - 1;
        constexpr int NV = kFRS * kFKP;                       // 80 values per stage
        double gu[4] = {0.0, 0.0, 0.0, 0.0};                  // this lane's entries of U_new^T U_new (panel 0 only)
        double* ex_base = p.Ex + (size_t)group * kFExSlots * p.panels * NV;
        unsigned long long* fl_base = p.Flags + (size_t)group * kFExSlots * p.panels;
        for (int s = ew; s < ns; s += kFExchWarps) {
            const int pslot = s % kFPaSlots;                  // == ew
            const uint32_t pph = (uint32_t)(s / kFPaSlots) & 1u;
            mbar_wait(&pa_full[pslot], pph);
            // CTA partial = sum of the consumer warps' partials (fixed order)
            double mine[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int idx = lane + 32 * q;
                double a = 0.0;
                if (idx < NV) {
#pragma unroll
                    for (int w = 0; w < kFConsWarps; ++w) a += sPA[(pslot * kFConsWarps + w) * NV + idx];
                }
                mine[q] = a;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pa_empty[pslot]);
            // publish
            const int xs = s % kFExSlots;
            double* ex = ex_base + (size_t)xs * p.panels * NV;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int idx = lane + 32 * q;
                if (idx < NV) ex[(size_t)panel * NV + idx] = mine[q];
            }
            __threadfence();
            __syncwarp();
            const unsigned long long want = p.epoch + (unsigned long long)s + 1ull;
            if (lane == 0) st_release_u64(&fl_base[(size_t)xs * p.panels + panel], want);
            // wait for every panel of the group
            if (lane < p.panels) {
                const unsigned long long* f = &fl_base[(size_t)xs * p.panels + lane];
                while (ld_acquire_u64(f) < want) { }
            }
            __syncwarp();
            // A = sum over panels in panel order (identical in every CTA of the group)
            double A[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int idx = lane + 32 * q;
                double a = 0.0;
                if (idx < NV)
                    for (int pp = 0; pp < p.panels; ++pp) a += (pp == panel) ? mine[q] : ld_cg_f64(ex + (size_t)pp * NV + idx);
                A[q] = a;
            }
            // U update of the 8 rows (:421-422)
            const int slot = s % S;
            mbar_wait(&full_bar[slot], (uint32_t)(s / S) & 1u);          // the old U rows travelled with the tile
            const double* uold = reinterpret_cast<const double*>(sm_stage + (size_t)slot * stage_bytes + x_stage_bytes);
            double* un = sUn + (size_t)slot * NV;
            const int64_t r0 = rbeg + (int64_t)s * kFRS;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int idx = lane + 32 * q;
                if (idx < NV) {
                    const int r = idx / kFKP, c = idx - r * kFKP;
                    double v = 0.0;
                    if (c < K) {
                        const double* ur = uold + r * K;
                        double den = 0.0;
#pragma unroll
                        for (int l = 0; l < K; ++l) den = fma(ur[l], sGv[l * K + c], den);
                        const double u = ur[c];
                        den += u;
                        const double f = (den != 0.0) ? A[q] / den : 1.0;
                        v = u * f;
                        if (panel == 0 && r0 + r < p.m) p.Unew[(r0 + r) * K + c] = v;
                    }
                    un[idx] = v;
                }
            }
            __syncwarp();
            if (panel == 0) {
                const int rows = (int)min((int64_t)kFRS, rend - r0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = lane + 32 * q;
                    if (e < K * K) {
                        const int a = e / K, b = e - a * K;
                        double g = gu[q];
                        for (int r = 0; r < rows; ++r) g = fma(un[r * kFKP + a], un[r * kFKP + b], g);
                        gu[q] = g;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&un_bar[slot]);
        }
        if (panel == 0) {
            double* out = p.Gu_part + ((size_t)group * kFExchWarps + ew) * K * K;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = lane + 32 * q;
                if (e < K * K) out[e] = gu[q];
            }
        }
        return;
    }

    // =================================== consumer warps ===================================
    const int cw = warp - 1 - kFExchWarps;                   // 0..7
    const int t = cw * 32 + lane;                            // 0..255: owns double2 column t (genes c0+2t, c0+2t+1)
    const int H2 = p.panel_w >> 1;                           // double2 columns in the panel
    const bool own = t < H2;
    const int64_t g0 = c0 + 2 * (int64_t)t;
    // phase A fragments: this warp's 4-gene steps
    const int ksteps_total = p.panel_w >> 2;
    const int ks_per_warp = (ksteps_total + kFConsWarps - 1) / kFConsWarps;      // <= kFMaxKSteps
    const int ks0 = cw * ks_per_warp;
    const int arow = lane >> 2, acol = lane & 3;             // A fragment: X[row = lane/4][gene = 4*step + lane%4]
    double bfrag[kFMaxKSteps];                               // B fragment: V[gene = 4*step + lane%4][factor = lane/4]
#pragma unroll
    for (int j = 0; j < kFMaxKSteps; ++j) {
        const int64_t gene = c0 + 4 * (int64_t)(ks0 + j) + acol;
        const bool ok = j < ks_per_warp && (ks0 + j) < ksteps_total && gene < p.n && arow < K;
        bfrag[j] = ok ? p.V[gene * K + arow] : 0.0;
    }
    double vx[2][NX > 0 ? NX : 1];                           // V[own genes][factors 8..]
#pragma unroll
    for (int e = 0; e < (NX > 0 ? NX : 1); ++e) {
        vx[0][e] = (NX > 0 && own && g0 < p.n) ? p.V[g0 * K + 8 + e] : 0.0;
        vx[1][e] = (NX > 0 && own && g0 + 1 < p.n) ? p.V[(g0 + 1) * K + 8 + e] : 0.0;
    }
    double acc[2][K];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int c = 0; c < K; ++c) acc[g][c] = 0.0;

    for (int s = 0; s < ns + kFLag; ++s) {
        if (s < ns) {
            // ---------------- phase A of stage s ----------------
            const int slot = s % S;
            mbar_wait(&full_bar[slot], (uint32_t)(s / S) & 1u);
            const unsigned char* tile = sm_stage + (size_t)slot * stage_bytes;
            const int pslot = s % kFPaSlots;
            mbar_wait(&pa_empty[pslot], ((uint32_t)(s / kFPaSlots) & 1u) ^ 1u);
            double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;  // two accumulator chains
            const unsigned char* arow_ptr = tile + (size_t)arow * p.pitch + ((size_t)(4 * ks0 + acol) << 3);
#pragma unroll
            for (int j = 0; j < kFMaxKSteps; j += 2) {
                if (j < ks_per_warp) {
                    const double a = *reinterpret_cast<const double*>(arow_ptr + (size_t)j * 32);
                    dmma884(d0, d1, a, bfrag[j]);
                }
                if (j + 1 < ks_per_warp) {
                    const double a = *reinterpret_cast<const double*>(arow_ptr + (size_t)(j + 1) * 32);
                    dmma884(e0, e1, a, bfrag[j + 1]);
                }
            }
            d0 += e0; d1 += e1;
            double* pa = sPA + (size_t)(pslot * kFConsWarps + cw) * (kFRS * kFKP);
            {   // D fragment: D[row = lane/4][col = 2*(lane%4) + {0,1}]
                const int col = 2 * acol;
                pa[arow * kFKP + col] = d0;
                pa[arow * kFKP + col + 1] = d1;
            }
            if constexpr (NX > 0) {
                double q[8 * NX];
#pragma unroll
                for (int r = 0; r < kFRS; ++r) {
                    const double2 x = own ? *reinterpret_cast<const double2*>(tile + (size_t)r * p.pitch + ((size_t)t << 4))
                                          : make_double2(0.0, 0.0);
#pragma unroll
                    for (int e = 0; e < NX; ++e) q[r * NX + e] = fma(x.y, vx[1][e], x.x * vx[0][e]);
                }
                constexpr int NVX = 8 * NX;
                const double tot = butterfly_reduce<NVX>(q, lane);
                if ((lane & 1) == 0) {
                    const int idx = butterfly_index<NVX>(lane);           // = r*NX + e
                    pa[(idx / NX) * kFKP + 8 + (idx % NX)] = tot;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pa_full[pslot]);
        }
        if (s >= kFLag) {
            // ---------------- phase B of stage s - LAG ----------------
            const int sb = s - kFLag;
            const int slot = sb % S;
            mbar_wait(&un_bar[slot], (uint32_t)(sb / S) & 1u);
            const unsigned char* tile = sm_stage + (size_t)slot * stage_bytes;
            const double* un = sUn + (size_t)slot * (kFRS * kFKP);
            if (own) {
#pragma unroll
                for (int r = 0; r < kFRS; ++r) {
                    const double2 x = *reinterpret_cast<const double2*>(tile + (size_t)r * p.pitch + ((size_t)t << 4));
#pragma unroll
                    for (int c = 0; c < K; ++c) {
                        const double u = un[r * kFKP + c];
                        acc[0][c] = fma(x.x, u, acc[0][c]);
                        acc[1][c] = fma(x.y, u, acc[1][c]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
        }
    }
    if (own) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int64_t j = g0 + g;
            if (j < p.n) {
                double* out = p.Bpart + ((size_t)group * p.n + j) * K;
#pragma unroll
                for (int c = 0; c < K; ++c) out[c] = acc[g][c];
            }
        }
    }
}

}  // namespace prmf
