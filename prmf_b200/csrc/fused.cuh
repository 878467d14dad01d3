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
// Inside a CTA (16 warps x 128 registers):
//   warp 0        producer: TMA bulk copies of 8 row pieces + the 8 old U rows per stage into a ring
//   warps 1..6    exchange (2 teams x 3 warps, one value per lane): sum the consumer warps' partials, publish
//                 them as tagged words, collect the other panels' words, U update
//   warps 8..15   consumers: phase A on FP64 tensor cores (DMMA m8n8k4: 8 samples x 4 genes x 8 factors per
//                 instruction, so the reduction over genes happens in the accumulator fragment, no shuffles;
//                 factors 8,9 by DFMA + one transposed butterfly), phase B by DFMA (thread owns 2 genes x k)
// Phase B of block s-LAG is interleaved with phase A of block s so the exchange latency is off the critical path.
#pragma once
#include "kernels.cuh"

namespace prmf {

constexpr int kFConsWarps = 8;
constexpr int kFExParts = 3;       // the 80 values of a stage are split over 3 exchange warps (one value per lane)
constexpr int kFExTeams = 2;       // teams of 3 exchange warps alternate over the stages
constexpr int kFExchWarps = kFExParts * kFExTeams;
constexpr int kFFirstCons = 8;     // warps 8..15 are the consumers (warp 0 producer, 1..6 exchange, 7 spare)
constexpr int kFThreads = (kFFirstCons + kFConsWarps) * 32;      // 512 threads x 128 registers = the whole file
constexpr int kFRS = 8;            // samples per stage (the DMMA M dimension)
constexpr int kFKP = 10;           // factor pitch of the per-stage 8 x k blocks
constexpr int kFExSlots = 16;      // exchange ring depth per row group (>= 2 x smem ring)
constexpr int kFPaSlots = kFExTeams;     // partial-sum buffers: one per exchange team, so nobody waits on a barrier
                                         // that can run two phases ahead
constexpr int kFLag = 3;           // phase B runs this many stages behind phase A
constexpr int kFMaxPanels = 16;    // gene panels per row group (n <= 8192)
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
    double* Gu_part;               // [groups][k*k]
    unsigned long long* Ex;        // [groups][kFExSlots][panels][kFRS * kFKP][2]: {32 data bits | 32-bit tag} words
    unsigned long long* Flags;     // unused (kept for layout compatibility)
    unsigned long long epoch;      // tags of this launch are (uint32)(epoch + stage + 1)
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
// "LL" exchange words: an aligned 8-byte store is single-copy atomic, so a word that carries the expected tag also
// carries valid data -- no fence and no separate flag round trip between producer and consumer CTAs.
__device__ __forceinline__ void st_ll(unsigned long long* p, uint32_t data, uint32_t tag) {
    const unsigned long long w = (unsigned long long)data | ((unsigned long long)tag << 32);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void st_ll2(unsigned long long* p, double v, uint32_t tag) {   // p 16-byte aligned
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long lo = (bits & 0xffffffffull) | ((unsigned long long)tag << 32);
    const unsigned long long hi = (bits >> 32) | ((unsigned long long)tag << 32);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void ld_ll2(const unsigned long long* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
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

#ifdef PRMF_FUSED_TIMING
__device__ unsigned long long g_fused_dbg[32];
#define FT_DECL unsigned long long ft_t0 = clock64(), ft_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define FT_MARK(i) do { const unsigned long long now_ = clock64(); ft_acc[i] += now_ - ft_t0; ft_t0 = now_; } while (0)
#define FT_DUMP(base, cond) do { if ((cond) && lane == 0) for (int i_ = 0; i_ < 8; ++i_) g_fused_dbg[(base) + i_] = ft_acc[i_]; } while (0)
#else
#define FT_DECL
#define FT_MARK(i) do { } while (0)
#define FT_DUMP(base, cond) do { } while (0)
#endif

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
            mbar_init(&un_bar[s], kFExParts);
        }
        for (int s = 0; s < kFPaSlots; ++s) {
            mbar_init(&pa_full[s], kFConsWarps);
            mbar_init(&pa_empty[s], kFExParts);
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
        const int team = (warp - 1) / kFExParts, part = (warp - 1) % kFExParts;
        constexpr int NV = kFRS * kFKP;                       // 80 values per stage
        const int idx = part * 32 + lane;                     // the value of this lane
        const bool have = idx < NV;
        const int r = idx / kFKP, c = idx - r * kFKP;
        unsigned long long* ex_base = p.Ex + (size_t)group * kFExSlots * p.panels * NV * 2;
        FT_DECL;
        for (int s = team; s < ns; s += kFExTeams) {
            const int pslot = s % kFPaSlots;                  // == team
            FT_MARK(7);
            mbar_wait(&pa_full[pslot], (uint32_t)(s / kFPaSlots) & 1u);
            FT_MARK(0);
            double mine = 0.0;                                // CTA partial = consumer warps' partials in order
            if (have) {
#pragma unroll
                for (int w = 0; w < kFConsWarps; ++w) mine += sPA[(pslot * kFConsWarps + w) * NV + idx];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pa_empty[pslot]);
            const int xs = s % kFExSlots;
            unsigned long long* ex = ex_base + (size_t)xs * p.panels * NV * 2;
            const uint32_t tag = (uint32_t)(p.epoch + (unsigned long long)s + 1ull);
            if (have) st_ll2(ex + ((size_t)panel * NV + idx) * 2, mine, tag);
            FT_MARK(1);
            // A = sum over the panels in panel order (bitwise identical in every CTA of the group); one round of
            // loads, words whose tag has not arrived yet are polled again
            double A = 0.0;
            if (have) {
                double vals[kFMaxPanels];
                uint32_t pending = 0;
#pragma unroll
                for (int pp = 0; pp < kFMaxPanels; ++pp) {
                    vals[pp] = 0.0;
                    if (pp < p.panels && pp != panel) pending |= 1u << pp;
                }
                while (pending) {
                    unsigned long long lo[kFMaxPanels], hi[kFMaxPanels];
#pragma unroll
                    for (int pp = 0; pp < kFMaxPanels; ++pp)
                        if (pending & (1u << pp)) ld_ll2(ex + ((size_t)pp * NV + idx) * 2, lo[pp], hi[pp]);
#pragma unroll
                    for (int pp = 0; pp < kFMaxPanels; ++pp)
                        if (pending & (1u << pp)) {
                            if ((uint32_t)(lo[pp] >> 32) == tag && (uint32_t)(hi[pp] >> 32) == tag) {
                                vals[pp] = __longlong_as_double((long long)((lo[pp] & 0xffffffffull) | (hi[pp] << 32)));
                                pending &= ~(1u << pp);
                            }
                        }
                }
#pragma unroll
                for (int pp = 0; pp < kFMaxPanels; ++pp)
                    if (pp < p.panels) A += (pp == panel) ? mine : vals[pp];
            }
            FT_MARK(2);
            // U update of this lane's entry (:421-422)
            const int slot = s % S;
            mbar_wait(&full_bar[slot], (uint32_t)(s / S) & 1u);          // the old U rows travelled with the tile
            const double* uold = reinterpret_cast<const double*>(sm_stage + (size_t)slot * stage_bytes + x_stage_bytes);
            double* un = sUn + (size_t)slot * NV;
            const int64_t r0 = rbeg + (int64_t)s * kFRS;
            if (have) {
                double v = 0.0;
                if (c < K) {
                    const double* ur = uold + r * K;
                    double den = 0.0;
#pragma unroll
                    for (int l = 0; l < K; ++l) den = fma(ur[l], sGv[l * K + c], den);
                    const double u = ur[c];
                    den += u;
                    const double f = (den != 0.0) ? A / den : 1.0;
                    v = u * f;
                    if (panel == 0 && r0 + r < p.m) p.Unew[(r0 + r) * K + c] = v;
                }
                un[idx] = v;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&un_bar[slot]);
            FT_MARK(3);
        }
        FT_DUMP(0, panel == 0 && group == 0 && warp == 1);
        return;
    }
    if (warp < kFFirstCons) return;                           // spare warp

    // =================================== consumer warps ===================================
    const int cw = warp - kFFirstCons;                       // 0..7
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
    double gu = 0.0;

    FT_DECL;
    for (int s = 0; s < ns + kFLag; ++s) {
        FT_MARK(7);
        if (s < ns) {
            // ---------------- phase A of stage s ----------------
            const int slot = s % S;
            mbar_wait(&full_bar[slot], (uint32_t)(s / S) & 1u);
            FT_MARK(0);
            const unsigned char* tile = sm_stage + (size_t)slot * stage_bytes;
            const int pslot = s % kFPaSlots;
            mbar_wait(&pa_empty[pslot], ((uint32_t)(s / kFPaSlots) & 1u) ^ 1u);
            FT_MARK(1);
            // 8 independent accumulator chains of 2 DMMAs: a dependent DMMA has a latency of a few hundred cycles
            constexpr int NCH = 8;
            double dch[NCH][2];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) { dch[ch][0] = 0.0; dch[ch][1] = 0.0; }
            const unsigned char* arow_ptr = tile + (size_t)arow * p.pitch + ((size_t)(4 * ks0 + acol) << 3);
            double afrag[kFMaxKSteps];
#pragma unroll
            for (int j = 0; j < kFMaxKSteps; ++j)
                afrag[j] = (j < ks_per_warp) ? *reinterpret_cast<const double*>(arow_ptr + (size_t)j * 32) : 0.0;
#pragma unroll
            for (int j = 0; j < kFMaxKSteps; ++j)
                if (j < ks_per_warp) dmma884(dch[j % NCH][0], dch[j % NCH][1], afrag[j], bfrag[j]);
            double d0 = ((dch[0][0] + dch[1][0]) + (dch[2][0] + dch[3][0])) + ((dch[4][0] + dch[5][0]) + (dch[6][0] + dch[7][0]));
            double d1 = ((dch[0][1] + dch[1][1]) + (dch[2][1] + dch[3][1])) + ((dch[4][1] + dch[5][1]) + (dch[6][1] + dch[7][1]));
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
            FT_MARK(2);
        }
        if (s >= kFLag) {
            // ---------------- phase B of stage s - LAG ----------------
            const int sb = s - kFLag;
            const int slot = sb % S;
            mbar_wait(&un_bar[slot], (uint32_t)(sb / S) & 1u);
            FT_MARK(3);
            const unsigned char* tile = sm_stage + (size_t)slot * stage_bytes;
            const double* un = sUn + (size_t)slot * (kFRS * kFKP);
            if (own) {
#pragma unroll
                for (int r = 0; r < kFRS; ++r) {
                    const double2 x = *reinterpret_cast<const double2*>(tile + (size_t)r * p.pitch + ((size_t)t << 4));
                    const double2* u2 = reinterpret_cast<const double2*>(un + r * kFKP);     // 80-byte rows: 16-byte aligned
#pragma unroll
                    for (int c2 = 0; c2 < (K + 1) / 2; ++c2) {
                        const double2 u = u2[c2];
                        acc[0][2 * c2] = fma(x.x, u.x, acc[0][2 * c2]);
                        acc[1][2 * c2] = fma(x.y, u.x, acc[1][2 * c2]);
                        if (2 * c2 + 1 < K) {
                            acc[0][2 * c2 + 1] = fma(x.x, u.y, acc[0][2 * c2 + 1]);
                            acc[1][2 * c2 + 1] = fma(x.y, u.y, acc[1][2 * c2 + 1]);
                        }
                    }
                }
            }
            if (panel == 0 && t < K * K) {                    // U_new^T U_new (:425): one entry per thread
                const int ga = t / K, gb = t - ga * K;
                const int rows = (int)min((int64_t)kFRS, rend - (rbeg + (int64_t)sb * kFRS));
                for (int r = 0; r < rows; ++r) gu = fma(un[r * kFKP + ga], un[r * kFKP + gb], gu);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
            FT_MARK(4);
        }
    }
    FT_DUMP(8, panel == 0 && group == 0 && cw == 0);
    FT_DUMP(16, panel == 5 && group == 3 && cw == 3);
    if (panel == 0 && t < K * K) p.Gu_part[(size_t)group * K * K + t] = gu;
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
