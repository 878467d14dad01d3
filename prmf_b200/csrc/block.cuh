// block.cuh -- the persistent step kernel (k <= 10, fp64): a whole block of inner steps in ONE cooperative launch.
//
// An inner step (reference prmf_runner.py:419-444) is two passes over X separated by full reductions:
//   half-step "pass 1":  A = X.V  (streams the transposed copy Xt)  -> U update  (:420-422) + partials of U_new^T U_new
//   half-step "pass 2":  B = X^T.U_new (streams X)                  -> V update  (:424-444) + what the objective needs
// skinny_tma_kernel runs each half as its own launch; every launch pays the ramp-up of its TMA ring, its fused tail
// with the SMs idle on HBM, and a kernel boundary.  Here one CTA per SM stays resident for `nh` consecutive halves:
//   * the producer warp runs ahead across the half boundary: while the 8 consumer warps are still in the tail of
//     half h, it has already refilled the whole ring with the X tiles of half h+1 (they depend on nothing); only
//     the small W operand (rows of V / U_new) waits for the data dependency, a monotone counter in L2;
//   * kernel boundaries are replaced by those counters: `udone` (all sample panels updated and folded) gates the
//     W rows of pass 2 and the V update; `vdone` gates the W rows of the next pass 1 and the next U update;
//   * on several GPUs the sum over ranks of [X^T U | U^T U] is part of the pass-2 tail: every CTA pushes the sums of
//     its share of genes straight into every peer's receive buffer (NVLink stores), raises a flag there, waits for
//     the same share from every peer and adds the ranks' contributions in rank order from LOCAL memory -- one
//     one-way hop, bitwise identical on all ranks, no collective launch.  U^T U is pushed at the START of pass 2
//     (it is complete after pass 1), so it has the whole pass to arrive.
// Every spin wait has a %globaltimer deadline: on expiry the kernel sets an error word, stops waiting and drains;
// the host reports PRMF_ERR_TIMEOUT instead of hanging (dead peer, lost launch).
// The objective is deferred exactly as on the two-launch path: every step leaves its Gram partials, U^T U and the
// active-set values of V_new in per-step slots; objective_deferred_kernel evaluates the block afterwards.
#pragma once

#include "kernels.cuh"

namespace prmf {

constexpr int kBlkRS = 8;            // rows of M per ring stage (as skinny_tma_kernel<.,8,.>)
constexpr int kBlkTile = 64;         // rows of a share updated per sub-tile of the tail (bounds the scratch)

constexpr unsigned int kErrTimeoutLocal = 1u;   // a wait on another CTA of this GPU expired
constexpr unsigned int kErrTimeoutPeer = 2u;    // a wait on a peer GPU's flag expired

struct BlockParams {
    // X (m x n, leading dimension ldx) and its transposed copy (n x m, ldxt)
    const double* X;
    const double* Xt;
    int64_t ldx, ldxt, m, n;
    // pass 1: column panels over samples, row chunks over genes; pass 2: panels over genes, chunks over samples
    int panels1, panel_w1, chunks1;
    int panels2, panel_w2, chunks2;
    int64_t rpc1, rpc2;
    int stages;
    uint32_t ring_stage_bytes;       // max over the two passes of (RS * panel_w * 8 + W rows, 128-byte padded)
    // state: U[0] / V[0] are current at launch; every pass 1 flips U, every pass 2 flips V
    double* U[2];
    double* V[2];
    double* Apart;                   // [chunks1][m][K]
    double* Bpart;                   // [chunks2][n][K]
    double* Gu_part;                 // [panels1][K*K]   U_new^T U_new per sample panel (this rank's rows)
    double* part2;                   // scratch [tiles][K*K] per-CTA Gram partials
    double* vb2;                     // scratch [tiles]
    const double* Gv0;               // V^T V of the V at block start (for the U update of half 0)
    // monotone counters (never reset): arrive / done per panel, and the two grid-wide ones
    unsigned long long* arrive1;
    unsigned long long* done1;
    unsigned long long* arrive2;
    unsigned long long* done2;
    unsigned long long* udone;       // += 1 per folded sample panel
    unsigned long long* vdone;       // += 1 per folded gene panel
    unsigned long long base1, base2; // pass-1 / pass-2 executions of this kernel on this handle before this launch
    int h0, nh;                      // halves [h0, h0 + nh) of the block: even = pass 1, odd = pass 2; step = half / 2
    // V update
    Pathways pw;
    const int32_t* active;
    const int32_t* pos;
    const double* gd;
    // deferred objective: per-step slots
    double* hist_Gu;                 // [steps][K*K]
    double* hist_Gvp;                // [steps][panels2][K*K]
    double* hist_VBp;                // [steps][panels2]
    double* hist_vh;                 // [steps][kVhCap]
    const int64_t* doff;
    // bounded waits
    unsigned int* err;
    unsigned long long timeout_ns;
    // exchange over ranks (nranks <= 1: none)
    int nranks, rank;
    double* xbuf[kMaxPeers];         // rank r's receive buffer: [parity 2][src rank kMaxPeers][xcount]
    unsigned long long* xflag[kMaxPeers];   // rank r's flags: [src rank kMaxPeers][tiles2 + 1]
    size_t xcount;                   // doubles per (parity, src) slot: n*K + K*K, padded
    unsigned long long xbase;        // exchanges (= pass-2 executions of this kernel) before this launch
};

__device__ __forceinline__ unsigned long long blk_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Spin until *p >= target (acquire; SYS: the writer is a peer GPU).  Gives up when the deadline passes or when
// another waiter of this kernel already gave up (err != 0), so a failed launch drains instead of hanging.
template <bool SYS>
__device__ __forceinline__ bool blk_wait_ge(const unsigned long long* p, unsigned long long target, unsigned int* err,
                                            unsigned long long timeout_ns) {
    unsigned long long t0 = 0;
    unsigned int polls = 0;
    for (;;) {
        const unsigned long long v = SYS ? ld_acquire_sys_u64(p) : ld_acquire_gpu_u64(p);
        if (v >= target) return true;
        if ((++polls & 255u) == 0u) {
            const unsigned long long now = blk_gtime();
            if (t0 == 0) t0 = now;
            if (now - t0 > timeout_ns || ld_volatile_u32(err) != 0u) {
                atomicOr(err, SYS ? kErrTimeoutPeer : kErrTimeoutLocal);
                return false;
            }
        }
    }
}

// per-thread running Gram accumulation over the sub-tiles of a share (thread = (pair a,b ; row slice))
template <int K>
struct BlkGram {
    static constexpr int NP = K * K;
    static constexpr int NS = (256 / NP) > 8 ? 8 : (256 / NP);
    double g = 0.0;
    __device__ __forceinline__ void add(const double* __restrict__ sT, int rows) {
        const int t = threadIdx.x;
        if (t < NP * NS) {
            const int pair = t % NP, sl = t / NP;
            const int a = pair / K, b = pair - a * K;
            const int rps = (rows + NS - 1) / NS;
            const int rb = sl * rps, re = min(rows, rb + rps);
            for (int r = rb; r < re; ++r) g = fma(sT[r * K + a], sT[r * K + b], g);
        }
    }
    // slice sums in order -> out[K*K]   (all 256 consumer threads call this)
    __device__ __forceinline__ void finish(double* __restrict__ sBuf, double* __restrict__ out) {
        const int t = threadIdx.x;
        if (t < NP * NS) sBuf[t] = g;             // sBuf[sl * NP + pair]
        cons_bar();
        if (t < NP) {
            double s = 0.0;
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) s += sBuf[sl * NP + t];
            out[t] = s;
        }
    }
};

struct BlkTile {          // this CTA's tile of a pass
    bool in;              // CTA takes part in the pass
    int panel, chunk;
    int64_t c0;           // first column of the panel
    int width;            // columns held (multiple of 4)
    int64_t rbeg, rend;   // row range of the chunk
    int iters;            // ring stages this tile streams
};

__device__ __forceinline__ BlkTile blk_tile(int b, int panels, int panel_w, int chunks, int64_t rpc, int64_t rows_total,
                                            int64_t ldm) {
    BlkTile tl;
    tl.in = b < panels * chunks;
    tl.panel = b % panels;
    tl.chunk = b / panels;
    tl.c0 = (int64_t)tl.panel * panel_w;
    tl.width = (int)max((int64_t)0, min((int64_t)panel_w, ldm - tl.c0));
    tl.rbeg = (int64_t)tl.chunk * rpc;
    tl.rend = min(rows_total, tl.rbeg + rpc);
    tl.iters = (tl.in && tl.rend > tl.rbeg) ? (int)((tl.rend - tl.rbeg + kBlkRS - 1) / kBlkRS) : 0;
    return tl;
}

// rows [rb, rb + rows) of the panel's columns are this CTA's share of the update
__device__ __forceinline__ void blk_share(const BlkTile& tl, int chunks, int64_t cols, int& rb, int& rows) {
    const int all = (int)max((int64_t)0, min((int64_t)tl.width, cols - tl.c0));
    const int per = (all + chunks - 1) / chunks;
    rb = tl.chunk * per;
    rows = max(0, min(all, rb + per) - rb);
}

template <int K>
__global__ void __launch_bounds__(kTmaThreads, 1)
block_kernel(const BlockParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stages = p.stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * p.ring_stage_bytes);
    uint64_t* empty_bar = full_bar + stages;
    double* scratch = reinterpret_cast<double*>(empty_bar + stages);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    constexpr uint32_t w_bytes = (uint32_t)kBlkRS * K * 8u;
    __shared__ int s_last;

    if (threadIdx.x == 0) {
        for (int s2 = 0; s2 < stages; ++s2) {
            mbar_init(&full_bar[s2], 1);
            mbar_init(&empty_bar[s2], kTmaConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const BlkTile t1 = blk_tile(b, p.panels1, p.panel_w1, p.chunks1, p.rpc1, p.n, p.ldxt);   // M = Xt: n rows, m columns
    const BlkTile t2 = blk_tile(b, p.panels2, p.panel_w2, p.chunks2, p.rpc2, p.m, p.ldx);    // M = X:  m rows, n columns

    if (warp == kTmaConsumerWarps) {
        // ================================= producer warp =================================
        const uint64_t pol_x = l2_policy_evict_first();
        const uint64_t pol_w = l2_policy_evict_last();
        int s2 = 0;
        uint32_t phase = 0;
        unsigned long long n1 = p.base1, n2 = p.base2;
        int ui = 0, vi = 0;                               // current U / V buffer
        for (int i = 0; i < p.nh; ++i) {
            const bool pass1 = ((p.h0 + i) & 1) == 0;
            const BlkTile& tl = pass1 ? t1 : t2;
            const double* M = pass1 ? p.Xt : p.X;
            const int64_t ldm = pass1 ? p.ldxt : p.ldx;
            const int panel_w = pass1 ? p.panel_w1 : p.panel_w2;
            // W of pass 1: the current V (that pass writes U[ui ^ 1] and flips ui); W of pass 2: the U this step's
            // pass 1 has just produced (that pass writes V[vi ^ 1] and flips vi)
            const double* Wsrc;
            if (pass1) { ++n1; Wsrc = p.V[vi]; ui ^= 1; }
            else { ++n2; Wsrc = p.U[ui]; vi ^= 1; }
            if (tl.iters > 0) {
                const uint32_t row_bytes = (uint32_t)tl.width * 8u;
                const uint32_t x_stage_bytes = (uint32_t)kBlkRS * (uint32_t)panel_w * 8u;
                // the data dependency of the W rows: every panel of the previous half folded
                const unsigned long long* dep = pass1 ? p.vdone : p.udone;
                const unsigned long long dep_target = pass1 ? (n2 * (unsigned long long)p.panels2) : (n1 * (unsigned long long)p.panels1);
                const int pre = min(stages, tl.iters);
                // X tiles of the first `pre` stages: no dependency, they go out as soon as the ring slots are free
                {
                    int s3 = s2;
                    uint32_t ph3 = phase;
                    for (int it = 0; it < pre; ++it) {
                        const int64_t r0 = tl.rbeg + (int64_t)it * kBlkRS;
                        const int rows = (int)min((int64_t)kBlkRS, tl.rend - r0);
                        if (lane == 0) {
                            mbar_wait(&empty_bar[s3], ph3 ^ 1u);
                            mbar_arrive_expect_tx(&full_bar[s3], (uint32_t)rows * row_bytes + w_bytes);
                        }
                        __syncwarp();
                        unsigned char* sx = smem_raw + (size_t)s3 * p.ring_stage_bytes;
                        if (lane < rows)
                            bulk_g2s(sx + (size_t)lane * panel_w * 8, M + (r0 + lane) * ldm + tl.c0, row_bytes, &full_bar[s3], pol_x);
                        if (++s3 == stages) { s3 = 0; ph3 ^= 1u; }
                    }
                }
                if (lane == 0) blk_wait_ge<false>(dep, dep_target, p.err, p.timeout_ns);
                __syncwarp();
                for (int it = 0; it < pre; ++it) {
                    const int64_t r0 = tl.rbeg + (int64_t)it * kBlkRS;
                    unsigned char* sx = smem_raw + (size_t)s2 * p.ring_stage_bytes;
                    if (lane == 0) bulk_g2s(sx + x_stage_bytes, Wsrc + r0 * K, w_bytes, &full_bar[s2], pol_w);
                    if (++s2 == stages) { s2 = 0; phase ^= 1u; }
                }
                for (int it = pre; it < tl.iters; ++it) {
                    const int64_t r0 = tl.rbeg + (int64_t)it * kBlkRS;
                    const int rows = (int)min((int64_t)kBlkRS, tl.rend - r0);
                    if (lane == 0) mbar_wait(&empty_bar[s2], phase ^ 1u);
                    __syncwarp();
                    unsigned char* sx = smem_raw + (size_t)s2 * p.ring_stage_bytes;
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full_bar[s2], (uint32_t)rows * row_bytes + w_bytes);
                        bulk_g2s(sx + x_stage_bytes, Wsrc + r0 * K, w_bytes, &full_bar[s2], pol_w);
                    }
                    __syncwarp();
                    if (lane < rows)
                        bulk_g2s(sx + (size_t)lane * panel_w * 8, M + (r0 + lane) * ldm + tl.c0, row_bytes, &full_bar[s2], pol_x);
                    if (++s2 == stages) { s2 = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ================================= consumer warps =================================
    const int t = threadIdx.x;                             // 0..255
    double* sG = scratch;                                  // K*K Gram of the other factor matrix (padded to 128)
    double* sW8 = sG + 112;                                // 8 warp sums
    double* sBuf = sG + 128;                               // 8 * K*K slice sums
    double* sT = sBuf + 8 * K * K;                         // kBlkTile x K new rows
    double* sS = sT + kBlkTile * K;                        // kBlkTile x K sums of the pass partials
    double* sO = sS + kBlkTile * K;                        // kBlkTile x K old rows
    int s2 = 0;
    uint32_t phase = 0;
    unsigned long long n1 = p.base1, n2 = p.base2;
    int ui = 0, vi = 0;
    for (int i = 0; i < p.nh; ++i) {
        const int half = p.h0 + i;
        const bool pass1 = (half & 1) == 0;
        const int step = half >> 1;
        const BlkTile& tl = pass1 ? t1 : t2;
        const int panel_w = pass1 ? p.panel_w1 : p.panel_w2;
        const int chunks = pass1 ? p.chunks1 : p.chunks2;
        const int64_t cols = pass1 ? p.m : p.n;            // columns of M = rows of the matrix being updated
        double* OutPart = pass1 ? p.Apart : p.Bpart;
        if (pass1) ++n1; else ++n2;

        // on several GPUs CTA 0 pushes this rank's U^T U to every peer at the start of pass 2 (complete since pass 1)
        if (!pass1 && p.nranks > 1 && b == 0) {
            if (t == 0) blk_wait_ge<false>(p.udone, n1 * (unsigned long long)p.panels1, p.err, p.timeout_ns);
            cons_bar();
            const unsigned long long xs = p.xbase + (n2 - p.base2);
            const size_t slot = ((size_t)(xs & 1ull) * kMaxPeers + p.rank) * p.xcount + (size_t)p.n * K;
            if (t < K * K) {
                const double g = sum_strided_cg(p.Gu_part + t, p.panels1, K * K);
                for (int r = 0; r < p.nranks; ++r) p.xbuf[r][slot + t] = g;
            }
            __threadfence_system();
            cons_bar();
            if (t < p.nranks)
                st_release_sys_u64(p.xflag[t] + (size_t)p.rank * (p.panels2 * p.chunks2 + 1) + p.panels2 * p.chunks2, xs);
        }

        // ---- main loop: acc[g][c] += M[row][col g] * W[row][c] over the chunk's rows ----
        if (tl.in) {
            const int H = panel_w >> 2;
            const bool act = t < H && (2 * t) < tl.width;
            const bool act2 = t < H && (2 * (t + H)) < tl.width;
            const uint32_t x_stage_bytes = (uint32_t)kBlkRS * (uint32_t)panel_w * 8u;
            double acc[4][K];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int c = 0; c < K; ++c) acc[g][c] = 0.0;
            for (int it = 0; it < tl.iters; ++it) {
                const int rows = (int)min((int64_t)kBlkRS, tl.rend - (tl.rbeg + (int64_t)it * kBlkRS));
                mbar_wait(&full_bar[s2], phase);
                const unsigned char* sx = smem_raw + (size_t)s2 * p.ring_stage_bytes;
                const double* sw = reinterpret_cast<const double*>(sx + x_stage_bytes);
                if (rows == kBlkRS) {
#pragma unroll
                    for (int r = 0; r < kBlkRS; ++r) {
                        const double2* xrow = reinterpret_cast<const double2*>(sx + (size_t)r * panel_w * 8);
                        const double2 xa = act ? xrow[t] : make_double2(0.0, 0.0);
                        const double2 xb = act2 ? xrow[t + H] : make_double2(0.0, 0.0);
#pragma unroll
                        for (int c = 0; c < K; ++c) {
                            const double u = sw[r * K + c];
                            acc[0][c] = fma(xa.x, u, acc[0][c]);
                            acc[1][c] = fma(xa.y, u, acc[1][c]);
                            acc[2][c] = fma(xb.x, u, acc[2][c]);
                            acc[3][c] = fma(xb.y, u, acc[3][c]);
                        }
                    }
                } else {
                    for (int r = 0; r < rows; ++r) {
                        const double2* xrow = reinterpret_cast<const double2*>(sx + (size_t)r * panel_w * 8);
                        const double2 xa = act ? xrow[t] : make_double2(0.0, 0.0);
                        const double2 xb = act2 ? xrow[t + H] : make_double2(0.0, 0.0);
#pragma unroll
                        for (int c = 0; c < K; ++c) {
                            const double u = sw[r * K + c];
                            acc[0][c] = fma(xa.x, u, acc[0][c]);
                            acc[1][c] = fma(xa.y, u, acc[1][c]);
                            acc[2][c] = fma(xb.x, u, acc[2][c]);
                            acc[3][c] = fma(xb.y, u, acc[3][c]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s2]);
                if (++s2 == stages) { s2 = 0; phase ^= 1u; }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int64_t j = tl.c0 + 2 * (int64_t)(g < 2 ? t : t + H) + (g & 1);
                const bool ok = (g < 2 ? act : act2) && j < cols;
                if (ok) {
                    double* out = OutPart + ((int64_t)tl.chunk * cols + j) * K;
#pragma unroll
                    for (int c = 0; c < K; ++c) out[c] = acc[g][c];
                }
            }
        }

        // ---- tail: this CTA's share of the panel's rows ----
        if (pass1) {
            const double* Uold = p.U[ui];
            double* Unew = p.U[ui ^ 1];
            ui ^= 1;
            if (!tl.in) continue;
            int rb, rows;
            blk_share(tl, chunks, cols, rb, rows);
            // Gv of the current V: the published one for the first half of a block, else the per-panel partials the
            // previous step's V update left (complete once vdone reached the previous pass 2)
            if (half == 0) {
                if (t < K * K) sG[t] = p.Gv0[t];
            } else {
                if (t == 0) blk_wait_ge<false>(p.vdone, n2 * (unsigned long long)p.panels2, p.err, p.timeout_ns);
                cons_bar();
                if (t < K * K)
                    sG[t] = sum_strided_cg(p.hist_Gvp + (size_t)(step - 1) * p.panels2 * K * K + t, p.panels2, K * K);
            }
            // every CTA of the panel has stored its partials
            __threadfence();
            cons_bar();
            if (t == 0) {
                atomicAdd(p.arrive1 + tl.panel, 1ull);
                blk_wait_ge<false>(p.arrive1 + tl.panel, n1 * (unsigned long long)chunks, p.err, p.timeout_ns);
            }
            cons_bar();
            BlkGram<K> gram;
            for (int r0 = 0; r0 < rows; r0 += kBlkTile) {
                const int nr = min(kBlkTile, rows - r0);
                const int n_el = nr * K;
                const int64_t base = (tl.c0 + rb + r0) * K;
                for (int e = t; e < n_el; e += 256) sO[e] = __ldcg(Uold + base + e);
                epi_stage_sums(sS, p.Apart + base, n_el, chunks, cols * K);                      // X.V   (:420)
                cons_bar();
                for (int e = t; e < n_el; e += 256) {
                    const int r = e / K, c = e - r * K;
                    const double* urow = sO + r * K;
                    double den = 0.0;
#pragma unroll
                    for (int l = 0; l < K; ++l) den = fma(urow[l], sG[l * K + c], den);
                    const double u = urow[c];
                    den += u;
                    const double f = (den != 0.0) ? sS[e] / den : 1.0;                           // 0/0 := 1 (:422)
                    const double un = u * f;
                    sT[e] = un;
                    Unew[base + e] = un;
                }
                cons_bar();
                gram.add(sT, nr);
                cons_bar();
            }
            const int64_t me = (int64_t)tl.panel * chunks + tl.chunk;
            gram.finish(sBuf, p.part2 + me * K * K);
            __threadfence();
            cons_bar();
            if (t == 0) {
                const unsigned long long prev = atomicAdd(p.done1 + tl.panel, 1ull);
                s_last = prev + 1ull == n1 * (unsigned long long)chunks;
            }
            cons_bar();
            if (s_last) {                                   // last CTA of the panel: fold the panel's Gram partials
                __threadfence();
                if (t < K * K)
                    p.Gu_part[(int64_t)tl.panel * K * K + t] =
                        sum_strided_cg(p.part2 + (int64_t)tl.panel * chunks * K * K + t, chunks, K * K);
                __threadfence();
                cons_bar();
                if (t == 0) atomicAdd(p.udone, 1ull);
            }
            cons_bar();                                     // s_last is rewritten by the next tail
        } else {
            const double* Vold = p.V[vi];
            double* Vnew = p.V[vi ^ 1];
            vi ^= 1;
            if (!tl.in) continue;
            int rb, rows;
            blk_share(tl, chunks, cols, rb, rows);
            const int tiles2 = p.panels2 * p.chunks2;
            const unsigned long long xs = p.xbase + (n2 - p.base2);                 // this exchange's sequence number
            const size_t xpar = (size_t)(xs & 1ull) * kMaxPeers;
            // Gu = U_new^T U_new: sum over this rank's sample panels (complete once udone reached this step's pass 1),
            // then over ranks
            if (t == 0) blk_wait_ge<false>(p.udone, n1 * (unsigned long long)p.panels1, p.err, p.timeout_ns);
            cons_bar();
            if (p.nranks <= 1) {
                if (t < K * K) {
                    const double g = sum_strided_cg(p.Gu_part + t, p.panels1, K * K);
                    sG[t] = g;
                    if (b == 0) p.hist_Gu[(size_t)step * K * K + t] = g;
                }
            } else {
                if (t < p.nranks)
                    blk_wait_ge<true>(p.xflag[p.rank] + (size_t)t * (tiles2 + 1) + tiles2, xs, p.err, p.timeout_ns);
                cons_bar();
                if (t < K * K) {
                    double g = 0.0;
                    for (int r = 0; r < p.nranks; ++r) g += __ldcg(p.xbuf[p.rank] + (xpar + r) * p.xcount + (size_t)p.n * K + t);
                    sG[t] = g;
                    if (b == 0) p.hist_Gu[(size_t)step * K * K + t] = g;
                }
            }
            const double gamma = p.gd[0], delta = p.gd[1];
            __threadfence();
            cons_bar();
            if (t == 0) {
                atomicAdd(p.arrive2 + tl.panel, 1ull);
                blk_wait_ge<false>(p.arrive2 + tl.panel, n2 * (unsigned long long)chunks, p.err, p.timeout_ns);
            }
            cons_bar();
            BlkGram<K> gram;
            double vb = 0.0;
            int sub = 0;
            const int nsub = (rows + kBlkTile - 1) / kBlkTile;
            for (int r0 = 0; r0 < rows; r0 += kBlkTile, ++sub) {
                const int nr = min(kBlkTile, rows - r0);
                const int n_el = nr * K;
                const int64_t base = (tl.c0 + rb + r0) * K;
                for (int e = t; e < n_el; e += 256) sO[e] = __ldcg(Vold + base + e);
                epi_stage_sums(sS, p.Bpart + base, n_el, chunks, cols * K);                      // local X^T U  (:424)
                cons_bar();
                if (p.nranks > 1) {
                    // push this rank's sums of the sub-tile into every rank's receive slot, flag, wait for all, add in order
                    const unsigned long long fs = (xs - 1) * (unsigned long long)nsub + sub + 1;   // per-CTA flag sequence
                    for (int e = t; e < n_el; e += 256) {
                        const double v = sS[e];
                        for (int r = 0; r < p.nranks; ++r) p.xbuf[r][(xpar + p.rank) * p.xcount + base + e] = v;
                    }
                    __threadfence_system();
                    cons_bar();
                    if (t < p.nranks) {
                        st_release_sys_u64(p.xflag[t] + (size_t)p.rank * (tiles2 + 1) + b, fs);
                        blk_wait_ge<true>(p.xflag[p.rank] + (size_t)t * (tiles2 + 1) + b, fs, p.err, p.timeout_ns);
                    }
                    cons_bar();
                    for (int e = t; e < n_el; e += 256) {
                        double s = 0.0;
                        for (int r = 0; r < p.nranks; ++r) s += __ldcg(p.xbuf[p.rank] + (xpar + r) * p.xcount + base + e);
                        sS[e] = s;
                    }
                    cons_bar();
                }
                for (int e = t; e < n_el; e += 256) {
                    const int r = e / K, c = e - r * K;
                    const int64_t j = tl.c0 + rb + r0 + r;
                    const double* vrow = sO + r * K;
                    const double bb = sS[e];
                    double cden = 0.0;
#pragma unroll
                    for (int l = 0; l < K; ++l) cden = fma(vrow[l], sG[l * K + c], cden);       // V.Gu   (:425)
                    const double v = vrow[c];
                    double num = bb, den = cden;
                    const int32_t pr = p.pos[j * K + c];
                    if (pr >= 0) {
                        const Pathways& pw = p.pw;
                        const int64_t pbase = pw.path_ptr[p.active[c]];
                        double wv = 0.0;
                        for (int64_t e2 = pw.row_ptr[pr]; e2 < pw.row_ptr[pr + 1]; ++e2)
                            wv = fma(pw.w[e2], __ldcg(Vold + (int64_t)pw.support_idx[pbase + pw.col_local[e2]] * K + c), wv);   // rows other CTAs wrote in this launch: L2
                        const double vp1 = v + 1.0;
                        const double man = gamma * wv;                                           // :434
                        const double ign = delta * (1.0 / (vp1 * vp1));                          // :438
                        num = bb + (man + ign);                                                  // :440
                        den = cden + gamma * (pw.deg[pr] * v);                                   // :435,:441
                        // (hist_vh below)
                    }
                    if (den < kEps) den = kEps;                                                  // :442
                    double vn = v * (num / den);                                                 // :443
                    if (vn < kEps) vn = kEps;                                                    // :444
                    sT[e] = vn;
                    Vnew[j * K + c] = vn;
                    if (pr >= 0) p.hist_vh[(size_t)step * kVhCap + p.doff[c] + (pr - p.pw.path_ptr[p.active[c]])] = vn;
                    vb = fma(vn, bb, vb);
                }
                cons_bar();
                gram.add(sT, nr);
                cons_bar();
            }
            // sum(V_new * B) of this share: warp sums, then the 8 warp sums in order
            vb = warp_sum(vb);
            if ((t & 31) == 0) sW8[t >> 5] = vb;
            cons_bar();
            const int64_t me = (int64_t)tl.panel * chunks + tl.chunk;
            if (t == 0) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += sW8[w];
                p.vb2[me] = s;
            }
            gram.finish(sBuf, p.part2 + me * K * K);
            __threadfence();
            cons_bar();
            if (t == 0) {
                const unsigned long long prev = atomicAdd(p.done2 + tl.panel, 1ull);
                s_last = prev + 1ull == n2 * (unsigned long long)chunks;
            }
            cons_bar();
            if (s_last) {
                __threadfence();
                if (t < K * K)
                    p.hist_Gvp[((size_t)step * p.panels2 + tl.panel) * K * K + t] =
                        sum_strided_cg(p.part2 + (int64_t)tl.panel * chunks * K * K + t, chunks, K * K);
                if (t == 128) p.hist_VBp[(size_t)step * p.panels2 + tl.panel] = sum_strided_cg(p.vb2 + (int64_t)tl.panel * chunks, chunks, 1);
                __threadfence();
                cons_bar();
                if (t == 0) atomicAdd(p.vdone, 1ull);
            }
            cons_bar();
        }
    }
}

}  // namespace prmf
