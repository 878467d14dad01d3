// Host-side helper of the candidate pruning (`restrict`, script/prmf_runner.py:129-194): plain C++, no CUDA.
// The decision logic stays on the host, as in the reference; this only removes numpy's per-call overhead from
// a loop that runs once per factor per outer iteration while the GPU is waiting for the next active set
// (0.27 ms per outer iteration at config 2 -- 10 % of an outer iteration on 8 GPUs).  Every value is computed by
// the same IEEE operations, in the same order, as prmf_b200/solver.py:restrict_from_tables / percentile_19_9
// (np.sqrt, +, -, *, np.partition's order statistics, numpy's _lerp), so the results are bit-identical
// (tests/test_host_logic.py).  Built without FMA contraction.
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/prmf_b200.h"

extern "C" int64_t prmf_host_restrict(const double* mass_row, const double* quad_row, const int64_t* ids, int64_t n,
                                      double q, double* scores_out, int64_t* keep_out) {
    if (!mass_row || !quad_row || !ids || !scores_out || !keep_out || n <= 0) return -1;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t p = ids[i];
        const double sm = std::sqrt(mass_row[p]);
        const double one_minus = 1.0 - quad_row[p];
        scores_out[i] = sm + one_minus;                               // :123-125
    }
    // np.percentile(scores, 19.9), method 'linear': virtual index (n-1)*q, the two neighbouring order statistics
    const double vi = (double)(n - 1) * q;
    const int64_t lo = (int64_t)std::floor(vi);
    const int64_t hi = std::min<int64_t>(lo + 1, n - 1);
    const double g = vi - (double)lo;
    static thread_local std::vector<double> tmp;                      // scratch kept across calls (one per host thread)
    tmp.assign(scores_out, scores_out + n);
    std::nth_element(tmp.begin(), tmp.begin() + lo, tmp.end());
    const double a = tmp[lo];
    double b = a;
    if (hi != lo) b = *std::min_element(tmp.begin() + lo + 1, tmp.end());   // the (lo+1)-th order statistic
    const double diff = b - a;
    const double thr = (g >= 0.5) ? (b - diff * (1.0 - g)) : (a + diff * g);     // numpy's _lerp
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i)
        if (scores_out[i] > thr) keep_out[cnt++] = i;                 // strict >, :171
    return cnt;
}

// All factors of one `restrict` call at once (one FFI crossing per outer iteration): factor f has the candidate ids
// ids[off[f] .. off[f+1]) and reads row factor[f] of the P-column tables.  Outputs are compacted per factor:
// kept_ids / kept_scores hold the survivors of factor f at [kept_off[f], kept_off[f+1]).  Returns 0, or -(f+1) when
// factor f has no survivor (all scores equal -- the case in which the reference raises), or -1000000 on bad arguments.
extern "C" int64_t prmf_host_restrict_batch(const double* mass, const double* quad, int64_t P, int32_t nf,
                                            const int32_t* factor, const int64_t* ids, const int64_t* off, double q,
                                            int64_t* kept_ids, double* kept_scores, int64_t* kept_off) {
    if (!mass || !quad || !factor || !ids || !off || !kept_ids || !kept_scores || !kept_off || nf < 0 || P <= 0)
        return -1000000;
    std::vector<double> scores;
    std::vector<int64_t> keep;
    int64_t w = 0;
    kept_off[0] = 0;
    for (int32_t f = 0; f < nf; ++f) {
        const int64_t n = off[f + 1] - off[f];
        if (n <= 0) return -1000000;
        scores.resize((size_t)n);
        keep.resize((size_t)n);
        const int64_t* fid = ids + off[f];
        const int64_t cnt = prmf_host_restrict(mass + (int64_t)factor[f] * P, quad + (int64_t)factor[f] * P, fid, n, q,
                                               scores.data(), keep.data());
        if (cnt < 0) return -1000000;
        if (cnt == 0) return -(int64_t)(f + 1);
        for (int64_t i = 0; i < cnt; ++i) {
            kept_ids[w + i] = fid[keep[(size_t)i]];
            kept_scores[w + i] = scores[(size_t)keep[(size_t)i]];
        }
        w += cnt;
        kept_off[f + 1] = w;
    }
    return 0;
}
