// gapped_kernel.cu — stage 3: score-only gapped extension of every ungapped HSP.
//
// Replaces (semantics, not code):
//   greedy (megablast)  BLAST_GreedyGappedAlignment core/blast_gapalign.c:2620-2751
//                       BLAST_AffineGreedyAlign dispatch core/greedy_align.c:801-815
//                       BLAST_GreedyAlign :385-681, s_FindFirstMismatch :318-381
//   DP (blastn)         s_BlastDynProgNtGappedAlignment core/blast_gapalign.c:2763-2825
//                       s_BlastAlignPackedNucl :2843-3056
//   start points        BLAST_GetGappedScore :3466-3497 (middle of the ungapped HSP for greedy,
//                       seed + 3 for DP)
//
// Parallel formulation: an extension's result depends only on (query context,
// subject, start point, parameters), never on earlier extensions, so ALL init-HSPs are extended
// speculatively, one per thread, and the host afterwards replays the reference's sequential
// interval-tree containment filter over the sorted list, discarding results whose init-HSP the
// reference would have skipped.  Each thread keeps the exact serial recurrence (row-major
// best_score pruning for the DP, per-distance diagonal bounds for greedy), so scores, end points
// and seed estimates are bit-identical.
//
// Scratch is tiered: tier 1 gives every thread a small window (greedy: distances <= tier_d,
// DP: a ring of tier_d cells); an extension that outgrows it reports status 1 and is redone
// in tier 2 with worst-case scratch (greedy max_dist 10000 rows, DP ring >= query length).
#include <algorithm>
#include "bn_device.cuh"

namespace bn {

constexpr int GAP_THREADS = 64;
constexpr int GAP_BLOCKS = 592;          // 4 x 148 SMs
int gapped_threads() { return GAP_THREADS * GAP_BLOCKS; }
int gapped_dp_smem_blocks() { return 148 * 3; }     // 64 KB of rings per block: three blocks per SM
int gapped_dp_ring16_blocks() { return 148 * 10; }  // 16 KB of 16-bit rings per block: ten blocks (640 threads) per SM
int gapped_threads_per_block() { return GAP_THREADS; }

constexpr int32_t GREEDY_MAX_COST = 10000;
constexpr int32_t GREEDY_INVALID = -2;
constexpr int32_t MININT = INT32_MIN / 2;

// s_FindFirstMismatch (compressed seq2 branch) on 16-base windows.  seq1 = query (absolute
// concatenated position qbase + i), seq2 = subject (absolute volume base sbase + i).
struct SeqPair {
    const DevQuery *q;
    const uint8_t *packed;
    int32_t qbase;      // forward: position of seq1[0]; reverse: position of seq1[0] as well
    int64_t sbase;      // absolute base index of seq2[0]
    int32_t len1, len2;
    bool reverse;
};
__device__ __forceinline__ int32_t first_mismatch(const SeqPair &p, int32_t i1, int32_t i2)
{
    const int32_t n = min(p.len1 - i1, p.len2 - i2);
    if (n <= 0) return 0;
    if (p.reverse)
        return match_run_rev(*p.q, p.packed, p.qbase + p.len1 - i1, p.sbase + p.len2 - i2, n);
    return match_run_fwd(*p.q, p.packed, p.qbase + i1, p.sbase + i2, n);
}

struct GreedySeed { int32_t start_q, start_s, match_length; };

// BLAST_GreedyAlign, score only.  rows: 2 x (2*D + 6) ints; max_score: D + 1 + xdrop_offset ints.
// Diagonal k of the reference is stored at index k - diag_origin + D + 2.
__device__ int32_t greedy_align(const SeqPair &sp, int32_t xdrop_threshold, int32_t match_cost,
                                int32_t mismatch_cost, int32_t &seq1_len, int32_t &seq2_len,
                                int32_t *row0, int32_t *row1, int32_t *max_score_mem, int32_t D,
                                GreedySeed &seed, bool &overflow)
{
    const int32_t len1 = sp.len1, len2 = sp.len2;
    int32_t best_dist = 0;
    const int32_t max_dist = min(GREEDY_MAX_COST, len2 / 2 + 1);
    const int32_t origin = D + 2;                 // re-biased diag_origin
    const int32_t xdrop_offset = (xdrop_threshold + match_cost / 2) / (match_cost + mismatch_cost) + 1;

    int32_t index = first_mismatch(sp, 0, 0);
    seq1_len = index; seq2_len = index;
    int32_t seq1_index = index, seq2_index;
    seed.start_q = 0; seed.start_s = 0;
    int32_t longest_match_run = index;
    seed.match_length = index;
    if (index == len1 || index == len2) return 0;

    int32_t *max_score = max_score_mem + xdrop_offset;
    for (int32_t i = 0; i < xdrop_offset; i++) max_score_mem[i] = 0;
    row0[origin] = seq1_index;
    max_score[0] = seq1_index * match_cost;
    int32_t diag_lower = origin - 1, diag_upper = origin + 1;
    bool end1_reached = false, end2_reached = false;

    for (int32_t d = 1; d <= max_dist; d++) {
        if (d > D) { overflow = true; return best_dist; }
        int32_t curr_extent = 0, curr_seq2_index = 0, curr_diag = 0;
        const int32_t tmp_lower = diag_lower, tmp_upper = diag_upper;
        int32_t *prev = ((d - 1) & 1) ? row1 : row0;
        int32_t *cur = (d & 1) ? row1 : row0;
        prev[diag_lower - 1] = GREEDY_INVALID;
        prev[diag_lower] = GREEDY_INVALID;
        prev[diag_upper] = GREEDY_INVALID;
        prev[diag_upper + 1] = GREEDY_INVALID;

        int32_t xdrop_score = max_score[d - xdrop_offset] + (match_cost + mismatch_cost) * d - xdrop_threshold;
        // (Int4)ceil((double)x / (match_cost/2)); match_cost/2 >= 1, exact integer ceiling
        {
            const int32_t h = match_cost / 2;
            int32_t qd = xdrop_score / h, r = xdrop_score % h;
            if (r > 0) ++qd;
            xdrop_score = qd;
        }
        for (int32_t k = tmp_lower; k <= tmp_upper; k++) {
            seq2_index = max(prev[k + 1], prev[k]) + 1;
            seq2_index = max(seq2_index, prev[k - 1]);
            seq1_index = seq2_index + k - origin;
            if (seq2_index < 0 || seq1_index + seq2_index < xdrop_score) {
                if (k == diag_lower) diag_lower++;
                else cur[k] = GREEDY_INVALID;
                continue;
            }
            diag_upper = k;
            index = first_mismatch(sp, seq1_index, seq2_index);
            if (index > longest_match_run) {
                seed.start_q = seq1_index; seed.start_s = seq2_index;
                seed.match_length = longest_match_run = index;
            }
            seq1_index += index; seq2_index += index;
            cur[k] = seq2_index;
            if (seq1_index + seq2_index > curr_extent) {
                curr_extent = seq1_index + seq2_index;
                curr_seq2_index = seq2_index;
                curr_diag = k;
            }
            if (seq2_index == len2) { diag_lower = k + 1; end2_reached = true; }
            if (seq1_index == len1) { diag_upper = k - 1; end1_reached = true; }
        }
        const int32_t curr_score = curr_extent * (match_cost / 2) - d * (match_cost + mismatch_cost);
        if (curr_score > max_score[d - 1]) {
            max_score[d] = curr_score;
            best_dist = d;
            seq2_len = curr_seq2_index;
            seq1_len = curr_seq2_index + curr_diag - origin;
        } else max_score[d] = max_score[d - 1];
        if (diag_lower > diag_upper) break;
        if (!end2_reached) diag_lower--;
        if (!end1_reached) diag_upper++;
    }
    return best_dist;
}

// BLAST_AffineGreedyAlign, affine body (core/greedy_align.c:817-1237), score only.  One thread, exact
// serial recurrence.  Scratch (ints): (max_penalty + 1) rows x (2D + 6) diagonals x {insert, match,
// delete}, diagonal bounds for D * gap_extend + 1 (+ max_penalty) distances, per-distance best scores.
// Row of distance d = slot d % (max_penalty + 1): the reference recycles rows the same way when no
// traceback is kept (:1165-1172).  Diagonal k of the reference is stored at k - diag_origin + D + 2; a
// gap costs at least gap_extend, so |k - diag_origin| <= d / gap_extend + 1 <= D + 1 while d <= D * gap_extend.
__device__ int32_t greedy_align_affine(const SeqPair &sp, const AffineCosts &ac, int32_t &seq1_len, int32_t &seq2_len,
                                       int32_t *scratch, int32_t D, GreedySeed &seed, bool &overflow)
{
    const int32_t kInvalidDiag = 100000000;
    const int32_t len1 = sp.len1, len2 = sp.len2;
    const int32_t match_half = ac.match / 2;
    const int32_t op_cost = ac.op_cost, gap_extend = ac.gap_extend, goe = ac.gap_open + ac.gap_extend;
    const int32_t max_penalty = ac.max_penalty, nrows = ac.max_penalty + 1;
    const int32_t max_dist = min(GREEDY_MAX_COST, len2 / 2 + 1);
    const int32_t scaled_max_dist = max_dist * gap_extend;
    const int32_t d_cap = D * gap_extend;
    const int32_t origin = D + 2, width = 2 * D + 6;

    int32_t index = first_mismatch(sp, 0, 0);
    seq1_len = index; seq2_len = index;
    int32_t seq1_index = index, seq2_index;
    seed.start_q = 0; seed.start_s = 0;
    int32_t longest_match_run = index;
    seed.match_length = index;
    if (index == len1 || index == len2) return index * ac.match;

    int32_t *rows = scratch;                                            // [nrows][width][3]
    int32_t *diag_lower = rows + 3 * (int64_t)nrows * width + max_penalty;
    int32_t *diag_upper = diag_lower + d_cap + 1 + max_penalty;
    int32_t *max_score_mem = diag_upper + d_cap + 1;
    int32_t *max_score = max_score_mem + ac.xdrop_offset;
#define AFF(dd, kk, f) rows[(((int64_t)((dd) % nrows)) * width + (kk)) * 3 + (f)]       // f: 0 insert, 1 match, 2 delete
    for (int32_t i = 0; i < ac.xdrop_offset; i++) max_score_mem[i] = 0;
    for (int32_t i = 1; i <= max_penalty; i++) { diag_lower[-i] = kInvalidDiag; diag_upper[-i] = -kInvalidDiag; }
    AFF(0, origin, 1) = seq1_index;
    AFF(0, origin, 0) = GREEDY_INVALID;
    AFF(0, origin, 2) = GREEDY_INVALID;
    max_score[0] = seq1_index * ac.match;
    diag_lower[0] = origin; diag_upper[0] = origin;
    int32_t curr_lower = origin - 1, curr_upper = origin + 1;
    int32_t end1_diag = 0, end2_diag = 0, num_nonempty = 1;
    int32_t best_dist = 0;
    int32_t d = 1;

    while (d <= scaled_max_dist) {
        if (d > d_cap) { overflow = true; return 0; }
        int32_t curr_extent = 0, curr_seq2_index = 0, curr_diag = 0;
        const int32_t tmp_lower = curr_lower, tmp_upper = curr_upper;
        int32_t xdrop_score = max_score[d - ac.xdrop_offset] + ac.common_factor * d - ac.xdrop;
        {   // (Int4)ceil((double)x / match_half), match_half >= 1
            int32_t qd = xdrop_score / match_half;
            if (xdrop_score % match_half > 0) ++qd;
            xdrop_score = qd < 0 ? 0 : qd;
        }
        const int32_t lo_goe = diag_lower[d - goe], up_goe = diag_upper[d - goe];
        const int32_t lo_ge = diag_lower[d - gap_extend], up_ge = diag_upper[d - gap_extend];
        const int32_t lo_op = diag_lower[d - op_cost], up_op = diag_upper[d - op_cost];
        for (int32_t k = tmp_lower; k <= tmp_upper; k++) {
            seq2_index = GREEDY_INVALID;
            if (k + 1 <= up_goe && k + 1 >= lo_goe) seq2_index = AFF(d - goe, k + 1, 1);
            if (k + 1 <= up_ge && k + 1 >= lo_ge) {
                const int32_t v = AFF(d - gap_extend, k + 1, 2);
                if (seq2_index < v) seq2_index = v;
            }
            const int32_t del = (seq2_index == GREEDY_INVALID) ? GREEDY_INVALID : seq2_index + 1;
            AFF(d, k, 2) = del;

            seq2_index = GREEDY_INVALID;
            if (k - 1 <= up_goe && k - 1 >= lo_goe) seq2_index = AFF(d - goe, k - 1, 1);
            if (k - 1 <= up_ge && k - 1 >= lo_ge) {
                const int32_t v = AFF(d - gap_extend, k - 1, 0);
                if (seq2_index < v) seq2_index = v;
            }
            AFF(d, k, 0) = seq2_index;

            seq2_index = max(seq2_index, del);
            if (k <= up_op && k >= lo_op) seq2_index = max(seq2_index, AFF(d - op_cost, k, 1) + 1);
            seq1_index = seq2_index + k - origin;
            if (seq2_index < 0 || seq1_index + seq2_index < xdrop_score) {
                if (k == curr_lower) curr_lower++;
                else AFF(d, k, 1) = GREEDY_INVALID;
                continue;
            }
            curr_upper = k;
            index = first_mismatch(sp, seq1_index, seq2_index);
            if (index > longest_match_run) {
                seed.start_q = seq1_index; seed.start_s = seq2_index;
                seed.match_length = longest_match_run = index;
            }
            seq1_index += index; seq2_index += index;
            AFF(d, k, 1) = seq2_index;
            if (seq1_index + seq2_index > curr_extent) {
                curr_extent = seq1_index + seq2_index;
                curr_seq2_index = seq2_index;
                curr_diag = k;
            }
            if (seq1_index == len1) { curr_upper = k; end1_diag = k - 1; }
            if (seq2_index == len2) { curr_lower = k; end2_diag = k + 1; }
        }
        const int32_t curr_score = curr_extent * match_half - d * ac.common_factor;
        if (curr_score > max_score[d - 1]) {
            max_score[d] = curr_score;
            best_dist = d;
            seq2_len = curr_seq2_index;
            seq1_len = curr_seq2_index + curr_diag - origin;
        } else max_score[d] = max_score[d - 1];
        if (curr_lower <= curr_upper) {
            num_nonempty++;
            diag_lower[d] = curr_lower; diag_upper[d] = curr_upper;
        } else { diag_lower[d] = kInvalidDiag; diag_upper[d] = -kInvalidDiag; }
        if (diag_lower[d - max_penalty] <= diag_upper[d - max_penalty]) num_nonempty--;
        if (num_nonempty == 0) break;

        d++;
        if (d > d_cap) { if (d <= scaled_max_dist) { overflow = true; return 0; } break; }
        curr_lower = min(diag_lower[d - goe], diag_lower[d - gap_extend]) - 1;
        curr_lower = min(curr_lower, diag_lower[d - op_cost]);
        if (end2_diag > 0) curr_lower = max(curr_lower, end2_diag);
        curr_upper = max(diag_upper[d - goe], diag_upper[d - gap_extend]) + 1;
        curr_upper = max(curr_upper, diag_upper[d - op_cost]);
        if (end1_diag > 0) curr_upper = min(curr_upper, end1_diag);
    }
#undef AFF
    return max_score[best_dist];
}

__device__ void greedy_gapped(const DevQuery &q, const uint8_t *packed, int32_t ctx_off, int32_t qlen,
                              int64_t chunk_base, int32_t slen, int32_t q_off, int32_t s_off,
                              int32_t *scratch, int32_t D, DevGapResult &g)
{
    int32_t match = q.reward, mismatch = -q.penalty, xd = q.gap_x_dropoff;
    if (match % 2 == 1) { match *= 2; mismatch *= 2; xd *= 2; }
    int32_t *row0 = scratch, *row1 = scratch + (2 * D + 6), *ms = scratch + 2 * (2 * D + 6);
    int32_t q_ext_r, s_ext_r, q_ext_l, s_ext_l;
    GreedySeed fwd, rev;
    bool overflow = false;
    const bool affine = q.gap_open != 0 || q.gap_extend != 0;
    const AffineCosts ac = affine_costs(q.reward, q.penalty, q.gap_open, q.gap_extend, q.gap_x_dropoff);
    SeqPair sp;
    sp.q = &q; sp.packed = packed;
    sp.qbase = ctx_off + q_off; sp.sbase = chunk_base + s_off;
    sp.len1 = qlen - q_off; sp.len2 = slen - s_off; sp.reverse = false;
    int32_t score = affine ? greedy_align_affine(sp, ac, q_ext_r, s_ext_r, scratch, D, fwd, overflow)
                           : greedy_align(sp, xd, match, mismatch, q_ext_r, s_ext_r, row0, row1, ms, D, fwd, overflow);
    if (!overflow) {
        sp.qbase = ctx_off; sp.sbase = chunk_base; sp.len1 = q_off; sp.len2 = s_off; sp.reverse = true;
        score += affine ? greedy_align_affine(sp, ac, q_ext_l, s_ext_l, scratch, D, rev, overflow)
                        : greedy_align(sp, xd, match, mismatch, q_ext_l, s_ext_l, row0, row1, ms, D, rev, overflow);
    }
    if (overflow) { g.status = 1; return; }
    // the basic algorithm returns distances, the affine one scores in (possibly doubled) units (:2683-2690)
    if (!affine) score = (q_ext_r + s_ext_r + q_ext_l + s_ext_l) * q.reward / 2 - score * (q.reward - q.penalty);
    else if (q.reward % 2 == 1) score /= 2;

    const int32_t q_box_l = q_off - q_ext_l, s_box_l = s_off - s_ext_l;
    const int32_t q_box_r = q_off + q_ext_r, s_box_r = s_off + s_ext_r;
    int32_t q_seed_l = q_off - rev.start_q, s_seed_l = s_off - rev.start_s;
    int32_t q_seed_r = q_off + fwd.start_q, s_seed_r = s_off + fwd.start_s;
    int32_t vl = 0, vr = 0;
    if (q_seed_r < q_box_r && s_seed_r < s_box_r) {
        vr = min(q_box_r - q_seed_r, s_box_r - s_seed_r);
        vr = min(vr, fwd.match_length) / 2;
    } else { q_seed_r = q_off; s_seed_r = s_off; }
    if (q_seed_l > q_box_l && s_seed_l > s_box_l) {
        vl = min(q_seed_l - q_box_l, s_seed_l - s_box_l);
        vl = min(vl, rev.match_length) / 2;
    } else { q_seed_l = q_off; s_seed_l = s_off; }
    if (vr > vl) { g.q_seed = q_seed_r + vr; g.s_seed = s_seed_r + vr; }
    else { g.q_seed = q_seed_l - vl; g.s_seed = s_seed_l - vl; }
    g.q_start = q_box_l; g.s_start = s_box_l; g.q_stop = q_box_r; g.s_stop = s_box_r;
    g.score = score;
    g.status = 0;
}

// ================================================================================================
// Warp-parallel greedy: one warp per extension, one lane per diagonal of the current distance.
// Within a distance d the reference visits diagonals k in ascending order; each diagonal's new
// offset depends only on row d-1, so all lanes compute theirs at once (each walks its own run of
// matches on 16-base windows) and the ORDER-DEPENDENT bookkeeping — diag_lower++ while the lowest
// diagonals fail, diag_upper = last success, the end-of-sequence clamps, "first maximum wins" for
// the extent and for the longest match run — is then replayed from ballots in ascending k.
// ================================================================================================
constexpr unsigned FULLW = 0xffffffffu;

// exact-match run measured by the whole warp (used for the long initial run of each direction)
__device__ int32_t first_mismatch_warp(const SeqPair &p, int32_t i1, int32_t i2, int lane)
{
    const int32_t n = min(p.len1 - i1, p.len2 - i2);
    if (n <= 0) return 0;
    for (int32_t base = 0; base < n; base += 512) {
        const int32_t off = base + 16 * lane;
        int32_t c = 16;
        if (off < n) {
            uint32_t qb, qa, m;
            if (p.reverse) {
                qwin(*p.q, p.qbase + p.len1 - i1 - off - 16, qb, qa);
                m = mismatch_bits(qb, qa, swin(p.packed, p.sbase + p.len2 - i2 - off - 16));
                if (m) c = (__ffs(m) - 1) >> 1;
            } else {
                qwin(*p.q, p.qbase + i1 + off, qb, qa);
                m = mismatch_bits(qb, qa, swin(p.packed, p.sbase + i2 + off));
                if (m) c = __clz(m) >> 1;
            }
        }
        const unsigned stop = __ballot_sync(FULLW, off < n && c < 16);
        if (stop) {
            const int src = __ffs(stop) - 1;
            const int32_t cc = __shfl_sync(FULLW, c, src);
            return min(base + 16 * src + cc, n);
        }
    }
    return n;
}

__device__ int32_t greedy_align_warp(const SeqPair &sp, int32_t xdrop_threshold, int32_t match_cost,
                                     int32_t mismatch_cost, int32_t &seq1_len, int32_t &seq2_len,
                                     int32_t *row0, int32_t *row1, int32_t *max_score_mem, int32_t D,
                                     GreedySeed &seed, bool &overflow, int lane)
{
    const int32_t len1 = sp.len1, len2 = sp.len2;
    int32_t best_dist = 0;
    const int32_t max_dist = min(GREEDY_MAX_COST, len2 / 2 + 1);
    const int32_t origin = D + 2;
    const int32_t xdrop_offset = (xdrop_threshold + match_cost / 2) / (match_cost + mismatch_cost) + 1;

    int32_t index = first_mismatch_warp(sp, 0, 0, lane);
    seq1_len = index; seq2_len = index;
    seed.start_q = 0; seed.start_s = 0;
    int32_t longest_match_run = index;
    seed.match_length = index;
    if (index == len1 || index == len2) return 0;

    int32_t *max_score = max_score_mem + xdrop_offset;
    for (int32_t i = lane; i < xdrop_offset; i += 32) max_score_mem[i] = 0;
    if (lane == 0) { row0[origin] = index; max_score[0] = index * match_cost; }
    __syncwarp();
    int32_t diag_lower = origin - 1, diag_upper = origin + 1;
    bool end1_reached = false, end2_reached = false;

    for (int32_t d = 1; d <= max_dist; d++) {
        if (d > D) { overflow = true; return best_dist; }
        int32_t curr_extent = 0, curr_seq2_index = 0, curr_diag = 0;
        const int32_t tmp_lower = diag_lower, tmp_upper = diag_upper;
        int32_t *prev = ((d - 1) & 1) ? row1 : row0;
        int32_t *cur = (d & 1) ? row1 : row0;
        if (lane == 0) {
            prev[diag_lower - 1] = GREEDY_INVALID;
            prev[diag_lower] = GREEDY_INVALID;
            prev[diag_upper] = GREEDY_INVALID;
            prev[diag_upper + 1] = GREEDY_INVALID;
        }
        __syncwarp();
        int32_t xdrop_score = max_score[d - xdrop_offset] + (match_cost + mismatch_cost) * d - xdrop_threshold;
        {
            const int32_t h = match_cost / 2;
            int32_t qd = xdrop_score / h;
            if (xdrop_score % h > 0) ++qd;
            xdrop_score = qd;
        }
        for (int32_t kb = tmp_lower; kb <= tmp_upper; kb += 32) {
            const int32_t k = kb + lane;
            const bool active = k <= tmp_upper;
            bool ok = false;
            int32_t seq1_index = 0, seq2_index = 0, run = 0;
            if (active) {
                seq2_index = max(prev[k + 1], prev[k]) + 1;
                seq2_index = max(seq2_index, prev[k - 1]);
                seq1_index = seq2_index + k - origin;
                ok = !(seq2_index < 0 || seq1_index + seq2_index < xdrop_score);
                if (ok) {
                    run = first_mismatch(sp, seq1_index, seq2_index);
                    seq1_index += run; seq2_index += run;
                }
            }
            const unsigned act = __ballot_sync(FULLW, active);
            const unsigned succ = __ballot_sync(FULLW, ok);
            const unsigned e2 = __ballot_sync(FULLW, ok && seq2_index == len2);
            const unsigned e1 = __ballot_sync(FULLW, ok && seq1_index == len1);
            // ordered replay of the bookkeeping
            unsigned inv = 0;
            if (e2 == 0) {
                // no diagonal reached the end of seq2 in this round: diag_lower only moves over the
                // leading run of failures (and only if it still sits on the round's first diagonal),
                // every later failure is marked invalid, diag_upper follows the last success
                const unsigned fail = act & ~succ;
                unsigned lead = 0;
                if (diag_lower == kb) {
                    lead = (unsigned)__ffs(~fail) - 1u;          // fail has no bits above act, so ~fail != 0 ... unless act is full
                    if (fail == 0xffffffffu) lead = 32;
                    diag_lower += (int32_t)lead;
                }
                inv = lead >= 32 ? 0u : (fail & ~((1u << lead) - 1u));
                if (succ) {
                    const int top = 31 - __clz(succ);
                    diag_upper = kb + top;
                    if (e1) {
                        end1_reached = true;
                        if ((e1 >> top) & 1u) diag_upper = kb + top - 1;
                    }
                }
            } else {
                for (unsigned rem = act; rem; rem &= rem - 1) {
                    const int b = __ffs(rem) - 1;
                    const unsigned bit = 1u << b;
                    const int32_t kk = kb + b;
                    if (!(succ & bit)) {
                        if (kk == diag_lower) diag_lower++;
                        else inv |= bit;
                    } else {
                        diag_upper = kk;
                        if (e2 & bit) { diag_lower = kk + 1; end2_reached = true; }
                        if (e1 & bit) { diag_upper = kk - 1; end1_reached = true; }
                    }
                }
            }
            if (ok) cur[k] = seq2_index;
            else if (inv & (1u << lane)) cur[k] = GREEDY_INVALID;
            if (succ) {
                // longest run of matches: first diagonal (ascending k) holding the strict maximum
                const int32_t best_run = __reduce_max_sync(FULLW, ok ? run : -1);          // REDUX.MAX
                if (best_run > longest_match_run) {
                    const int src = __ffs(__ballot_sync(FULLW, ok && run == best_run)) - 1;
                    seed.start_q = __shfl_sync(FULLW, seq1_index - run, src);
                    seed.start_s = __shfl_sync(FULLW, seq2_index - run, src);
                    seed.match_length = longest_match_run = best_run;
                }
                // extent: first diagonal with the strict maximum of seq1 + seq2
                int32_t ext = ok ? seq1_index + seq2_index : -1;
                const int32_t best_ext = __reduce_max_sync(FULLW, ext);
                if (best_ext > curr_extent) {
                    const int src = __ffs(__ballot_sync(FULLW, ok && ext == best_ext)) - 1;
                    curr_extent = best_ext;
                    curr_seq2_index = __shfl_sync(FULLW, seq2_index, src);
                    curr_diag = kb + src;
                }
            }
        }
        const int32_t curr_score = curr_extent * (match_cost / 2) - d * (match_cost + mismatch_cost);
        const int32_t prev_best = max_score[d - 1];
        __syncwarp();
        if (curr_score > prev_best) {
            if (lane == 0) max_score[d] = curr_score;
            best_dist = d;
            seq2_len = curr_seq2_index;
            seq1_len = curr_seq2_index + curr_diag - origin;
        } else if (lane == 0) max_score[d] = prev_best;
        __syncwarp();
        if (diag_lower > diag_upper) break;
        if (!end2_reached) diag_lower--;
        if (!end1_reached) diag_upper++;
    }
    return best_dist;
}

// TWO warps per init-HSP — the forward (right) and the reverse (left) extension are independent, so
// they run side by side and meet at a named barrier; rows live in shared memory (tier 1) or in global
// scratch (tier 2).  Block = 4 warps = 2 init-HSPs in flight.
__global__ void __launch_bounds__(128)
greedy_kernel(const DevQuery q, const GappedLaunch L, int use_smem)
{
    extern __shared__ int32_t smem_rows[];
    __shared__ int32_t xch[2][8];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int pair = wib >> 1, dir = wib & 1;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const int64_t pair0 = (int64_t)blockIdx.x * 2 + pair;
    const int64_t npairs = (int64_t)gridDim.x * 2;
    const int64_t n = L.todo ? (int64_t)L.n_todo : (int64_t)min((unsigned long long)L.max_init, *L.n_init);
    const int32_t D = L.tier_d;
    int32_t *scratch = use_smem ? smem_rows + (size_t)wib * L.scratch_ints_per_thread
                                : L.scratch + warp * L.scratch_ints_per_thread;
    int32_t *row0 = scratch, *row1 = scratch + (2 * D + 6), *ms = scratch + 2 * (2 * D + 6);
    int32_t match = q.reward, mismatch = -q.penalty, xd = q.gap_x_dropoff;
    if (match % 2 == 1) { match *= 2; mismatch *= 2; xd *= 2; }

    for (int64_t w = pair0; w < n; w += npairs) {
        const int64_t i = L.todo ? (int64_t)L.todo[w] : w;
        const DevInitHit h = L.init[i];
        const DevChunk ch = L.chunks[h.chunk];
        const int32_t context = ctx_search_warp(q, h.q_off, lane);      // 32 pivots per round: 2-3 rounds, not 11-17
        const DevContext c = q.ctx[context];
        const int32_t q_off = (h.q_start - c.query_offset) + h.length / 2;
        const int32_t s_off = h.s_start + h.length / 2;
        const int64_t chunk_base = ch.byte_off * 4;
        int32_t q_ext = 0, s_ext = 0;
        GreedySeed seed{0, 0, 0};
        bool overflow = false;
        SeqPair sp;
        sp.q = &q; sp.packed = L.packed;
        if (dir == 0) {
            sp.qbase = c.query_offset + q_off; sp.sbase = chunk_base + s_off;
            sp.len1 = c.query_length - q_off; sp.len2 = ch.len - s_off; sp.reverse = false;
        } else {
            sp.qbase = c.query_offset; sp.sbase = chunk_base; sp.len1 = q_off; sp.len2 = s_off; sp.reverse = true;
        }
        const int32_t dist_mine = greedy_align_warp(sp, xd, match, mismatch, q_ext, s_ext, row0, row1, ms, D, seed, overflow, lane);
        __syncwarp();
        if (dir == 1 && lane == 0) {
            xch[pair][0] = dist_mine; xch[pair][1] = q_ext; xch[pair][2] = s_ext;
            xch[pair][3] = seed.start_q; xch[pair][4] = seed.start_s; xch[pair][5] = seed.match_length;
            xch[pair][6] = overflow ? 1 : 0;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
        if (dir == 0) {
            DevGapResult g;
            g.q_start = g.q_stop = g.s_start = g.s_stop = g.score = g.q_seed = g.s_seed = 0;
            g.status = 0;
            const int32_t q_ext_r = q_ext, s_ext_r = s_ext;
            const GreedySeed fwd = seed;
            const int32_t dist = dist_mine + xch[pair][0];
            const int32_t q_ext_l = xch[pair][1], s_ext_l = xch[pair][2];
            const GreedySeed rev{xch[pair][3], xch[pair][4], xch[pair][5]};
            if (overflow || xch[pair][6]) g.status = 1;
            else {
                const int32_t score = (q_ext_r + s_ext_r + q_ext_l + s_ext_l) * q.reward / 2 - dist * (q.reward - q.penalty);
                const int32_t q_box_l = q_off - q_ext_l, s_box_l = s_off - s_ext_l;
                const int32_t q_box_r = q_off + q_ext_r, s_box_r = s_off + s_ext_r;
                int32_t q_seed_l = q_off - rev.start_q, s_seed_l = s_off - rev.start_s;
                int32_t q_seed_r = q_off + fwd.start_q, s_seed_r = s_off + fwd.start_s;
                int32_t vl = 0, vr = 0;
                if (q_seed_r < q_box_r && s_seed_r < s_box_r) {
                    vr = min(q_box_r - q_seed_r, s_box_r - s_seed_r);
                    vr = min(vr, fwd.match_length) / 2;
                } else { q_seed_r = q_off; s_seed_r = s_off; }
                if (q_seed_l > q_box_l && s_seed_l > s_box_l) {
                    vl = min(q_seed_l - q_box_l, s_seed_l - s_box_l);
                    vl = min(vl, rev.match_length) / 2;
                } else { q_seed_l = q_off; s_seed_l = s_off; }
                if (vr > vl) { g.q_seed = q_seed_r + vr; g.s_seed = s_seed_r + vr; }
                else { g.q_seed = q_seed_l - vl; g.s_seed = s_seed_l - vl; }
                g.q_start = q_box_l; g.s_start = s_box_l; g.q_stop = q_box_r; g.s_stop = s_box_r;
                g.score = score;
            }
            if (lane == 0) L.out[i] = g;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");     // xch is reused by the next init-HSP
    }
}

cudaError_t launch_greedy_warp(const DevQuery &q, const GappedLaunch &g, int warps_per_block, int blocks,
                               bool use_smem, cudaStream_t st)
{
    const size_t smem = use_smem ? (size_t)warps_per_block * (size_t)g.scratch_ints_per_thread * sizeof(int32_t) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    greedy_kernel<<<blocks, warps_per_block * 32, smem, st>>>(q, g, use_smem ? 1 : 0);
    return cudaGetLastError();
}

// s_BlastAlignPackedNucl with the score_array kept in a ring of C cells (C a power of two): only
// indices in [first_b_index, b_size] are live, so index i lives at slot i & (C - 1) as long as the live
// span <= C.  Two ring homes:
//   SmemRing   tier 1: DP_SMEM_CELLS cells per thread in shared memory, interleaved by thread
//              (cell c of thread t at [c][t]) so a warp's 8-byte accesses never conflict, whatever
//              cells its lanes are at.  A dependent global-memory round trip per DP cell made the
//              longest alignment of a batch (one thread, ~4e5 cells for a 10 kb query) set the kernel time.
//   GlobalRing tier 2: worst-case capacity in global scratch for the rare band wider than tier 1.
constexpr int DP_SMEM_CELLS = 128;
struct SmemRing {
    int2 *base;                 // &ring[0][thread]
    __device__ __forceinline__ int2 get(int32_t i) const { return base[(i & (DP_SMEM_CELLS - 1)) * GAP_THREADS]; }
    __device__ __forceinline__ void set(int32_t i, int2 v) const { base[(i & (DP_SMEM_CELLS - 1)) * GAP_THREADS] = v; }
    __device__ __forceinline__ int32_t capacity() const { return DP_SMEM_CELLS; }
};
// The same ring with 16-bit cells: a tier-1 extension walks at most dp_max_rows rows, so every live score lies in
// (-(X + 2 (gap_open + gap_extend)), reward * rows]; the launcher picks this ring only when that interval fits 16 bits.
// The reference's MININT is the one value outside it and travels as -32768.  (A stored score is either a live cell's
// or exactly MININT: pruned cells are written as MININT and keep their old gap score, so no "MININT + small" value
// is ever stored.)  Half the bytes per thread: 2.5 x the resident threads of the 32-bit ring, which is what this
// latency-bound loop needs.
constexpr int DP16_CELLS = 64;
struct SmemRing16 {
    short2 *base;               // &ring[0][thread]
    __device__ __forceinline__ int2 get(int32_t i) const
    {
        const short2 v = base[(i & (DP16_CELLS - 1)) * GAP_THREADS];
        return make_int2(v.x == -32768 ? MININT : (int32_t)v.x, v.y == -32768 ? MININT : (int32_t)v.y);
    }
    __device__ __forceinline__ void set(int32_t i, int2 v) const
    {
        base[(i & (DP16_CELLS - 1)) * GAP_THREADS] =
            make_short2(v.x < -32000 ? (short)-32768 : (short)v.x, v.y < -32000 ? (short)-32768 : (short)v.y);
    }
    __device__ __forceinline__ int32_t capacity() const { return DP16_CELLS; }
};
struct GlobalRing {
    int2 *base;
    int32_t mask;               // capacity - 1
    __device__ __forceinline__ int2 get(int32_t i) const { return base[i & mask]; }
    __device__ __forceinline__ void set(int32_t i, int2 v) const { base[i & mask] = v; }
    __device__ __forceinline__ int32_t capacity() const { return mask + 1; }
};

template <typename Ring>
__device__ int32_t dp_packed(const uint8_t *B, const uint8_t *A, int32_t N, int32_t M,
                             int32_t &b_offset, int32_t &a_offset, const int32_t *matrix,
                             int32_t gap_open, int32_t gap_extend, int32_t x_dropoff, bool reverse,
                             const Ring ring, bool &overflow, int32_t max_rows, bool &too_long)
{
    const int32_t gap_open_extend = gap_open + gap_extend;
    const int32_t C = ring.capacity();
    a_offset = 0; b_offset = 0;
    if (x_dropoff < gap_open_extend) x_dropoff = gap_open_extend;
    if (N <= 0 || M <= 0) return 0;

    int32_t score = -gap_open_extend;
    ring.set(0, make_int2(0, -gap_open_extend));
    int32_t i;
    for (i = 1; i <= N; i++) {
        if (score < -x_dropoff) break;
        if (i >= C) { overflow = true; return 0; }
        ring.set(i, make_int2(score, score - gap_open_extend));
        score -= gap_extend;
    }
    int32_t b_size = i, best_score = 0, first_b_index = 0;
    const int32_t b_inc = reverse ? -1 : 1;

    for (int32_t a_index = 1; a_index <= M; a_index++) {
        if (a_index > max_rows) { too_long = true; return 0; }      // a long alignment: the warp-parallel kernel takes it
        int a_bp;
        if (reverse) a_bp = (__ldg(A + (M - a_index) / 4) >> (2 * ((a_index - 1) % 4))) & 3;
        else a_bp = (__ldg(A + 1 + (a_index - 1) / 4) >> (2 * (3 - (a_index - 1) % 4))) & 3;
        const int32_t *mrow = matrix + 16 * a_bp;
        const uint8_t *b_ptr = reverse ? (B + N - first_b_index) : (B + first_b_index);
        score = MININT;
        int32_t score_gap_row = MININT, last_b_index = first_b_index;

        // the next cell and the next query byte are fetched one step ahead of their use
        int2 cell = ring.get(first_b_index);
        int qb = (first_b_index < b_size) ? (int)__ldg(b_ptr + b_inc) : 0;
        for (int32_t b_index = first_b_index; b_index < b_size; b_index++) {
            b_ptr += b_inc;
            const bool has_next = b_index + 1 < b_size;
            const int2 next_cell = has_next ? ring.get(b_index + 1) : cell;
            const int next_qb = has_next ? (int)__ldg(b_ptr + b_inc) : 0;
            int32_t score_gap_col = cell.y;
            const int32_t next_score = cell.x + mrow[qb];
            if (score < score_gap_col) score = score_gap_col;
            if (score < score_gap_row) score = score_gap_row;
            if (best_score - score > x_dropoff) {
                if (b_index == first_b_index) first_b_index++;
                else { cell.x = MININT; ring.set(b_index, cell); }
            } else {
                last_b_index = b_index;
                if (score > best_score) { best_score = score; a_offset = a_index; b_offset = b_index; }
                score_gap_row -= gap_extend;
                score_gap_col -= gap_extend;
                cell.y = max(score - gap_open_extend, score_gap_col);
                score_gap_row = max(score - gap_open_extend, score_gap_row);
                cell.x = score;
                ring.set(b_index, cell);
            }
            score = next_score;
            cell = next_cell;
            qb = next_qb;
        }
        if (first_b_index == b_size) break;
        if (last_b_index < b_size - 1) b_size = last_b_index + 1;
        else {
            while (score_gap_row >= (best_score - x_dropoff) && b_size <= N) {
                if (b_size - first_b_index + 2 >= C) { overflow = true; return 0; }
                ring.set(b_size, make_int2(score_gap_row, score_gap_row - gap_open_extend));
                score_gap_row -= gap_extend;
                b_size++;
            }
        }
        if (b_size <= N) {
            if (b_size - first_b_index + 2 >= C) { overflow = true; return 0; }
            ring.set(b_size, make_int2(MININT, MININT));
            b_size++;
        }
    }
    return best_score;
}

template <typename Ring>
__device__ void dp_gapped(const DevQuery &q, const int32_t *matrix, const uint8_t *query, int32_t qlen, const uint8_t *S,
                          int32_t slen, int32_t q_off, int32_t s_off, const Ring ring, int32_t max_rows, DevGapResult &g)
{
    const int32_t adj = 4 - (s_off % 4);
    int32_t q_length = q_off + adj, s_length = s_off + adj;
    if (q_length > qlen || s_length > slen) { q_length -= 4; s_length -= 4; }
    bool overflow = false, too_long = false;
    int32_t pq, ps, right = 0;
    const int32_t left = dp_packed(query, S, q_length, s_length, pq, ps, matrix, q.gap_open,
                                   q.gap_extend, q.gap_x_dropoff, true, ring, overflow, max_rows, too_long);
    if (too_long) { g.status = 2; return; }
    if (overflow) { g.status = 1; return; }
    g.q_start = q_length - pq; g.s_start = s_length - ps;
    if (q_length < qlen && s_length < slen) {
        int32_t qs, ss;
        right = dp_packed(query + q_length - 1, S + (s_length + 3) / 4 - 1, qlen - q_length,
                          slen - s_length, qs, ss, matrix, q.gap_open, q.gap_extend,
                          q.gap_x_dropoff, false, ring, overflow, max_rows, too_long);
        if (too_long) { g.status = 2; return; }
        if (overflow) { g.status = 1; return; }
        g.q_stop = qs + q_length; g.s_stop = ss + s_length;
    } else { g.q_stop = q_length; g.s_stop = s_length; }
    g.score = left + right;
    g.q_seed = q_off; g.s_seed = s_off;
    g.status = 0;
}

// ================================================================================================
// Warp-parallel packed DP for LONG alignments (the thread-per-HSP kernel hands over whatever exceeds
// its row budget): one warp per extension, one lane per band cell, exact.
//
// A row of s_BlastAlignPackedNucl visits the live cells b = first_b .. b_size-1 in order.  With
//   v_b = max(best_old[b-1] + matrix[a][B_b], best_gap_old[b])            (no dependency inside the row)
// the row is the recurrence
//   s_b = max(v_b, r_b);  pruned_b = (best_b - s_b > X)
//   unpruned: best_{b+1} = max(best_b, s_b),  r_{b+1} = max(s_b - goe, r_b - ge) = max(v_b - goe, r_b - ge)
//   pruned  : best_{b+1} = best_b,            r_{b+1} = r_b           (a pruned cell neither pays nor feeds the gap)
// For a GIVEN set of prune flags both chains are prefix maxima:  with u_b = unpruned cells before b,
//   r_b = max(r_in, max_{unpruned j<b} (v_j - goe + ge (u_j + 1))) - ge u_b,   best_b = max(best_in, max_{unpruned j<b} s_j)
// so a 32-cell segment costs two warp scans.  The flags are found by fixed-point iteration from the guess
// "pruned iff best_in - v_b > X": every pass makes at least one more leading flag final (flag b depends on
// flags < b only), the fixed point is unique and equals the serial result; in practice 1-2 passes.
// Segments of a row run in order with (r, best, left-edge state, previous old best) carried across.
// ================================================================================================
constexpr int DPW_WARPS = 4;            // warps per block
constexpr int DPW_CELLS = 512;          // ring cells per warp (power of two)
constexpr int32_t NEGINF = INT32_MIN / 2 - (1 << 24);

__device__ __forceinline__ int32_t warp_excl_prefix_max(int32_t x, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(FULLW, x, o);
        if (lane >= o) x = max(x, y);
    }
    const int32_t e = __shfl_up_sync(FULLW, x, 1);
    return lane == 0 ? NEGINF : e;
}

__device__ int32_t dp_packed_warp(const uint8_t *B, const uint8_t *A, int32_t N, int32_t M, int32_t &b_offset,
                                  int32_t &a_offset, const int32_t *matrix, int32_t gap_open, int32_t gap_extend,
                                  int32_t x_dropoff, bool reverse, int2 *ring, bool &overflow, int lane)
{
    const int32_t goe = gap_open + gap_extend, ge = gap_extend;
    constexpr int32_t C = DPW_CELLS, MASK = DPW_CELLS - 1;
    a_offset = 0; b_offset = 0;
    if (x_dropoff < goe) x_dropoff = goe;
    if (N <= 0 || M <= 0) return 0;
    const uint32_t lt = (1u << lane) - 1u;

    // row 0: cell 0 = (0, -goe); cell i >= 1 = (-goe - (i-1) ge, that - goe) while the score is >= -X
    int32_t b_size;
    {
        int32_t k = (ge == 0) ? N : (x_dropoff - goe) / ge + 1;     // number of cells i >= 1 with -goe - (i-1) ge >= -X
        k = min(k, N);
        if (k + 1 >= C) { overflow = true; return 0; }
        for (int32_t i = lane; i <= k; i += 32) {
            const int32_t sc = (i == 0) ? 0 : -goe - (i - 1) * ge;
            ring[i & MASK] = make_int2(sc, sc - goe);
        }
        b_size = k + 1;
        __syncwarp();
    }
    int32_t best_score = 0, first_b = 0;

    for (int32_t a_index = 1; a_index <= M; a_index++) {
        int a_bp;
        if (reverse) a_bp = (__ldg(A + (M - a_index) / 4) >> (2 * ((a_index - 1) % 4))) & 3;
        else a_bp = (__ldg(A + 1 + (a_index - 1) / 4) >> (2 * (3 - (a_index - 1) % 4))) & 3;
        const int32_t *mrow = matrix + 16 * a_bp;
        const int32_t row_first = first_b;
        int32_t r_in = MININT, best_in = best_score, prev_old_best = MININT;
        int32_t last_b = first_b, new_first = first_b;
        bool lead = true;

        for (int32_t seg = row_first; seg < b_size; seg += 32) {
            const int32_t b = seg + lane;
            const bool active = b < b_size;
            const uint32_t amask = __ballot_sync(FULLW, active);
            const int2 cell = active ? ring[b & MASK] : make_int2(MININT, MININT);
            int32_t up = __shfl_up_sync(FULLW, cell.x, 1);
            if (lane == 0) up = prev_old_best;
            prev_old_best = __shfl_sync(FULLW, cell.x, 31);
            int32_t v = NEGINF;
            if (active) {
                int32_t d = MININT;
                if (b != row_first) d = up + mrow[(int)__ldg(reverse ? B + N - b : B + b)];
                v = max(d, cell.y);
            }
            // ---- fixed point over the prune flags ------------------------------------------------------
            uint32_t p = __ballot_sync(FULLW, active && (best_in - v > x_dropoff));
            int32_t s = v, R = r_in;
            for (;;) {
                const uint32_t um = amask & ~p;
                const bool unpruned = (um >> lane) & 1u;
                const int32_t u = __popc(um & lt);
                const int32_t w = unpruned ? v - goe + ge * (u + 1) : NEGINF;
                R = max(r_in, warp_excl_prefix_max(w, lane)) - ge * u;
                s = max(v, R);
                const int32_t best_b = max(best_in, warp_excl_prefix_max(unpruned ? s : NEGINF, lane));
                // flags are final from the left: keep what was decided, re-derive from the values they imply
                const uint32_t pn = __ballot_sync(FULLW, active && (best_b - s > x_dropoff));
                if (pn == p) break;
                p = pn;
            }
            const uint32_t um = amask & ~p;
            const bool unpruned = (um >> lane) & 1u;
            // ---- commit the segment --------------------------------------------------------------------
            const int nact = __popc(amask);
            int dropped = 0;                                    // leading pruned cells leave the band
            if (lead) {
                dropped = um ? (__ffs(um) - 1) : nact;
                new_first += dropped;
                lead = (dropped == nact);
            }
            if (active) {
                if (unpruned) ring[b & MASK] = make_int2(s, max(s - goe, cell.y - ge));
                else if (lane >= dropped) ring[b & MASK] = make_int2(MININT, cell.y);
            }
            if (um) {
                last_b = seg + (31 - __clz(um));
                const int32_t m = __reduce_max_sync(FULLW, unpruned ? s : NEGINF);
                if (m > best_in) {
                    const uint32_t at = __ballot_sync(FULLW, unpruned && s == m);
                    best_in = m; a_offset = a_index; b_offset = seg + (__ffs(at) - 1);
                }
            }
            // gap_row leaving the segment: state after its last active cell
            const int32_t r_next = unpruned ? max(s - goe, R - ge) : R;
            r_in = __shfl_sync(FULLW, r_next, nact - 1);
        }
        __syncwarp();
        best_score = best_in;
        first_b = new_first;
        if (first_b == b_size) break;
        if (last_b < b_size - 1) b_size = last_b + 1;
        else {
            // while (gap_row >= best - X && b_size <= N) append (gap_row, gap_row - goe), gap_row -= ge
            int32_t k = 0;
            if (r_in >= best_score - x_dropoff) k = (ge == 0) ? (N - b_size + 1) : ((r_in - (best_score - x_dropoff)) / ge + 1);
            k = max(0, min(k, N - b_size + 1));
            if (b_size + k - first_b + 2 >= C) { overflow = true; return 0; }
            for (int32_t i = lane; i < k; i += 32) {
                const int32_t sc = r_in - i * ge;
                ring[(b_size + i) & MASK] = make_int2(sc, sc - goe);
            }
            b_size += k;
        }
        if (b_size <= N) {
            if (b_size - first_b + 2 >= C) { overflow = true; return 0; }
            if (lane == 0) ring[b_size & MASK] = make_int2(MININT, MININT);
            b_size++;
        }
        __syncwarp();
    }
    return best_score;
}

__global__ void __launch_bounds__(DPW_WARPS * 32)
gapped_warp_kernel(const DevQuery q, const GappedLaunch L)
{
    __shared__ int2 rings[DPW_WARPS][DPW_CELLS];
    __shared__ int32_t s_matrix[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_matrix[i] = q.matrix[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * DPW_WARPS + wib, nwarps = (int64_t)gridDim.x * DPW_WARPS;
    const int64_t n = L.todo ? (int64_t)L.n_todo : (int64_t)min((unsigned long long)L.max_init, *L.n_init);
    int2 *ring = rings[wib];
    for (int64_t w = warp0; w < n; w += nwarps) {
        const int64_t i = L.todo ? (int64_t)L.todo[w] : w;
        const DevInitHit h = L.init[i];
        const DevChunk ch = L.chunks[h.chunk];
        const uint8_t *S = L.packed + ch.byte_off;
        const int32_t context = ctx_search(q, h.q_off);
        const DevContext c = q.ctx[context];
        const uint8_t *query = q.query + c.query_offset;
        const int32_t qlen = c.query_length, slen = ch.len;
        int32_t q_off = h.q_off - c.query_offset, s_off = h.s_off;
        if (h.s_start + h.length >= s_off + 8) { s_off += 3; q_off += 3; }
        // s_BlastDynProgNtGappedAlignment (same as dp_gapped above)
        DevGapResult g;
        g.q_start = g.q_stop = g.s_start = g.s_stop = g.score = g.q_seed = g.s_seed = 0;
        g.status = 0;
        const int32_t adj = 4 - (s_off % 4);
        int32_t q_length = q_off + adj, s_length = s_off + adj;
        if (q_length > qlen || s_length > slen) { q_length -= 4; s_length -= 4; }
        bool overflow = false;
        int32_t pq, ps, right = 0;
        const int32_t left = dp_packed_warp(query, S, q_length, s_length, pq, ps, s_matrix, q.gap_open, q.gap_extend,
                                            q.gap_x_dropoff, true, ring, overflow, lane);
        if (!overflow) {
            g.q_start = q_length - pq; g.s_start = s_length - ps;
            if (q_length < qlen && s_length < slen) {
                int32_t qs, ss;
                __syncwarp();
                right = dp_packed_warp(query + q_length - 1, S + (s_length + 3) / 4 - 1, qlen - q_length, slen - s_length,
                                       qs, ss, s_matrix, q.gap_open, q.gap_extend, q.gap_x_dropoff, false, ring, overflow, lane);
                g.q_stop = qs + q_length; g.s_stop = ss + s_length;
            } else { g.q_stop = q_length; g.s_stop = s_length; }
        }
        if (overflow) g.status = 1;
        else { g.score = left + right; g.q_seed = q_off; g.s_seed = s_off; }
        if (lane == 0) L.out[i] = g;
        __syncwarp();
    }
}

// ================================================================================================
// LONG alignments, latency first: one WARP per (extension, direction); the row's dependency chain runs in ONE lane.
// A real 10 kb alignment is ~10^4 rows of a ~40-cell band, and a batch has a few hundred of them that run side by
// side, so the stage ends when the longest DIRECTION ends: what counts is the time of one row.  Inside a row of
// s_BlastAlignPackedNucl the cells form one chain (score -> horizontal gap / running best -> next score):
//      s_b = max(v_b, r);  pruned_b = best - s_b > X;  unpruned: best = max(best, s_b), r = max(s_b - goe, r - ge)
// with v_b = max(best_old[b-1] + matrix[a][B_b], best_gap_old[b]) free of it.  So per row
//   1. all lanes compute v_b of the band into shared memory            (parallel, two rounds for a 40-cell band)
//   2. lane 0 walks the chain: three dependent operations per cell     (~16 cycles a cell, loads run ahead)
//   3. all lanes write the cells back (score, vertical gap; MININT for a pruned cell)
// ~800 cycles for a 40-cell row.  The lane-per-cell kernel above resolves the same chain with two warp scans and a
// fixed-point loop per 32-cell segment (~4500 cycles per row measured); one thread doing everything pays the
// shared-memory and matrix look-ups inside the chain (~15000).  Values are those of dp_packed, cell for cell.
// ================================================================================================
constexpr int DPL_WARPS = 4;
constexpr int DPL_CELLS = 512;

__device__ int32_t dp_packed_chain(const uint8_t *B, const uint8_t *A, int32_t N, int32_t M, int32_t &b_offset, int32_t &a_offset,
                                   const int32_t *matrix, int32_t gap_open, int32_t gap_extend, int32_t x_dropoff, bool reverse,
                                   int2 *ring, int32_t *vbuf, int32_t *sbuf, bool &overflow, int lane)
{
    const int32_t goe = gap_open + gap_extend, ge = gap_extend;
    constexpr int32_t C = DPL_CELLS, MASK = DPL_CELLS - 1;
    a_offset = 0; b_offset = 0;
    if (x_dropoff < goe) x_dropoff = goe;
    if (N <= 0 || M <= 0) return 0;
    // row 0: cell 0 = (0, -goe); cell i >= 1 = (-goe - (i-1) ge, that - goe) while that score is >= -X
    int32_t b_size;
    {
        int32_t k = (ge == 0) ? N : (x_dropoff - goe) / ge + 1;
        k = min(k, N);
        if (k + 1 >= C) { overflow = true; return 0; }
        for (int32_t i = lane; i <= k; i += 32) {
            const int32_t sc = (i == 0) ? 0 : -goe - (i - 1) * ge;
            ring[i & MASK] = make_int2(sc, sc - goe);
        }
        b_size = k + 1;
        __syncwarp();
    }
    int32_t best = 0, first_b = 0, a_off = 0, b_off = 0;
    for (int32_t a_index = 1; a_index <= M; a_index++) {
        int a_bp;
        if (reverse) a_bp = (__ldg(A + (M - a_index) / 4) >> (2 * ((a_index - 1) % 4))) & 3;
        else a_bp = (__ldg(A + 1 + (a_index - 1) / 4) >> (2 * (3 - (a_index - 1) % 4))) & 3;
        const int32_t *mrow = matrix + 16 * a_bp;
        // 1. v_b for the whole band (the lane keeps its first cell's vertical gap score for step 3)
        int32_t my_y = MININT;
        for (int32_t b = first_b + lane; b < b_size; b += 32) {
            const int2 cell = ring[b & MASK];
            if (b < first_b + 32) my_y = cell.y;
            int32_t diag = MININT;
            if (b > first_b) diag = ring[(b - 1) & MASK].x + mrow[(int)__ldg(reverse ? (B + N - b) : (B + b))];
            vbuf[b & MASK] = max(diag, cell.y);
        }
        __syncwarp();
        // 2. the chain, without branches: per cell  s = max(v, r);  keep = s >= best - X;  r = keep ? max(v - goe, r - ge) : r
        // (= max(s - goe, r - ge): when r > v, s - goe = r - goe <= r - ge as gap_open >= 0);  best = max(best, s).  Three
        // dependent operations from r to r; the loads of v run ahead (unrolled), the stores trail.
        int32_t fb = first_b, last_b = first_b, r = MININT;
        if (lane == 0) {
            int32_t thr = best - x_dropoff;             // prune a cell if s < thr
            bool lead = true;
#pragma unroll 8
            for (int32_t b = first_b; b < b_size; b++) {
                const int32_t v = vbuf[b & MASK];
                const int32_t s = max(v, r);
                const bool keep = s >= thr;
                const bool improved = s > best;          // implies keep
                r = keep ? max(v - goe, r - ge) : r;
                sbuf[b & MASK] = keep ? s : MININT;
                b_off = improved ? b : b_off;
                a_off = improved ? a_index : a_off;
                best = max(best, s);
                thr = max(thr, s - x_dropoff);
                last_b = keep ? b : last_b;
                lead = lead && !keep;
                fb += lead ? 1 : 0;
            }
        }
        __syncwarp();                        // lane 0's sbuf stores before everybody reads them
        fb = __shfl_sync(FULLW, fb, 0); last_b = __shfl_sync(FULLW, last_b, 0); r = __shfl_sync(FULLW, r, 0);
        best = __shfl_sync(FULLW, best, 0);
        // 3. write back: an unpruned cell gets its score and vertical gap, a pruned one inside the band MININT
        for (int32_t b = first_b + lane; b < b_size; b += 32) {
            const int32_t sc = sbuf[b & MASK];
            if (sc != MININT) {
                const int32_t y = (b < first_b + 32) ? my_y : ring[b & MASK].y;
                ring[b & MASK] = make_int2(sc, max(sc - goe, y - ge));
            } else if (b >= fb) ring[b & MASK].x = MININT;
        }
        first_b = fb;
        if (first_b == b_size) break;
        if (last_b < b_size - 1) b_size = last_b + 1;
        else {
            while (r >= (best - x_dropoff) && b_size <= N) {
                if (b_size - first_b + 2 >= C) { overflow = true; return 0; }
                if (lane == 0) ring[b_size & MASK] = make_int2(r, r - goe);
                r -= ge;
                b_size++;
            }
        }
        if (b_size <= N) {
            if (b_size - first_b + 2 >= C) { overflow = true; return 0; }
            if (lane == 0) ring[b_size & MASK] = make_int2(MININT, MININT);
            b_size++;
        }
        __syncwarp();
    }
    a_offset = __shfl_sync(FULLW, a_off, 0); b_offset = __shfl_sync(FULLW, b_off, 0);
    return best;
}

__global__ void __launch_bounds__(DPL_WARPS * 32)
gapped_long_kernel(const DevQuery q, const GappedLaunch L, int2 *halves)
{
    __shared__ int2 rings[DPL_WARPS][DPL_CELLS];
    __shared__ int32_t vbufs[DPL_WARPS][DPL_CELLS], sbufs[DPL_WARPS][DPL_CELLS];
    __shared__ int32_t s_matrix[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_matrix[i] = q.matrix[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * DPL_WARPS + wib, nwarps = (int64_t)gridDim.x * DPL_WARPS;
    const int64_t n = (int64_t)L.n_todo;
    for (int64_t w = warp0; w < 2 * n; w += nwarps) {
        const int64_t i = (int64_t)L.todo[w >> 1];
        const bool do_right = (w & 1) != 0;
        const DevInitHit h = L.init[i];
        const DevChunk ch = L.chunks[h.chunk];
        const uint8_t *S = L.packed + ch.byte_off;
        const DevContext c = q.ctx[ctx_search(q, h.q_off)];
        const uint8_t *query = q.query + c.query_offset;
        const int32_t qlen = c.query_length, slen = ch.len;
        int32_t q_off = h.q_off - c.query_offset, s_off = h.s_off;
        if (h.s_start + h.length >= s_off + 8) { s_off += 3; q_off += 3; }
        const int32_t adj = 4 - (s_off % 4);                           // s_BlastDynProgNtGappedAlignment, as in dp_gapped
        int32_t q_length = q_off + adj, s_length = s_off + adj;
        if (q_length > qlen || s_length > slen) { q_length -= 4; s_length -= 4; }
        bool overflow = false;
        int32_t pq = 0, ps = 0, score = 0;
        DevGapResult *g = &L.out[i];        // the two directions own different fields; score and status are joined afterwards
        __syncwarp();
        if (!do_right) {
            score = dp_packed_chain(query, S, q_length, s_length, pq, ps, s_matrix, q.gap_open, q.gap_extend, q.gap_x_dropoff,
                                    true, rings[wib], vbufs[wib], sbufs[wib], overflow, lane);
            if (lane == 0) { g->q_start = q_length - pq; g->s_start = s_length - ps; g->q_seed = q_off; g->s_seed = s_off; }
        } else if (q_length < qlen && s_length < slen) {
            score = dp_packed_chain(query + q_length - 1, S + (s_length + 3) / 4 - 1, qlen - q_length, slen - s_length, pq, ps,
                                    s_matrix, q.gap_open, q.gap_extend, q.gap_x_dropoff, false, rings[wib], vbufs[wib], sbufs[wib],
                                    overflow, lane);
            if (lane == 0) { g->q_stop = pq + q_length; g->s_stop = ps + s_length; }
        } else if (lane == 0) { g->q_stop = q_length; g->s_stop = s_length; }
        if (lane == 0) halves[w] = make_int2(score, overflow ? 1 : 0);
        __syncwarp();
    }
}

__global__ void gapped_long_combine_kernel(const GappedLaunch L, const int2 *halves)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (int64_t)L.n_todo) return;
    const int64_t i = (int64_t)L.todo[w];
    const int2 l = halves[2 * w], r = halves[2 * w + 1];
    L.out[i].score = l.x + r.x;
    L.out[i].status = (l.y || r.y) ? 1 : 0;           // band wider than the ring: tier 2 takes it
}

// halves: 2 * n_todo int2 of scratch
cudaError_t launch_gapped_long(const DevQuery &q, const GappedLaunch &g, int2 *halves, cudaStream_t st)
{
    if (g.n_todo <= 0) return cudaSuccess;
    const int64_t items = 2 * (int64_t)g.n_todo;
    const int blocks = (int)std::min<int64_t>((items + DPL_WARPS - 1) / DPL_WARPS, 148 * 16);
    gapped_long_kernel<<<blocks, DPL_WARPS * 32, 0, st>>>(q, g, halves);
    gapped_long_combine_kernel<<<(unsigned)((g.n_todo + 127) / 128), 128, 0, st>>>(g, halves);
    return cudaGetLastError();
}

cudaError_t launch_gapped_warp(const DevQuery &q, const GappedLaunch &g, int blocks, cudaStream_t st)
{
    gapped_warp_kernel<<<blocks, DPW_WARPS * 32, 0, st>>>(q, g);
    return cudaGetLastError();
}
int gapped_warp_per_block() { return DPW_WARPS; }

__global__ void __launch_bounds__(GAP_THREADS)
gapped_kernel(const DevQuery q, const GappedLaunch L)
{
    extern __shared__ int2 dp_smem[];                 // tier-1 DP rings [DP_SMEM_CELLS][GAP_THREADS] (DP launches only)
    __shared__ int32_t s_matrix[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_matrix[i] = q.matrix[i];
    __syncthreads();
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t n = L.todo ? (int64_t)L.n_todo : (int64_t)min((unsigned long long)L.max_init, *L.n_init);
    int32_t *scratch = L.scratch ? L.scratch + tid * L.scratch_ints_per_thread : nullptr;

    // work is handed out one extension at a time when a counter is given (extensions differ in length by orders of
    // magnitude: a fixed stride leaves most threads idle behind the slowest), else by a fixed stride
    for (int64_t w = L.work_counter ? (int64_t)atomicAdd(L.work_counter, 1ull) : tid; w < n;
         w = L.work_counter ? (int64_t)atomicAdd(L.work_counter, 1ull) : w + nthreads) {
        const int64_t i = L.todo ? (int64_t)L.todo[w] : w;
        const DevInitHit h = L.init[i];
        const DevChunk ch = L.chunks[h.chunk];
        const uint8_t *S = L.packed + ch.byte_off;
        const int32_t context = ctx_search(q, h.q_off);
        const DevContext c = q.ctx[context];
        const uint8_t *query = q.query + c.query_offset;
        DevGapResult g;
        g.q_start = g.q_stop = g.s_start = g.s_stop = g.score = g.q_seed = g.s_seed = 0;
        g.status = 0;
        if (q.gap_algo == 1) {
            const int32_t q_off = (h.q_start - c.query_offset) + h.length / 2;
            const int32_t s_off = h.s_start + h.length / 2;
            greedy_gapped(q, L.packed, c.query_offset, c.query_length, ch.byte_off * 4, ch.len, q_off, s_off, scratch, L.tier_d, g);
        } else {
            int32_t q_off = h.q_off - c.query_offset, s_off = h.s_off;
            if (h.s_start + h.length >= s_off + 8) { s_off += 3; q_off += 3; }
            if (L.dp_smem_ring == 2)
                dp_gapped(q, s_matrix, query, c.query_length, S, ch.len, q_off, s_off,
                          SmemRing16{reinterpret_cast<short2 *>(dp_smem) + threadIdx.x}, L.dp_max_rows, g);
            else if (L.dp_smem_ring)
                dp_gapped(q, s_matrix, query, c.query_length, S, ch.len, q_off, s_off, SmemRing{dp_smem + threadIdx.x},
                          L.dp_max_rows > 0 ? L.dp_max_rows : INT32_MAX, g);
            else
                dp_gapped(q, s_matrix, query, c.query_length, S, ch.len, q_off, s_off,
                          GlobalRing{reinterpret_cast<int2 *>(scratch), L.tier_d - 1}, INT32_MAX, g);
        }
        L.out[i] = g;
    }
}

cudaError_t launch_gapped(const DevQuery &q, const GappedLaunch &g, cudaStream_t st)
{
    const size_t smem = g.dp_smem_ring == 2 ? (size_t)DP16_CELLS * GAP_THREADS * sizeof(short2)
                        : (g.dp_smem_ring ? (size_t)DP_SMEM_CELLS * GAP_THREADS * sizeof(int2) : 0);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gapped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    gapped_kernel<<<g.grid_blocks > 0 ? g.grid_blocks : GAP_BLOCKS, GAP_THREADS, smem, st>>>(q, g);
    return cudaGetLastError();
}

// Results of the fused pipeline straight into the pinned host mirrors: one small kernel instead of
// "copy the counters, synchronise, copy n_init init-HSPs and n_init gapped results, synchronise" - the number of
// results is only known on the device, and posted PCIe writes of a few tens of KB cost less than one more
// host round trip.  Both record types are 32 bytes (two 16-byte stores).
__global__ void mirror_results_kernel(const DevInitHit *init, const DevGapResult *gap, const unsigned long long *counters,
                                      int64_t cap, DevInitHit *h_init, DevGapResult *h_gap, unsigned long long *h_counters)
{
    static_assert(sizeof(DevInitHit) == 32 && sizeof(DevGapResult) == 32, "two uint4 per record");
    const int64_t n = min((int64_t)counters[2], cap);
    const uint4 *si = reinterpret_cast<const uint4 *>(init), *sg = reinterpret_cast<const uint4 *>(gap);
    uint4 *di = reinterpret_cast<uint4 *>(h_init), *dg = reinterpret_cast<uint4 *>(h_gap);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (int64_t)gridDim.x * blockDim.x) {
        di[i] = si[i];
        dg[i] = sg[i];
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) h_counters[threadIdx.x] = counters[threadIdx.x];
}
cudaError_t launch_mirror_results(const DevInitHit *init, const DevGapResult *gap, const unsigned long long *counters,
                                  int64_t cap, DevInitHit *h_init, DevGapResult *h_gap, unsigned long long *h_counters,
                                  cudaStream_t st)
{
    mirror_results_kernel<<<32, 256, 0, st>>>(init, gap, counters, cap, h_init, h_gap, h_counters);
    return cudaGetLastError();
}

}  // namespace bn
