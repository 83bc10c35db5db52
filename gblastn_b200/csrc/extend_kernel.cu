// extend_kernel.cu — stage 2: diagonal bookkeeping + ungapped X-drop extension.
//
// Replaces (semantics, not code):
//   s_BlastnDiagHashExtendInitialHit   core/na_ungapped.c:779-922  (container eDiagHash)
//   s_BlastnDiagTableExtendInitialHit  core/na_ungapped.c:612-757  (container eDiagArray)
//   s_BlastDiagHashRetrieve / Insert   core/na_ungapped.c:361-450
//   s_TypeOfWord / s_IsSeedMasked / s_MBLookup / s_SmallNaLookup  :489-588, :459-471, :52-106
//   s_NuclUngappedExtend / ...Exact    core/na_ungapped.c:263-350, :153-245
//
// Parallel formulation.  The reference walks all word hits of a subject serially.  With
// window_size == 0 (the blastn / megablast default) the only state shared between hits is
//   * eDiagArray: one cell per (s_off - q_off) & mask            -> independent per cell
//   * eDiagHash : one chain per bucket (512 buckets), INCLUDING the reference's recycling of
//                 "stale" cells while it walks a chain (core/na_ungapped.c:420-428), which makes
//                 a chain's content depend on the order of all insertions into that bucket
//                                                                  -> independent per bucket
// so hits are grouped by cell / bucket, kept in the reference's emission order inside a group, and
// ONE WARP replays a group serially.  The warp runs the control flow uniformly (all lanes hold the
// same values); the ungapped extension — the expensive, data-dependent part — is spread over the
// 32 lanes: the serial recurrence
//        sum += s_k;  if (sum > 0) { score += sum; sum = 0; best = k; }  if (sum < X) break;
// is the same as "T_k = prefix sum, M_k = running (strict) maximum of T, stop at the first k with
// T_k - M_k < X, report M and the first position where it was reached", which is computed per
// block of 32 lanes x (4 four-base steps | 16 bases) with two warp scans.  Sequences are compared
// as 16-base windows (bn_device.cuh); windows touching an ambiguous / sentinel query base take the
// byte-exact slow path so garbage-in semantics of the reference are preserved.
// The hash chains are bit-faithful: prepend at head, overwrite the first stale cell, carry-over
// across subjects through `offset`, reset when offset passes INT4_MAX/4 (chunk table epochs).
#include "bn_device.cuh"

namespace bn {

constexpr int EXT_WARPS_PER_BLOCK = 4;
constexpr int EXT_BLOCKS = 148 * 4;
constexpr int SPEC_BLOCKS = 148 * 12;     // speculative pass: 48 warps per SM
constexpr unsigned FULL = 0xffffffffu;

struct Ungapped { int32_t q_start, s_start, length, score; };

// ---- warp scans ---------------------------------------------------------------------------------
__device__ __forceinline__ long long warp_excl_sum(long long v, int lane, long long &total)
{
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}
__device__ __forceinline__ long long warp_excl_max(long long v, int lane, long long identity)
{
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x = max(x, y);
    }
    long long e = __shfl_up_sync(FULL, x, 1);
    return lane == 0 ? identity : e;
}

// State of one direction of an X-drop walk, uniform across the warp.
struct Walk {
    long long T;       // running total
    long long M;       // running strict maximum (starts at 0: "no improvement yet")
    int32_t best;      // index of the step that set M, -1 if none
    bool stopped;
};

// Consume one block of per-lane step values.  v[0..cnt) are this lane's consecutive steps, the
// lane's first step has global index k0.  NV = steps per lane.  Returns with W updated; W.stopped
// set when the X-drop test fired inside the block.
template <int NV>
__device__ __forceinline__ void walk_block(Walk &W, const int32_t (&v)[NV], int cnt, int32_t k0, int32_t X,
                                           int lane)
{
    long long S = 0, P = LLONG_MIN;
#pragma unroll
    for (int j = 0; j < NV; j++)
        if (j < cnt) { S += v[j]; P = max(P, S); }
    long long total;
    const long long off = warp_excl_sum(S, lane, total);
    const long long Tstart = W.T + off;
    const long long cand = (cnt > 0) ? Tstart + P : LLONG_MIN;
    const long long Mstart = max(W.M, warp_excl_max(cand, lane, LLONG_MIN));
    // second pass with the true running values
    long long T = Tstart, M = Mstart;
    int32_t arg = -1;
    int brk = -1;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        if (j < cnt && brk < 0) {
            T += v[j];
            if (T > M) { M = T; arg = k0 + j; }
            if (T - M < (long long)X) brk = j;
        }
    }
    const unsigned bmask = __ballot_sync(FULL, brk >= 0);
    const int last_lane = bmask ? (__ffs(bmask) - 1) : 31;
    const unsigned raised = __ballot_sync(FULL, arg >= 0) & (last_lane == 31 ? FULL : ((2u << last_lane) - 1));
    if (raised) {
        const int src = 31 - __clz(raised);
        W.best = __shfl_sync(FULL, arg, src);
    }
    W.M = __shfl_sync(FULL, M, last_lane);
    W.T = __shfl_sync(FULL, T, last_lane);   // only meaningful when not stopped
    W.stopped = bmask != 0;
}

// ---- approximate extension: 4 bases per step (s_NuclUngappedExtend) ------------------------------
// raw 4-base byte of the query exactly as the reference builds it (values >= 4 overlap bit fields)
__device__ __forceinline__ uint32_t qbyte_raw(const uint8_t *query, int32_t p)
{
    return (((uint32_t)__ldg(query + p) << 6) | ((uint32_t)__ldg(query + p + 1) << 4) |
            ((uint32_t)__ldg(query + p + 2) << 2) | (uint32_t)__ldg(query + p + 3)) & 0xFFu;
}

// Right: steps k = 0.. cover query [q_ext + 4k, q_ext + 4k + 4), subject bytes from s_ext (4-aligned).
// Left : steps k = 0.. cover query [q_ext - 4k - 4, q_ext - 4k).
template <bool RIGHT>
__device__ void approx_walk(const DevQuery &q, const uint8_t *packed, int64_t chunk_base,
                            const int32_t *s_tab, int32_t q_ext, int32_t s_ext, int32_t nsteps, int32_t X,
                            int lane, Walk &W)
{
    W.T = 0; W.M = 0; W.best = -1; W.stopped = false;
    for (int32_t base = 0; base < nsteps && !W.stopped; base += 128) {
        const int32_t k0 = base + 4 * lane;
        const int cnt = max(0, min(4, nsteps - k0));
        int32_t v[4] = {0, 0, 0, 0};
        if (cnt > 0) {
            const int32_t qpos = RIGHT ? q_ext + 4 * k0 : q_ext - 4 * k0 - 16;
            const int64_t spos = chunk_base + (RIGHT ? (int64_t)s_ext + 4 * k0 : (int64_t)s_ext - 4 * k0 - 16);
            uint32_t qb, qa;
            qwin(q, qpos, qb, qa);
            const uint32_t sb = swin(packed, spos);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j >= cnt) break;
                // step j of this lane = byte j of the window (right) or byte 3-j (left)
                const int byte = RIGHT ? j : 3 - j;
                const uint32_t sh = 24 - 8 * byte;
                uint32_t qq = (qb >> sh) & 0xFFu;
                if ((qa >> sh) & 0xFFu) qq = qbyte_raw(q.query, qpos + 4 * byte);
                v[j] = s_tab[qq ^ ((sb >> sh) & 0xFFu)];
            }
        }
        walk_block<4>(W, v, cnt, k0, X, lane);
    }
}

// ---- exact extension: 1 base per step (s_NuclUngappedExtendExact) ---------------------------------
template <bool RIGHT>
__device__ void exact_walk(const DevQuery &q, const uint8_t *packed, int64_t chunk_base, int32_t q_off,
                           int32_t s_off, int32_t nsteps, int32_t X, int lane, Walk &W)
{
    W.T = 0; W.M = 0; W.best = -1; W.stopped = false;
    const int32_t reward = q.reward, penalty = q.penalty;
    for (int32_t base = 0; base < nsteps && !W.stopped; base += 512) {
        const int32_t k0 = base + 16 * lane;
        const int cnt = max(0, min(16, nsteps - k0));
        int32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = 0;
        if (cnt > 0) {
            const int32_t qpos = RIGHT ? q_off + k0 : q_off - k0 - 16;
            const int64_t spos = chunk_base + (RIGHT ? (int64_t)s_off + k0 : (int64_t)s_off - k0 - 16);
            uint32_t qb, qa;
            qwin(q, qpos, qb, qa);
            const uint32_t sb = swin(packed, spos);
            const uint32_t mm = mismatch_bits(qb, 0u, sb);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j >= cnt) break;
                const int b = RIGHT ? j : 15 - j;          // base index inside the window
                const uint32_t sh = 30 - 2 * b;
                int32_t val = ((mm >> sh) & 1u) ? penalty : reward;
                if ((qa >> sh) & 1u)                        // ambiguity code or sentinel: matrix row
                    val = __ldg(&q.matrix[16 * (int)__ldg(q.query + qpos + b) + (int)((sb >> sh) & 3u)]);
                v[j] = val;
            }
        }
        walk_block<16>(W, v, cnt, k0, X, lane);
    }
}

__device__ void ungapped_exact(const DevQuery &q, const uint8_t *packed, int64_t chunk_base, int32_t slen,
                               int32_t q_off, int32_t s_off, int32_t X, int lane, Ungapped &u)
{
    Walk W;
    // left: while subject position > max(0, s_off - q_off); the sentinel row (INT4_MIN/2) stops it
    const int32_t nleft = (q_off < s_off) ? q_off : s_off;
    exact_walk<false>(q, packed, chunk_base, q_off, s_off, nleft, X, lane, W);
    int32_t score = (int32_t)W.M;
    const int32_t q_beg = W.best >= 0 ? q_off - 1 - W.best : q_off;
    u.q_start = q_beg;
    u.s_start = s_off - (q_off - q_beg);
    const int32_t q_avail = q.concat_len - q_off, s_avail = slen - s_off;
    const int32_t nright = (q_avail < s_avail) ? q_avail : s_avail;
    exact_walk<true>(q, packed, chunk_base, q_off, s_off, nright, X, lane, W);
    score += (int32_t)W.M;
    const int32_t q_end = W.best >= 0 ? q_off + W.best + 1 : q_off;
    u.length = q_end - q_beg;
    u.score = score;
}

__device__ void ungapped_extend(const DevQuery &q, const uint8_t *packed, int64_t chunk_base, int32_t slen,
                                const int32_t *s_tab, int32_t q_off, int32_t s_match_end, int32_t s_off,
                                int32_t X, int32_t reduced_cutoff, int lane, Ungapped &u)
{
    const int32_t shift = (4 - (s_off % 4)) % 4;
    const int32_t q_ext = q_off + shift, s_ext = s_off + shift;
    Walk W;
    approx_walk<false>(q, packed, chunk_base, s_tab, q_ext, s_ext, min(q_ext, s_ext) / 4, X, lane, W);
    int32_t score = (int32_t)W.M;
    int32_t new_q = W.best >= 0 ? q_ext - 4 * (W.best + 1) : q_ext;
    u.q_start = new_q;
    u.s_start = s_ext - (q_ext - new_q);
    approx_walk<true>(q, packed, chunk_base, s_tab, q_ext, s_ext,
                      min(q.concat_len - q_ext, slen - s_ext) / 4, X, lane, W);
    score += (int32_t)W.M;
    new_q = W.best >= 0 ? q_ext + 4 * W.best + 3 : q_ext;
    if (score >= reduced_cutoff) {
        ungapped_exact(q, packed, chunk_base, slen, q_off, s_off, X, lane, u);
    } else {
        u.score = score;
        u.length = max(s_match_end - u.s_start, new_q - u.q_start + 1);
    }
}

// ---- lookup probes (warp-uniform) ----------------------------------------------------------------
__device__ bool lut_contains(const DevQuery &q, uint32_t index, int32_t q_pos)
{
    if (q.lut_type == 0) {
        int32_t v = mb_cell(q, (uint32_t)index & q.hash_mask);
        ++q_pos;
        while (v) {
            if (v == q_pos) return true;
            v = __ldg(&q.next_pos[v]);
        }
        return false;
    }
    if (q.lut_type == 2) {          // s_NaLookup core/na_ungapped.c:112-138
        const int4 cell = __ldg(&q.na_cells[index & q.hash_mask]);
        const int32_t nh = cell.x;
        if (nh <= 3) return (nh > 0 && cell.y == q_pos) || (nh > 1 && cell.z == q_pos) || (nh > 2 && cell.w == q_pos);
        for (int32_t i = 0; i < nh; i++)
            if (__ldg(&q.na_overflow[cell.y + i]) == q_pos) return true;
        return false;
    }
    int32_t v = __ldg(&q.backbone[index & q.hash_mask]);
    if (v == q_pos) return true;
    if (v == -1 || v >= 0) return false;
    int32_t src = -v;
    v = __ldg(&q.overflow[src++]);
    do {
        if (v == q_pos) return true;
        v = __ldg(&q.overflow[src++]);
    } while (v >= 0);
    return false;
}

__device__ __forceinline__ bool seed_masked(const DevQuery &q, const uint8_t *S, int32_t s_off,
                                            int32_t lut, int32_t q_pos)
{
    uint32_t w = be32(S + s_off / 4);
    return !lut_contains(q, w >> (2 * (16 - s_off % 4 - lut)), q_pos);
}

// s_TypeOfWord (core/na_ungapped.c:489-588): 0 = not a word, 1 = single word, 2 = double word
__device__ int type_of_word(const DevQuery &q, const uint8_t *S, int32_t &q_off, int32_t &s_off,
                            bool has_locations, uint32_t s_range, int32_t word_length, int32_t lut,
                            bool check_double, int32_t &extended, int lane)
{
    extended = 0;
    if (word_length == lut) return 1;
    // one-hit mode without masked locations: q_end - q_off stays word_length, ext_to is 0, nothing runs
    if (!has_locations && !check_double) return 1;
    int32_t q_end = q_off + word_length, s_end = s_off + word_length;
    const int32_t context = ctx_search_warp(q, q_end, lane);
    const int32_t q_range = __ldg(&q.ctx[context].query_offset) + __ldg(&q.ctx[context].query_length);
    if (has_locations) {
        if (seed_masked(q, S, s_end - lut, lut, q_end - lut)) return 0;
        for (;; ++s_off, ++q_off)
            if (!seed_masked(q, S, s_off, lut, q_off)) break;
    }
    int32_t ext_to = word_length - (q_end - q_off);
    const uint32_t a = (uint32_t)(q_range - q_end), c = s_range - (uint32_t)s_end;
    int32_t ext_max = (int32_t)(a > c ? c : a);   // unsigned MIN, as in the reference (:534)
    if (ext_to || has_locations) {
        if (ext_to > ext_max) return 0;
        q_end += ext_to; s_end += ext_to;
        for (int32_t s_pos = s_end - lut, q_pos = q_end - lut; s_pos > s_off; s_pos -= lut, q_pos -= lut)
            if (seed_masked(q, S, s_pos, lut, q_pos)) return 0;
        extended = ext_to;
    }
    if (!check_double) return 1;
    // right extension to a double word: seed by seed, then base by base (:557-585)
    ext_to += word_length;
    ext_max = min(ext_max, ext_to);
    int32_t s_pos = s_end, q_pos = q_end;
    for (; (uint32_t)extended + (uint32_t)lut <= (uint32_t)ext_max; s_pos += lut, q_pos += lut, extended += lut)
        if (seed_masked(q, S, s_pos, lut, q_pos)) break;
    s_pos -= lut - 1; q_pos -= lut - 1;
    while (extended < ext_max) {
        if (seed_masked(q, S, s_pos, lut, q_pos)) return 1;
        ++extended; ++s_pos; ++q_pos;
    }
    return ext_max == ext_to ? 2 : 1;
}

// ---- one word hit: s_TypeOfWord + ungapped extension + cutoff test -----------------------------------
// The outcome depends only on (q_off, s_off, chunk), never on the diagonal container, so it can be
// computed for many hits at once (speculative pass) or inline by the replay warp.  Warp-uniform.
constexpr int32_t SPEC_NONE = 0;       // not computed speculatively
constexpr int32_t SPEC_MASKED = 2;     // s_TypeOfWord returned 0
constexpr int32_t SPEC_LOW = 3;        // extended, score below the cutoff
constexpr int32_t SPEC_READY = 4;      // extended, score >= cutoff_score
constexpr int32_t SPEC_SINGLE = 5;     // two-hit mode: single word, no neighbour yet -> recorded, not extended

struct CtxCache { int32_t lo, hi, x_dropoff, cutoff_score, reduced_cutoff; };   // context covering [lo, hi)

__device__ __forceinline__ void ctx_lookup(const DevQuery &q, int32_t q_off, int lane, CtxCache &cc)
{
    if (q_off >= cc.lo && q_off < cc.hi) return;
    const int32_t context = ctx_search_warp(q, q_off, lane);
    const DevContext c = q.ctx[context];
    cc.lo = c.query_offset;
    cc.hi = (context + 1 < q.num_contexts) ? __ldg(&q.ctx[context + 1].query_offset) : INT32_MAX;
    cc.x_dropoff = c.x_dropoff; cc.cutoff_score = c.cutoff_score; cc.reduced_cutoff = c.reduced_cutoff;
}

// touch the cache lines both X-drop walks are about to read (subject +-2 kb, query windows +-2 kb)
__device__ __forceinline__ void prefetch_around(const DevQuery &q, const uint8_t *packed, const DevChunk &ch,
                                                int32_t q_off, int32_t s_off, int lane)
{
    if (lane < 8) {
        int64_t b = ch.byte_off + (s_off >> 2) + (int64_t)(lane - 4) * 128;
        b = max(b, ch.byte_off); b = min(b, ch.byte_off + (ch.len >> 2));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(packed + b));
    } else if (lane < 24) {
        int32_t i = ((q_off + 16) >> 4) + (lane - 16) * 16;
        i = max(i, 0); i = min(i, (q.concat_len + 18) >> 4);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q.qpk + i));
    }
}

// s_range of the extension callbacks (core/na_ungapped.c:1634-1637): the right end of the unmasked range the
// seed was scanned in; the chunk length when the volume carries no database masks
__device__ __forceinline__ int32_t hit_s_range(const int2 *ranges, const DevChunk &ch, int32_t scan_pos)
{
    if (ch.n_ranges == 0) return ch.len;
    int32_t lo = 0, hi = ch.n_ranges - 1;
    while (lo < hi) {                         // last range with left <= scan_pos
        const int32_t m = (lo + hi + 1) >> 1;
        if (__ldg(&ranges[ch.range_first + m].x) <= scan_pos) lo = m; else hi = m - 1;
    }
    return __ldg(&ranges[ch.range_first + lo].y);
}

__device__ void extend_one(const DevQuery &q, const uint8_t *packed, const DevChunk &ch, int32_t s_range, const int32_t *s_tab,
                           bool is_hash, bool has_loc, int32_t word, int32_t lut, bool direct, bool check_double,
                           int32_t q_off, int32_t s_off, int lane, CtxCache &cc, SpecResult &r)
{
    int32_t extended = 0;
    const int32_t s_end0 = s_off + word;      // the reference fixes s_end before s_TypeOfWord may shift s_off
    const int word_type = type_of_word(q, packed + ch.byte_off, q_off, s_off, has_loc, (uint32_t)s_range, word,
                                       direct ? word : lut, check_double, extended, lane);
    if (!word_type) {
        r.status = SPEC_MASKED;
        return;
    }
    r.q_off = q_off; r.s_off = s_off; r.extended = extended;
    if (check_double && word_type == 1) {     // off-diagonal search needs scan_range > 0 (rejected at load)
        r.status = SPEC_SINGLE;
        return;
    }
    ctx_lookup(q, q_off, lane, cc);
    Ungapped u;
    const int64_t chunk_base = ch.byte_off * 4;
    if (!is_hash && word < 11)
        ungapped_exact(q, packed, chunk_base, ch.len, q_off, s_off, -cc.x_dropoff, lane, u);
    else
        ungapped_extend(q, packed, chunk_base, ch.len, s_tab, q_off, s_end0 + extended, s_off, -cc.x_dropoff,
                        cc.reduced_cutoff, lane, u);
    r.status = (u.score >= cc.cutoff_score) ? SPEC_READY : SPEC_LOW;
    r.q_start = u.q_start; r.s_start = u.s_start; r.length = u.length; r.score = u.score;
}

// ---- bucket chain (BLAST_DiagHash restricted to one bucket) -------------------------------------
// cell = int4 {diag, level, (hit_len << 1) | hit_saved, next}; index 0 = null.  All lanes execute
// the walk on identical values; lane 0 stores.
// The first CHAIN_SMEM cells of a chain live in shared memory (stale cells are recycled in place, so a
// bucket's chain stays far shorter than its hit count: a walk step costs a shared-memory access instead
// of a dependent L2 round trip); cells beyond that spill to the group's region of global memory.
constexpr int CHAIN_SMEM = 255;
struct Chain {
    int4 *cells;       // global region, index 1..
    int4 *fast;        // shared memory, index 1..CHAIN_SMEM
    int32_t head, used;
    __device__ __forceinline__ int4 load(int32_t i) const { return i <= CHAIN_SMEM ? fast[i] : cells[i]; }
    __device__ __forceinline__ void store(int32_t i, int4 v) const { if (i <= CHAIN_SMEM) fast[i] = v; else cells[i] = v; }
};

__device__ __forceinline__ bool chain_get(const Chain &c, int32_t diag, int32_t &level, int32_t &saved)
{
    int32_t i = c.head;
    while (i) {
        const int4 v = c.load(i);
        if (v.x == diag) { level = v.y; saved = v.z & 1; return true; }
        i = v.w;
    }
    return false;
}

__device__ __forceinline__ void chain_put(Chain &c, int32_t diag, int32_t level, int32_t len,
                                          int32_t saved, int32_t s_off_pos, int32_t window, int lane)
{
    int32_t i = c.head;
    while (i) {
        const int4 v = c.load(i);
        if (v.x == diag || s_off_pos - v.y > window) {
            if (lane == 0) c.store(i, make_int4(diag, level, (len << 1) | saved, v.w));
            __syncwarp();
            return;
        }
        i = v.w;
    }
    const int32_t n = ++c.used;
    if (lane == 0) c.store(n, make_int4(diag, level, (len << 1) | saved, c.head));
    __syncwarp();
    c.head = n;
}

__device__ __forceinline__ uint32_t diag_bucket(int32_t diag)
{
    return ((uint32_t)diag * 0x9E370001u) % 512u;
}

// Sorted key = (group << gbits) | global scan position.  Two consecutive hits belong to the same
// replay group when their group fields match and, for the diagonal ARRAY (whose cells are
// independent per subject chunk), they come from the same chunk.
__device__ __forceinline__ bool same_group(const uint64_t *keys, const SeedHit *hits, int64_t a, int64_t b,
                                           int gbits, bool is_hash)
{
    if ((keys[a] >> gbits) != (keys[b] >> gbits)) return false;
    return is_hash || hits[a].chunk == hits[b].chunk;
}

// heads[i] = index of the first hit of group i (any order); leaders[] = hits that open a run of
// consecutive hits on one diagonal of one chunk (the hits the replay is most likely to extend)
__global__ void group_heads_kernel(const uint64_t *keys, const SeedHit *hits, int64_t n, int gbits,
                                   int is_hash, int spec_enabled, uint32_t *heads, uint32_t *leaders,
                                   SpecResult *spec, unsigned long long *counters)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    bool head = true, leader = spec_enabled != 0;
    if (j > 0) {
        const SeedHit a = hits[j - 1], b = hits[j];
        const bool same_grp = (keys[j - 1] >> gbits) == (keys[j] >> gbits) && (is_hash || a.chunk == b.chunk);
        head = !same_grp;
        leader = spec_enabled && (head || a.chunk != b.chunk || (a.s_off - a.q_off) != (b.s_off - b.q_off));
    }
    spec[j].status = SPEC_NONE;
    if (head) heads[atomicAdd(&counters[4], 1ull)] = (uint32_t)j;
    if (leader) {
        // warp-aggregated append
        const unsigned m = __activemask();
        const int lane = threadIdx.x & 31, ldr = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == ldr) base = atomicAdd(&counters[5], (unsigned long long)__popc(m));
        base = __shfl_sync(m, base, ldr);
        leaders[base + __popc(m & ((1u << lane) - 1))] = (uint32_t)j;
    }
}


// ---- one leader per LANE (blastn mode, plain reward / penalty scoring) -----------------------------------------------
// The warp-wide walks above take 128 (approximate) or 512 (exact) bases per step: right for the long extensions of a
// megablast batch, ~3000 warp instructions too many for the hit of a blastn-mode batch whose extension dies after a
// dozen bases — and those are 5.6 M per pass of C3.  Here a lane runs the same two recurrences alone, 16-base windows,
// group scores in closed form (4 r - (r - p) * mismatches, what nucl_score_table holds for such a batch: Query::direct_ok).
// Returns false = undecided (an ambiguity code or sentinel in a window it needs, or an extension beyond SC_LIMIT bases):
// the warp then extends that hit together, as before.
constexpr int32_t SC_LIMIT = 128;
__device__ bool scalar_extend_direct(const DevQuery &q, const uint8_t *packed, int64_t chunk_base, int32_t slen, int32_t q_off,
                                     int32_t s_off, int32_t s_match_end, int32_t X, int32_t reduced, Ungapped &u)
{
    const int32_t r = q.reward, pen = q.penalty, r4 = 4 * r, dd = r - pen;
    const int32_t shift = (4 - (s_off % 4)) % 4;
    const int32_t q_ext = q_off + shift, s_ext = s_off + shift;
    int32_t score, new_q_r;
    {   // approximate pass, left: step k covers query [q_ext - 4k - 4, q_ext - 4k)
        const int32_t n = min(q_ext, s_ext) >> 2;
        int32_t M = 0, sum = 0, best = -1;
        bool stop = false;
        for (int32_t k = 0; k < n && !stop; k += 4) {
            if (4 * k >= SC_LIMIT) return false;
            uint32_t qb, qa;
            qwin(q, q_ext - 4 * k - 16, qb, qa);
            const int cnt = min(4, n - k);
            if (qa & (cnt == 4 ? 0xFFFFFFFFu : ((1u << (8 * cnt)) - 1u))) return false;
            const uint32_t m = mismatch_bits(qb, 0u, swin(packed, chunk_base + s_ext - 4 * k - 16));
            for (int j = 0; j < cnt; j++) {
                sum += r4 - dd * __popc((m >> (8 * j)) & 0xFFu);
                if (sum > 0) { M += sum; sum = 0; best = k + j; }
                if (sum < X) { stop = true; break; }
            }
        }
        score = M;
        const int32_t new_q = best >= 0 ? q_ext - 4 * (best + 1) : q_ext;
        u.q_start = new_q;
        u.s_start = s_ext - (q_ext - new_q);
    }
    {   // approximate pass, right: step k covers query [q_ext + 4k, q_ext + 4k + 4)
        const int32_t n = min(q.concat_len - q_ext, slen - s_ext) >> 2;
        int32_t M = 0, sum = 0, best = -1;
        bool stop = false;
        for (int32_t k = 0; k < n && !stop; k += 4) {
            if (4 * k >= SC_LIMIT) return false;
            uint32_t qb, qa;
            qwin(q, q_ext + 4 * k, qb, qa);
            const int cnt = min(4, n - k);
            if (qa & (cnt == 4 ? 0xFFFFFFFFu : ~((1u << (32 - 8 * cnt)) - 1u))) return false;
            const uint32_t m = mismatch_bits(qb, 0u, swin(packed, chunk_base + s_ext + 4 * k));
            for (int j = 0; j < cnt; j++) {
                sum += r4 - dd * __popc((m >> (24 - 8 * j)) & 0xFFu);
                if (sum > 0) { M += sum; sum = 0; best = k + j; }
                if (sum < X) { stop = true; break; }
            }
        }
        score += M;
        new_q_r = best >= 0 ? q_ext + 4 * best + 3 : q_ext;
    }
    if (score < reduced) {
        u.score = score;
        u.length = max(s_match_end - u.s_start, new_q_r - u.q_start + 1);
        return true;
    }
    // exact pass (s_NuclUngappedExtendExact), one base per step
    int32_t q_beg;
    {
        const int32_t n = min(q_off, s_off);
        int32_t M = 0, sum = 0, best = -1;
        bool stop = false;
        for (int32_t done = 0; done < n && !stop; done += 16) {
            if (done >= SC_LIMIT) return false;
            uint32_t qb, qa;
            qwin(q, q_off - done - 16, qb, qa);
            const int rem = min(16, n - done);
            if (qa & (rem == 16 ? 0xFFFFFFFFu : ((1u << (2 * rem)) - 1u))) return false;     // step t = base 15 - t, flag at bit 2t
            const uint32_t m = mismatch_bits(qb, 0u, swin(packed, chunk_base + s_off - done - 16));
            for (int t = 0; t < rem; t++) {
                sum += ((m >> (2 * t)) & 1u) ? pen : r;
                if (sum > 0) { M += sum; sum = 0; best = done + t; }
                if (sum < X) { stop = true; break; }
            }
        }
        score = M;
        q_beg = best >= 0 ? q_off - 1 - best : q_off;
        u.q_start = q_beg;
        u.s_start = s_off - (q_off - q_beg);
    }
    {
        const int32_t n = min(q.concat_len - q_off, slen - s_off);
        int32_t M = 0, sum = 0, best = -1;
        bool stop = false;
        for (int32_t done = 0; done < n && !stop; done += 16) {
            if (done >= SC_LIMIT) return false;
            uint32_t qb, qa;
            qwin(q, q_off + done, qb, qa);
            const int rem = min(16, n - done);
            if (qa & (rem == 16 ? 0xFFFFFFFFu : ~((1u << (32 - 2 * rem)) - 1u))) return false;   // step t = base t, flag at bit 30 - 2t
            const uint32_t m = mismatch_bits(qb, 0u, swin(packed, chunk_base + s_off + done));
            for (int t = 0; t < rem; t++) {
                sum += ((m >> (30 - 2 * t)) & 1u) ? pen : r;
                if (sum > 0) { M += sum; sum = 0; best = done + t; }
                if (sum < X) { stop = true; break; }
            }
        }
        score += M;
        const int32_t q_end = best >= 0 ? q_off + best + 1 : q_off;
        u.length = q_end - q_beg;
        u.score = score;
    }
    return true;
}

// Speculative pass: one warp per leader runs s_TypeOfWord + the ungapped extension and parks the
// outcome next to the hit; the replay consumes it if (and only if) the diagonal test lets the hit
// through, exactly where the reference would have extended.
__global__ void __launch_bounds__(EXT_WARPS_PER_BLOCK * 32)
extend_leaders_kernel(const DevQuery q, const ExtendLaunch e)
{
    __shared__ int32_t s_tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = q.score_table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * EXT_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * EXT_WARPS_PER_BLOCK;
    if (e.n_from_device && e.counters[6]) return;
    const int64_t n = (int64_t)e.counters[5];
    const bool is_hash = q.container_type == 1;
    const int32_t word = q.word_length, lut = q.lut_word_length;
    const bool direct = (word == lut);
    const bool has_loc = q.has_locations && !direct;
    CtxCache cc{0, -1, 0, 0, 0};
    if (e.scalar_ok && direct) {
        // one leader per lane; what a lane cannot decide alone is extended by the warp afterwards
        for (int64_t w0 = warp0 * 32; w0 < n; w0 += nwarps * 32) {
            const int64_t w = w0 + lane;
            uint32_t j = 0;
            bool pending = false;
            if (w < n) {
                j = e.leaders[w];
                const SeedHit h = e.hits[j];
                const DevChunk ch = e.chunks[h.chunk];
                int32_t xd = e.uni_x, co = e.uni_cutoff, rc = e.uni_reduced;
                if (!e.uni_ok) {
                    const DevContext c = q.ctx[ctx_search(q, (int32_t)h.q_off)];
                    xd = c.x_dropoff; co = c.cutoff_score; rc = c.reduced_cutoff;
                }
                Ungapped u;
                if (scalar_extend_direct(q, e.packed, ch.byte_off * 4, ch.len, (int32_t)h.q_off, (int32_t)h.s_off,
                                         (int32_t)h.s_off + word, -xd, rc, u)) {
                    SpecResult r;
                    r.status = (u.score >= co) ? SPEC_READY : SPEC_LOW;
                    r.q_off = (int32_t)h.q_off; r.s_off = (int32_t)h.s_off; r.extended = 0;
                    r.q_start = u.q_start; r.s_start = u.s_start; r.length = u.length; r.score = u.score;
                    e.spec[j] = r;
                } else pending = true;
            }
            unsigned m = __ballot_sync(0xffffffffu, pending);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1u;
                const uint32_t jj = __shfl_sync(0xffffffffu, j, src);
                const SeedHit h = e.hits[jj];
                const DevChunk ch = e.chunks[h.chunk];
                SpecResult r;
                r.status = SPEC_NONE; r.q_off = r.s_off = r.extended = r.q_start = r.s_start = r.length = r.score = 0;
                extend_one(q, e.packed, ch, hit_s_range(e.ranges, ch, (int32_t)h.scan_pos), s_tab, is_hash, has_loc, word, lut,
                           direct, false, (int32_t)h.q_off, (int32_t)h.s_off, lane, cc, r);
                if (lane == 0) e.spec[jj] = r;
            }
        }
        return;
    }
    for (int64_t w = warp0; w < n; w += nwarps) {
        const uint32_t j = e.leaders[w];
        const SeedHit h = e.hits[j];
        const DevChunk ch = e.chunks[h.chunk];
        prefetch_around(q, e.packed, ch, (int32_t)h.q_off, (int32_t)h.s_off, lane);
        SpecResult r;
        r.status = SPEC_NONE; r.q_off = r.s_off = r.extended = r.q_start = r.s_start = r.length = r.score = 0;
        extend_one(q, e.packed, ch, hit_s_range(e.ranges, ch, (int32_t)h.scan_pos), s_tab, is_hash, has_loc, word, lut,
                   direct, false, (int32_t)h.q_off, (int32_t)h.s_off, lane, cc, r);
        if (lane == 0) e.spec[j] = r;
    }
}

__global__ void __launch_bounds__(EXT_WARPS_PER_BLOCK * 32)
extend_kernel(const DevQuery q, const ExtendLaunch e, const uint64_t *keys, const uint32_t *heads,
              int64_t n_hits, int gbits)
{
    if (e.n_from_device) {
        if (e.counters[6]) return;
        n_hits = (int64_t)e.counters[0];
    }
    __shared__ int32_t s_tab[256];
    __shared__ int4 s_chain[EXT_WARPS_PER_BLOCK][CHAIN_SMEM + 1];
    // the batch of 32 hits a warp replays serially: staged in shared memory so that every step reads its
    // hit with broadcast loads whose addresses do not depend on the previous step (no shuffle chain)
    __shared__ SeedHit s_hit[EXT_WARPS_PER_BLOCK][32];
    __shared__ SpecResult s_spec[EXT_WARPS_PER_BLOCK][32];
    __shared__ uint64_t s_key[EXT_WARPS_PER_BLOCK][32];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = q.score_table[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * EXT_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * EXT_WARPS_PER_BLOCK;
    const int64_t n_groups = (int64_t)e.counters[4];

    const bool is_hash = q.container_type == 1;
    const int32_t word = q.word_length, lut = q.lut_word_length;
    const bool direct = (word == lut);
    const bool has_loc = q.has_locations && !direct;
    // one-hit mode: Delta = MIN(scan_range, -word_length); two-hit mode with scan_range == 0: Delta <= 0,
    // clamped to 0 on the single-word path only (core/na_ungapped.c:833)
    const int32_t window = q.window_size;
    const bool two_hits = window > 0;
    const int32_t Delta = min(q.scan_range, window - word);

    for (int64_t g = warp0; g < n_groups; g += nwarps) {
        const int64_t j0 = (int64_t)heads[g];
        Chain chain;
        chain.cells = reinterpret_cast<int4 *>(e.cells) + j0;   // region [j0+1 .. j0+group_size]
        chain.fast = s_chain[threadIdx.x >> 5];
        chain.head = 0; chain.used = 0;
        __syncwarp();
        int32_t last_hit_cell = 0, flag_cell = 0;      // eDiagArray: the cell's last_hit / flag
        uint32_t cur_chunk = 0xFFFFFFFFu;
        int32_t cur_epoch = -1;
        DevChunk ch{};
        unsigned long long n_extended = 0;
        CtxCache cc{0, -1, 0, 0, 0};

        // most-recently-stored diagonal of this bucket: a get() right after put(d) returns what was
        // stored (no other put intervened), so runs of seeds on one diagonal skip the chain walk
        int32_t cache_diag = 0, cache_level = 0, cache_saved = 0;
        bool cache_ok = false;
        bool group_done = false;
        for (int64_t base = j0; base < n_hits && !group_done; base += 32) {
            // prefetch up to 32 hits of the group, one per lane
            const int64_t jm = base + lane;
            SeedHit mine{0, 0, 0, 0};
            SpecResult myspec;
            myspec.status = SPEC_NONE;
            uint64_t mykey = 0;
            bool in_group = false;
            if (jm < n_hits) {
                mine = e.hits[jm];
                mykey = keys[jm];
                myspec = e.spec[jm];
                in_group = (jm == j0) || same_group(keys, e.hits, jm - 1, jm, gbits, is_hash);
            }
            const unsigned brk = __ballot_sync(FULL, !in_group);
            const int cnt = brk ? (__ffs(brk) - 1) : 32;
            if (cnt < 32) group_done = true;
            __syncwarp();
            s_hit[wib][lane] = mine; s_spec[wib][lane] = myspec; s_key[wib][lane] = mykey;
            __syncwarp();
            for (int t = 0; t < cnt; t++) {
                // A run of seeds on the diagonal that was just updated, all starting before its recorded
                // end, is what a long match produces: each of them is a plain `continue` for the reference
                // (s_off_pos < last_hit, no state change), so the whole run is stepped over at once.
                if (is_hash && cache_ok) {
                    const bool covered = lane >= t && lane < cnt && mine.chunk == cur_chunk &&
                                         (int32_t)(mine.s_off - mine.q_off) == cache_diag &&
                                         (int32_t)mine.s_off + ch.diag_offset < cache_level;
                    const unsigned cm = __ballot_sync(FULL, covered) >> t;
                    const int run = (~cm) ? (__ffs(~cm) - 1) : (32 - t);
                    if (run > 0) { t += run - 1; continue; }
                }
                const int64_t j = base + t;
                const SeedHit h = s_hit[wib][t];
                if (h.chunk != cur_chunk) {
                    cur_chunk = h.chunk;
                    ch = e.chunks[cur_chunk];
                    if (is_hash) {
                        // Blast_ExtendWordExit may have reset the container between chunks
                        if (cur_epoch >= 0 && ch.diag_epoch != cur_epoch) {
                            chain.head = 0; chain.used = 0;
                            chain.cells = reinterpret_cast<int4 *>(e.cells) + j;
                            cache_ok = false;
                        }
                        cur_epoch = ch.diag_epoch;
                    }
                }
                int32_t q_off = (int32_t)h.q_off, s_off = (int32_t)h.s_off;
                int32_t s_end = s_off + word;
                const int32_t s_off_pos = s_off + ch.diag_offset;
                int32_t s_end_pos = s_end + ch.diag_offset;
                const int32_t diag = s_off - q_off;
                int32_t last_hit = 0, hit_saved = 0;
                if (is_hash) {
                    if (cache_ok && cache_diag == diag) { last_hit = cache_level; hit_saved = cache_saved; }
                    else if (!chain_get(chain, diag, last_hit, hit_saved)) { last_hit = 0; hit_saved = 0; }
                } else { last_hit = last_hit_cell; hit_saved = flag_cell; }
                if (s_off_pos < last_hit) continue;
                // two-hit mode: a hit far from the previous one on its diagonal is only recorded unless it
                // is a double word by itself
                const bool check_double = two_hits && (hit_saved || s_end_pos > last_hit + window);

                SpecResult r = s_spec[wib][t];
                if (r.status != SPEC_NONE && !check_double) {
                    if (r.status == SPEC_MASKED) continue;
                } else {
                    extend_one(q, e.packed, ch, hit_s_range(e.ranges, ch, (int32_t)h.scan_pos), s_tab, is_hash, has_loc, word,
                               lut, direct, check_double, q_off, s_off, lane, cc, r);
                    if (r.status == SPEC_MASKED) continue;
                }
                q_off = r.q_off; s_off = r.s_off;
                s_end += r.extended; s_end_pos += r.extended;

                int32_t hit_ready = 0;
                if (r.status == SPEC_READY) {
                    hit_ready = 1;
                    const uint64_t kj = s_key[wib][t];
                    if (lane == 0) {
                        const unsigned long long slot = atomicAdd(&e.counters[2], 1ull);
                        if ((int64_t)slot < e.init_capacity) {
                            DevInitHit o;
                            o.chunk = (int32_t)cur_chunk; o.q_off = q_off; o.s_off = s_off;
                            o.q_start = r.q_start; o.s_start = r.s_start; o.length = r.length; o.score = r.score;
                            o.order = (uint32_t)(kj & ((1ull << gbits) - 1ull));
                            e.init[slot] = o;
                        }
                    }
                    s_end_pos = r.length + r.s_start + ch.diag_offset;
                    ++n_extended;
                }
                if (is_hash) {
                    const int32_t d_eff = (r.status == SPEC_SINGLE) ? max(Delta, 0) : Delta;
                    chain_put(chain, diag, s_end_pos, hit_ready ? 0 : s_end_pos - s_off_pos, hit_ready,
                              s_off_pos, window + d_eff + 1, lane);
                    cache_diag = diag; cache_level = s_end_pos; cache_saved = hit_ready; cache_ok = true;
                } else { last_hit_cell = s_end_pos; flag_cell = hit_ready; }
            }
        }
        if (lane == 0 && n_extended) atomicAdd(&e.counters[3], n_extended);
    }
}

// ---- serial replay: two-hit mode with an off-diagonal search (scan_range > 0) -----------------------
// With Delta = MIN(scan_range, window_size - word_length) > 0 a single-word hit looks up the diagonals
// diag +- 1 .. diag +- Delta (core/na_ungapped.c:697-726 array, :853-884 hash), which live in other
// buckets / cells, and the stale-cell recycling of a bucket then depends on every earlier insertion of the
// subject.  There is no independent sub-problem left, so ONE warp replays all hits in the reference's order
// (hits sorted by global scan position) against the complete container: 512 bucket heads in shared memory
// + one cell pool (eDiagHash), or the whole cell array {last_hit, (hit_len << 1) | flag} (eDiagArray).
// The extension itself is still spread over the 32 lanes.  A rarely used option; exactness over speed.
__global__ void __launch_bounds__(32)
extend_serial_kernel(const DevQuery q, const ExtendLaunch e, const uint64_t *keys, int64_t n_hits, int gbits,
                     int32_t diag_array_length)
{
    __shared__ int32_t s_tab[256];
    __shared__ int32_t s_heads[512];
    const int lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) s_tab[i] = q.score_table[i];
    for (int i = lane; i < 512; i += 32) s_heads[i] = 0;
    const bool is_hash = q.container_type == 1;
    int4 *cells = reinterpret_cast<int4 *>(e.cells);          // hash: pool, index 1..
    int2 *arr = reinterpret_cast<int2 *>(e.cells);            // array: one int2 per cell
    const int32_t dmask = diag_array_length - 1;
    if (!is_hash)
        for (int32_t i = lane; i < diag_array_length; i += 32) arr[i] = make_int2(0, 0);
    __syncwarp();

    const int32_t word = q.word_length, lut = q.lut_word_length;
    const bool direct = (word == lut);
    const bool has_loc = q.has_locations && !direct;
    const int32_t window = q.window_size;
    const int32_t Delta = min(q.scan_range, window - word);      // > 0 here
    int32_t used = 0;
    uint32_t cur_chunk = 0xFFFFFFFFu;
    int32_t cur_epoch = -1;
    DevChunk ch{};
    unsigned long long n_extended = 0;
    CtxCache cc{0, -1, 0, 0, 0};

    auto hget = [&](int32_t diag, int32_t &level, int32_t &len, int32_t &saved) -> bool {
        int32_t i = s_heads[diag_bucket(diag)];
        while (i) {
            const int4 v = cells[i];
            if (v.x == diag) { level = v.y; len = v.z >> 1; saved = v.z & 1; return true; }
            i = v.w;
        }
        return false;
    };

    for (int64_t j = 0; j < n_hits; j++) {
        const SeedHit h = e.hits[j];
        if (h.chunk != cur_chunk) {
            cur_chunk = h.chunk;
            ch = e.chunks[cur_chunk];
            if (cur_epoch >= 0 && ch.diag_epoch != cur_epoch) {      // Blast_ExtendWordExit reset (core/blast_extend.c:164-186)
                __syncwarp();
                if (is_hash) { for (int i = lane; i < 512; i += 32) s_heads[i] = 0; used = 0; }
                else for (int32_t i = lane; i < diag_array_length; i += 32) arr[i] = make_int2(-window, 0);
                __syncwarp();
            }
            cur_epoch = ch.diag_epoch;
        }
        int32_t q_off = (int32_t)h.q_off, s_off = (int32_t)h.s_off;
        int32_t s_end = s_off + word;
        const int32_t s_off_pos = s_off + ch.diag_offset;
        int32_t s_end_pos = s_end + ch.diag_offset;
        int32_t diag, real_diag = 0, last_hit = 0, hit_saved = 0, s_l = 0;
        if (is_hash) {
            diag = s_off - q_off;
            if (!hget(diag, last_hit, s_l, hit_saved)) { last_hit = 0; hit_saved = 0; }
        } else {
            diag = s_off + diag_array_length - q_off;
            real_diag = diag & dmask;
            const int2 c = arr[real_diag];
            last_hit = c.x; hit_saved = c.y & 1;
        }
        if (s_off_pos < last_hit) continue;

        const int32_t s_range = hit_s_range(e.ranges, ch, (int32_t)h.scan_pos);
        const uint8_t *S = e.packed + ch.byte_off;
        int32_t extended = 0;
        bool hit_ready = true, off_found = false;
        if (hit_saved || s_end_pos > last_hit + window) {
            const int wt = type_of_word(q, S, q_off, s_off, has_loc, (uint32_t)s_range, word, direct ? word : lut,
                                        true, extended, lane);
            if (!wt) continue;
            s_end += extended; s_end_pos += extended;
            if (wt == 1) {
                // a neighbouring diagonal whose recorded (unsaved) hit ends inside the window makes the pair
                const int32_t s_a = s_off_pos + word - window;
                const int32_t s_b = s_end_pos - 2 * word;
                for (int32_t delta = 1; delta <= Delta && !off_found; ++delta) {
                    int32_t lvl = 0, len = 0, sv = 0;
                    if (is_hash) {
                        if (hget(diag + delta, lvl, len, sv) && len && lvl - delta >= s_a && lvl - len <= s_b) { off_found = true; break; }
                        lvl = len = 0;
                        if (hget(diag - delta, lvl, len, sv) && len && lvl >= s_a && lvl - len + delta <= s_b) { off_found = true; break; }
                    } else {
                        const int32_t orig = real_diag + diag_array_length;
                        int2 c = arr[(orig + delta) & dmask];
                        lvl = c.x; len = c.y >> 1;
                        if (len && lvl - delta >= s_a && lvl - len <= s_b) { off_found = true; break; }
                        c = arr[(orig - delta) & dmask];
                        lvl = c.x; len = c.y >> 1;
                        if (len && lvl >= s_a && lvl - len + delta <= s_b) { off_found = true; break; }
                    }
                }
                if (!off_found) hit_ready = false;
            }
        } else {
            if (!type_of_word(q, S, q_off, s_off, has_loc, (uint32_t)s_range, word, direct ? word : lut, false,
                              extended, lane))
                continue;
            s_end += extended; s_end_pos += extended;
        }

        if (hit_ready) {
            ctx_lookup(q, q_off, lane, cc);
            Ungapped u;
            const int64_t chunk_base = ch.byte_off * 4;
            if (!is_hash && word < 11)
                ungapped_exact(q, e.packed, chunk_base, ch.len, q_off, s_off, -cc.x_dropoff, lane, u);
            else
                ungapped_extend(q, e.packed, chunk_base, ch.len, s_tab, q_off, s_end, s_off, -cc.x_dropoff,
                                cc.reduced_cutoff, lane, u);
            if (off_found || u.score >= cc.cutoff_score) {
                if (lane == 0) {
                    const unsigned long long slot = atomicAdd(&e.counters[2], 1ull);
                    if ((int64_t)slot < e.init_capacity) {
                        DevInitHit o;
                        o.chunk = (int32_t)cur_chunk; o.q_off = q_off; o.s_off = s_off;
                        o.q_start = u.q_start; o.s_start = u.s_start; o.length = u.length; o.score = u.score;
                        o.order = (uint32_t)(keys[j] & ((1ull << gbits) - 1ull));
                        e.init[slot] = o;
                    }
                }
                s_end_pos = u.length + u.s_start + ch.diag_offset;
                ++n_extended;
            } else hit_ready = false;
        }
        const int32_t len = hit_ready ? 0 : s_end_pos - s_off_pos;
        __syncwarp();
        if (is_hash) {
            // s_BlastDiagHashInsert (core/na_ungapped.c:395-450) with the stale horizon window + Delta + 1
            const uint32_t b = diag_bucket(diag);
            const int32_t horizon = window + Delta + 1;
            int32_t i = s_heads[b];
            bool done = false;
            while (i) {
                const int4 v = cells[i];
                if (v.x == diag || s_off_pos - v.y > horizon) {
                    if (lane == 0) cells[i] = make_int4(diag, s_end_pos, (len << 1) | (hit_ready ? 1 : 0), v.w);
                    done = true;
                    break;
                }
                i = v.w;
            }
            if (!done) {
                const int32_t n = ++used;
                if (lane == 0) { cells[n] = make_int4(diag, s_end_pos, (len << 1) | (hit_ready ? 1 : 0), s_heads[b]); }
                __syncwarp();
                if (lane == 0) s_heads[b] = n;
            }
        } else if (lane == 0) {
            arr[real_diag] = make_int2(s_end_pos, ((len & 0xFF) << 1) | (hit_ready ? 1 : 0));   // Uint1 hit_len
        }
        __syncwarp();
    }
    if (lane == 0 && n_extended) atomicAdd(&e.counters[3], n_extended);
}

int64_t extend_serial_cells(int64_t n_hits, bool is_hash, int32_t diag_array_length)
{
    return is_hash ? n_hits + 2 : (int64_t)diag_array_length / 2 + 2;       // int4 units
}

cudaError_t launch_extend_serial(const DevQuery &q, const ExtendLaunch &e, const uint64_t *keys, int64_t n_hits,
                                 int gbits, int32_t diag_array_length, cudaStream_t st)
{
    if (n_hits <= 0) return cudaSuccess;
    extend_serial_kernel<<<1, 32, 0, st>>>(q, e, keys, n_hits, gbits, diag_array_length);
    return cudaGetLastError();
}

// Fast path: hits were grouped on the device (group_sort.cu), their number is only known there; grids
// are sized for the buffer limit and the kernels read the count themselves.
cudaError_t launch_extend_grouped(const DevQuery &q, const ExtendLaunch &e, const uint64_t *keys,
                                  const uint32_t *heads, int gbits, cudaStream_t st)
{
    if (q.window_size <= 0) {
        extend_leaders_kernel<<<SPEC_BLOCKS, EXT_WARPS_PER_BLOCK * 32, 0, st>>>(q, e);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    extend_kernel<<<EXT_BLOCKS, EXT_WARPS_PER_BLOCK * 32, 0, st>>>(q, e, keys, heads, 0, gbits);
    return cudaGetLastError();
}

cudaError_t launch_extend_groups(const DevQuery &q, const ExtendLaunch &e, const uint64_t *keys,
                                 uint32_t *heads, int64_t n_hits, int gbits, cudaStream_t st)
{
    if (n_hits <= 0) return cudaSuccess;
    // two-hit mode: which s_TypeOfWord variant runs depends on the diagonal state, so nothing is
    // extended ahead of the replay there
    const int spec_enabled = q.window_size > 0 ? 0 : 1;
    group_heads_kernel<<<(unsigned)((n_hits + 255) / 256), 256, 0, st>>>(keys, e.hits, n_hits, gbits,
                                                                         q.container_type == 1, spec_enabled, heads,
                                                                         e.leaders, e.spec, e.counters);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    const int64_t want = (n_hits + EXT_WARPS_PER_BLOCK - 1) / EXT_WARPS_PER_BLOCK;
    const unsigned spec_blocks = (unsigned)(want < SPEC_BLOCKS ? (want < 1 ? 1 : want) : SPEC_BLOCKS);
    if (spec_enabled) {
        extend_leaders_kernel<<<spec_blocks, EXT_WARPS_PER_BLOCK * 32, 0, st>>>(q, e);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    const unsigned blocks = (unsigned)(want < EXT_BLOCKS ? (want < 1 ? 1 : want) : EXT_BLOCKS);
    extend_kernel<<<blocks, EXT_WARPS_PER_BLOCK * 32, 0, st>>>(q, e, keys, heads, n_hits, gbits);
    return cudaGetLastError();
}

}  // namespace bn
