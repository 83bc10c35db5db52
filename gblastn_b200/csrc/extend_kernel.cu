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
// one thread replays its group's hits serially, doing the ungapped extension inline (its result
// feeds the next hit's test).  The hash chains are bit-faithful: same prepend-at-head, same
// overwrite-first-stale-cell rule, same carry-over from earlier subjects (BLAST_DiagHash::offset
// grows, core/blast_extend.c:164-186) and the same reset when offset passes INT4_MAX/4.
#include "bn_device.cuh"

namespace bn {

struct Ungapped { int32_t q_start, s_start, length, score; };

// s_NuclUngappedExtendExact
__device__ void ungapped_exact(const DevQuery &q, const uint8_t *S, int32_t slen, int32_t q_off,
                               int32_t s_off, int32_t X, Ungapped &u)
{
    const uint8_t *query = q.query;
    int32_t sum = 0, score = 0;
    int32_t qp = q_off, q_beg = q_off, q_end = q_off;
    const int32_t q_avail = q.concat_len - q_off, s_avail = slen - s_off;
    const int32_t s_lo = (q_off < s_off) ? s_off - q_off : 0;
    int32_t sp = s_off;
    while (sp > s_lo) {
        --sp; --qp;
        sum += __ldg(&q.matrix[16 * (int)__ldg(query + qp) + sbase(S, sp)]);
        if (sum > 0) { q_beg = qp; score += sum; sum = 0; }
        else if (sum < X) break;
    }
    u.q_start = q_beg;
    u.s_start = s_off - (q_off - q_beg);
    const int32_t s_hi = (q_avail < s_avail) ? s_off + q_avail : slen;
    qp = q_off; sp = s_off; sum = 0;
    while (sp < s_hi) {
        sum += __ldg(&q.matrix[16 * (int)__ldg(query + qp) + sbase(S, sp)]);
        ++qp; ++sp;
        if (sum > 0) { q_end = qp; score += sum; sum = 0; }
        else if (sum < X) break;
    }
    u.length = q_end - q_beg;
    u.score = score;
}

__device__ __forceinline__ uint32_t qbyte(const uint8_t *query, int32_t p)
{
    // (q[0] << 6) | (q[1] << 4) | (q[2] << 2) | q[3], truncated to 8 bits like the reference's Uint1
    return (((uint32_t)__ldg(query + p) << 6) | ((uint32_t)__ldg(query + p + 1) << 4) |
            ((uint32_t)__ldg(query + p + 2) << 2) | (uint32_t)__ldg(query + p + 3)) & 0xFFu;
}

// s_NuclUngappedExtend
__device__ void ungapped_extend(const DevQuery &q, const uint8_t *S, int32_t slen, int32_t q_off,
                                int32_t s_match_end, int32_t s_off, int32_t X, int32_t reduced_cutoff,
                                Ungapped &u)
{
    const uint8_t *query = q.query;
    int32_t len = (4 - (s_off % 4)) % 4;
    const int32_t q_ext = q_off + len, s_ext = s_off + len;
    int32_t qp = q_ext, sb = s_ext / 4;
    int32_t sum = 0, score = 0, new_q = q_ext;

    len = min(q_ext, s_ext) / 4;
    for (int32_t i = 0; i < len; --sb, qp -= 4, ++i) {
        sum += __ldg(&q.score_table[qbyte(query, qp - 4) ^ (uint32_t)__ldg(S + sb - 1)]);
        if (sum > 0) { new_q = qp - 4; score += sum; sum = 0; }
        if (sum < X) break;
    }
    u.q_start = new_q;
    u.s_start = s_ext - (q_ext - new_q);

    qp = q_ext; sb = s_ext / 4;
    len = min(q.concat_len - q_ext, slen - s_ext) / 4;
    sum = 0; new_q = qp;
    for (int32_t i = 0; i < len; ++sb, qp += 4, ++i) {
        sum += __ldg(&q.score_table[qbyte(query, qp) ^ (uint32_t)__ldg(S + sb)]);
        if (sum > 0) { new_q = qp + 3; score += sum; sum = 0; }
        if (sum < X) break;
    }
    if (score >= reduced_cutoff) {
        ungapped_exact(q, S, slen, q_off, s_off, X, u);
    } else {
        u.score = score;
        u.length = max(s_match_end - u.s_start, new_q - u.q_start + 1);
    }
}

// s_MBLookup / s_SmallNaLookup
__device__ bool lut_contains(const DevQuery &q, uint32_t index, int32_t q_pos)
{
    if (q.lut_type == 0) {
        int32_t v = __ldg(&q.hashtable[index & q.hash_mask]);
        ++q_pos;
        while (v) {
            if (v == q_pos) return true;
            v = __ldg(&q.next_pos[v]);
        }
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

// s_IsSeedMasked
__device__ __forceinline__ bool seed_masked(const DevQuery &q, const uint8_t *S, int32_t s_off,
                                            int32_t lut, int32_t q_pos)
{
    uint32_t w = be32(S + s_off / 4);
    return !lut_contains(q, w >> (2 * (16 - s_off % 4 - lut)), q_pos);
}

// s_TypeOfWord with check_double == FALSE (window_size == 0)
__device__ int type_of_word(const DevQuery &q, const uint8_t *S, int32_t &q_off, int32_t &s_off,
                            bool has_locations, uint32_t s_range, int32_t word_length, int32_t lut,
                            int32_t &extended)
{
    extended = 0;
    if (word_length == lut) return 1;
    int32_t q_end = q_off + word_length, s_end = s_off + word_length;
    const int32_t context = ctx_search(q, q_end);
    const int32_t q_range = __ldg(&q.ctx[context].query_offset) + __ldg(&q.ctx[context].query_length);
    if (has_locations) {
        if (seed_masked(q, S, s_end - lut, lut, q_end - lut)) return 0;
        for (;; ++s_off, ++q_off)
            if (!seed_masked(q, S, s_off, lut, q_off)) break;
    }
    const int32_t ext_to = word_length - (q_end - q_off);
    const uint32_t a = (uint32_t)(q_range - q_end), c = s_range - (uint32_t)s_end;
    const int32_t ext_max = (int32_t)(a > c ? c : a);   // unsigned MIN, as in the reference (:534)
    if (ext_to || has_locations) {
        if (ext_to > ext_max) return 0;
        q_end += ext_to; s_end += ext_to;
        for (int32_t s_pos = s_end - lut, q_pos = q_end - lut; s_pos > s_off; s_pos -= lut, q_pos -= lut)
            if (seed_masked(q, S, s_pos, lut, q_pos)) return 0;
        extended = ext_to;
    }
    return 1;
}

// ---- bucket chain (BLAST_DiagHash restricted to one bucket) -------------------------------------
// cell = int4 {diag, level, (hit_len << 1) | hit_saved, next}; index 0 = null.
struct Chain {
    int4 *cells;        // cells[1..]: storage region of this group
    int32_t head;       // backbone[bucket]
    int32_t used;       // cells allocated so far in this region
};

__device__ __forceinline__ bool chain_get(const Chain &c, int32_t diag, int32_t &level)
{
    int32_t i = c.head;
    while (i) {
        int4 v = c.cells[i];
        if (v.x == diag) { level = v.y; return true; }
        i = v.w;
    }
    return false;
}

__device__ __forceinline__ void chain_put(Chain &c, int32_t diag, int32_t level, int32_t len,
                                          int32_t saved, int32_t s_off_pos, int32_t window)
{
    int32_t i = c.head;
    while (i) {
        int4 v = c.cells[i];
        if (v.x == diag || s_off_pos - v.y > window) {
            c.cells[i] = make_int4(diag, level, (len << 1) | saved, v.w);
            return;
        }
        i = v.w;
    }
    const int32_t n = ++c.used;
    c.cells[n] = make_int4(diag, level, (len << 1) | saved, c.head);
    c.head = n;
}

__device__ __forceinline__ uint32_t diag_bucket(int32_t diag)
{
    return ((uint32_t)diag * 0x9E370001u) % 512u;
}

// One thread per sorted hit; only the first hit of every group does work and replays the group.
__global__ void __launch_bounds__(128)
extend_kernel(const DevQuery q, const ExtendLaunch e, const uint64_t *group_key, int64_t n_hits)
{
    const int64_t j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j0 >= n_hits) return;
    const uint64_t gk = group_key[j0];
    if (j0 > 0 && group_key[j0 - 1] == gk) return;        // not a group head

    const bool is_hash = q.container_type == 1;
    const int32_t word = q.word_length, lut = q.lut_word_length;
    const bool direct = (word == lut);
    const bool has_loc = q.has_locations && !direct;
    // window_size == 0  =>  Delta = MIN(scan_range, -word_length); staleness window = Delta + 1
    const int32_t stale_window = min(q.scan_range, -word) + 1;

    Chain chain;
    chain.cells = reinterpret_cast<int4 *>(e.cells) + j0;   // region [j0+1 .. j0+group_size]
    chain.head = 0; chain.used = 0;
    int32_t last_hit_cell = 0;      // eDiagArray: the cell's last_hit
    uint32_t cur_chunk = 0xFFFFFFFFu;
    int32_t cur_epoch = -1;
    DevChunk ch;
    const uint8_t *S = nullptr;
    unsigned long long n_extended = 0;

    for (int64_t j = j0; j < n_hits && group_key[j] == gk; ++j) {
        const SeedHit h = e.hits[j];
        if (h.chunk != cur_chunk) {
            cur_chunk = h.chunk;
            ch = e.chunks[cur_chunk];
            S = e.packed + ch.byte_off;
            if (is_hash) {
                // Blast_ExtendWordExit resets happen between chunks; replay every reset that
                // occurred since the previous chunk this bucket saw (any one empties the chain).
                if (cur_epoch >= 0 && ch.diag_epoch != cur_epoch) {
                    chain.head = 0; chain.used = 0;
                    chain.cells = reinterpret_cast<int4 *>(e.cells) + j;
                }
                cur_epoch = ch.diag_epoch;
            }
        }
        int32_t q_off = (int32_t)h.q_off, s_off = (int32_t)h.s_off;
        const int32_t s_range = ch.len;
        int32_t s_end = s_off + word;
        const int32_t s_off_pos = s_off + ch.diag_offset;
        int32_t s_end_pos = s_end + ch.diag_offset;
        const int32_t diag = s_off - q_off;
        int32_t last_hit = 0;
        if (is_hash) { if (!chain_get(chain, diag, last_hit)) last_hit = 0; }
        else last_hit = last_hit_cell;
        if (s_off_pos < last_hit) continue;

        int32_t extended = 0;
        if (!type_of_word(q, S, q_off, s_off, has_loc, (uint32_t)s_range, word, direct ? word : lut, extended))
            continue;
        s_end += extended; s_end_pos += extended;

        const int32_t context = ctx_search(q, q_off);
        const DevContext c = q.ctx[context];
        Ungapped u;
        if (!is_hash && word < 11)
            ungapped_exact(q, S, ch.len, q_off, s_off, -c.x_dropoff, u);
        else
            ungapped_extend(q, S, ch.len, q_off, s_end, s_off, -c.x_dropoff, c.reduced_cutoff, u);

        int32_t hit_ready = 0;
        if (u.score >= c.cutoff_score) {
            hit_ready = 1;
            unsigned long long slot = atomicAdd(&e.counters[2], 1ull);
            if ((int64_t)slot < e.init_capacity) {
                DevInitHit o;
                o.chunk = (int32_t)cur_chunk; o.q_off = q_off; o.s_off = s_off;
                o.q_start = u.q_start; o.s_start = u.s_start; o.length = u.length; o.score = u.score;
                o.order = e.order[j];
                e.init[slot] = o;
            }
            s_end_pos = u.length + u.s_start + ch.diag_offset;
            ++n_extended;
        }
        if (is_hash)
            chain_put(chain, diag, s_end_pos, hit_ready ? 0 : s_end_pos - s_off_pos, hit_ready,
                      s_off_pos, stale_window);
        else
            last_hit_cell = s_end_pos;
    }
    if (n_extended) atomicAdd(&e.counters[3], n_extended);
}

// group key of every sorted hit: hash -> bucket id; array -> (chunk, real diagonal)
__global__ void group_key_kernel(const DevQuery q, const SeedHit *hits, const uint32_t *perm,
                                 int64_t n, int32_t diag_array_length, uint64_t *keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SeedHit h = hits[perm[i]];
    if (q.container_type == 1) {
        keys[i] = diag_bucket((int32_t)h.s_off - (int32_t)h.q_off);
    } else {
        uint32_t real = (uint32_t)((int32_t)h.s_off + diag_array_length - (int32_t)h.q_off) &
                        (uint32_t)(diag_array_length - 1);
        keys[i] = ((uint64_t)h.chunk << 32) | real;
    }
}

__global__ void gather_hits_kernel(const SeedHit *in, const uint32_t *perm, int64_t n, SeedHit *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

__global__ void iota_kernel(uint32_t *p, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

cudaError_t launch_group_keys(const DevQuery &q, const SeedHit *hits, const uint32_t *perm, int64_t n,
                              int32_t diag_array_length, uint64_t *keys, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    group_key_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, hits, perm, n, diag_array_length, keys);
    return cudaGetLastError();
}
cudaError_t launch_gather_hits(const SeedHit *in, const uint32_t *perm, int64_t n, SeedHit *out,
                               cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    gather_hits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, perm, n, out);
    return cudaGetLastError();
}
cudaError_t launch_iota(uint32_t *p, int64_t n, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n);
    return cudaGetLastError();
}

cudaError_t launch_extend_groups(const DevQuery &q, const ExtendLaunch &e, const uint64_t *group_key,
                                 int64_t n_hits, cudaStream_t st)
{
    if (n_hits <= 0) return cudaSuccess;
    extend_kernel<<<(unsigned)((n_hits + 127) / 128), 128, 0, st>>>(q, e, group_key, n_hits);
    return cudaGetLastError();
}

}  // namespace bn
