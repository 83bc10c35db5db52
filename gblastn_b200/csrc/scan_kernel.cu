// scan_kernel.cu — stage 1 of the blastn hot path on the GPU:
//   subject scan  +  lookup-chain expansion  +  mini-extension to the full word.
//
// Replaces (semantics, not code):
//   scanners      core/blast_nascan.c:1489-1591 (s_MBScanSubject_Any and its 30 specialisations,
//                 which all visit p = 0, step, 2*step ... <= len - lut and differ only in how they
//                 fetch the word), :445-560 (small-table scanners)
//   chain walk    s_BlastMBLookupRetrieve core/blast_nascan.c:1413-1427,
//                 s_BlastSmallNaRetrieveHits :312-335
//   mini-ext.     s_BlastNaExtend core/na_ungapped.c:1026-1148 (== ...Aligned :1166-1290 on aligned
//                 hits), s_BlastNaExtendDirect :942-1005, s_BlastSmallNaExtendAlignedOneByte
//                 :1347-1427, s_BlastSmallNaExtend :1450-1555
//
// One launch covers every chunk of a resident volume.  A block owns POS_PER_BLOCK consecutive scan
// positions of the volume-wide position space (prefix sums in the chunk table).
//
//   phase 0  the block's slice of the packed subject (a few KB, contiguous in the volume) is staged
//            into shared memory by ONE TMA bulk copy (cp.async.bulk + mbarrier); meanwhile the block
//            builds a shared-memory table of the chunks it spans, so position -> (chunk, offset) is a
//            32-bit cursor walk instead of a 64-bit binary search in global memory
//   phase A  every thread forms the lookup words of its 8 positions from the tile and probes the
//            compact table word {presence bits, rank} (one 8-byte L2 access per position, all 8 in
//            flight together); occupied cells are pushed to a shared-memory candidate queue with one
//            reservation per warp
//   phase B  candidates are processed DENSELY (one per thread, no idle lanes): cinfo[rank] holds the
//            first chain element together with the query's 16 bases on either side of the word, so
//            the common case is one 16-byte load + two tile windows; chains continue through
//            next_pos / qinfo; survivors are appended with warp-aggregated atomics
//
// A survivor carries the 64-bit key (group << gbits | global position): group = diagonal-hash bucket
// or diagonal-array cell.  One stable radix sort on that key both groups the hits for the diagonal
// stage and restores the reference's emission order inside a group (a position is handled by one
// thread, which emits its chain in chain order).
// The generic kernel (direct global loads, full hashtable) is kept for small tables and for blocks
// whose byte span does not fit the tile.
#include "bn_device.cuh"
#include <cstdlib>

namespace bn {

#ifndef BN_SCAN_THREADS
#define BN_SCAN_THREADS 256
#endif
#ifndef BN_POS_PER_THREAD
#define BN_POS_PER_THREAD 8
#endif
#ifndef BN_SCAN_MIN_BLOCKS
#define BN_SCAN_MIN_BLOCKS 6
#endif
constexpr int SCAN_THREADS = BN_SCAN_THREADS;
constexpr int POS_PER_THREAD = BN_POS_PER_THREAD;
constexpr int POS_PER_BLOCK = SCAN_THREADS * POS_PER_THREAD;
constexpr int TILE_BYTES = 20 * 1024;      // staged subject slice (incl. 64-byte margins)
constexpr int TILE_MARGIN = 64;
static_assert(POS_PER_BLOCK <= 2048, "candidate packing holds 11 bits of in-block position");

int scan_positions_per_block() { return POS_PER_BLOCK; }

// bytes of shared memory for the staged subject slice of one block (multiple of 16, <= TILE_BYTES)
int scan_tile_cap(int scan_step, int word_length)
{
    long need = (long)POS_PER_BLOCK * scan_step / 4 + word_length / 4 + 2 * TILE_MARGIN + 64;
    need = (need + 15) & ~15L;
    if (need < 1024) need = 1024;
    return (int)(need > TILE_BYTES ? TILE_BYTES : need);
}

__device__ __forceinline__ uint32_t load_window(const uint8_t *packed, int64_t byte)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(packed + (byte & ~int64_t(3)));
    uint32_t a = __byte_perm(__ldg(w), 0, 0x0123);       // big-endian view of the first word
    uint32_t b = __byte_perm(__ldg(w + 1), 0, 0x0123);
    return __funnelshift_l(b, a, (uint32_t)(byte & 3) * 8);
}

__device__ __forceinline__ uint32_t diag_group(const DevQuery &q, const ScanLaunch &s, int32_t q_off, int32_t s_off)
{
    if (s.raw_pairs || s.one_group) return 0;
    if (q.container_type == 1) return ((uint32_t)(s_off - q_off) * 0x9E370001u) % 512u;       // hash bucket
    return (uint32_t)(s_off + s.diag_array_length - q_off) & (uint32_t)(s.diag_array_length - 1);  // array cell
}

__device__ __forceinline__ void emit_hit(const DevQuery &q, const ScanLaunch &s, uint32_t chunk, uint32_t p,
                                         int64_t g, int32_t q_off, int32_t s_off)
{
    // warp-aggregated append
    unsigned mask = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&s.counters[0], (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    unsigned long long slot = base + __popc(mask & ((1u << lane) - 1));
    if ((int64_t)slot < s.capacity) {
        SeedHit h;
        h.chunk = chunk; h.scan_pos = p; h.q_off = (uint32_t)q_off; h.s_off = (uint32_t)s_off;
        const uint32_t grp = diag_group(q, s, q_off, s_off);
        s.hits[slot] = h;
        s.keys[slot] = ((uint64_t)grp << s.gbits) | (uint64_t)g;
        if (s.bucket_count) {       // device-side grouping: the count is also the hit's slot in its bucket's region
            const uint32_t pos = atomicAdd(&s.bucket_count[grp], 1u);
            if (pos < (uint32_t)s.bucket_cap)
                s.bucket_keys[(size_t)grp * (size_t)s.bucket_cap + pos] = ((uint64_t)g << 24) | (slot & 0xFFFFFFull);
        }
    }
}

// s_BlastNaExtend on one (q_offset, s_offset) pair; returns true and the shifted offsets when the
// full word is an exact match.
__device__ __forceinline__ bool mini_extend_mb(const DevQuery &q, const uint8_t *S, int32_t s_range,
                                               int32_t q_offset, int32_t s_offset, int32_t &q_out,
                                               int32_t &s_out)
{
    const int32_t lut = q.lut_word_length, ext_to = q.word_length - lut;
    int32_t ext_left = 0;
    if (ext_to > 0) {
        int32_t lim = min(ext_to, s_offset);
        int32_t sp = s_offset, qp = q_offset;
        for (; ext_left < lim; ++ext_left) {
            --sp; --qp;
            if (sbase(S, sp) != (int)__ldg(q.query + qp)) break;
        }
        if (ext_left < ext_to) {
            int32_t ext_right = 0, need = ext_to - ext_left;
            sp = s_offset + lut;
            if ((uint32_t)(sp + need) > (uint32_t)s_range) return false;
            qp = q_offset + lut;
            for (; ext_right < need; ++ext_right) {
                if (sbase(S, sp) != (int)__ldg(q.query + qp)) break;
                ++sp; ++qp;
            }
            if (ext_right < need) return false;
        }
    }
    q_out = q_offset - ext_left;
    s_out = s_offset - ext_left;
    return true;
}

// compressed_nuc_seq[i] of BlastCompressBlastnaSequence (core/blast_util.c:459-501), recomputed.
__device__ __forceinline__ uint32_t cq(const DevQuery &q, int32_t i)
{
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int32_t p = i + k;
        v <<= 2;
        if (p >= 0 && p < q.concat_len) v |= (uint32_t)(__ldg(q.query + p) & 3);
    }
    return v;
}
__device__ __forceinline__ int match_left(uint32_t x)   // s_ExactMatchExtendLeft
{
    x &= 0xFF;
    if (x == 0) return 4;
    return (__ffs(x) - 1) >> 1;
}
__device__ __forceinline__ int match_right(uint32_t x)  // s_ExactMatchExtendRight
{
    x &= 0xFF;
    if (x == 0) return 4;
    return (__clz(x) - 24) >> 1;
}

__device__ __forceinline__ bool mini_extend_small(const DevQuery &q, const uint8_t *S, int32_t s_range,
                                                  int32_t q_offset, int32_t s_offset, int32_t &q_out,
                                                  int32_t &s_out)
{
    const int32_t word = q.word_length, lut = q.lut_word_length, ext_to = word - lut;
    int32_t ext_left = 0, ext_right = 0;
    if (ext_to == 0) { q_out = q_offset; s_out = s_offset; return true; }
    int32_t context = ctx_search(q, q_offset);
    int32_t q_start = __ldg(&q.ctx[context].query_offset);
    int32_t q_range = q_start + __ldg(&q.ctx[context].query_length);

    if (lut % 4 == 0 && q.scan_step % 4 == 0 && ext_to <= 4) {
        if (s_offset > 0 && q_offset > 0) {
            ext_left = match_left(cq(q, q_offset - 4) ^ (uint32_t)__ldg(S + s_offset / 4 - 1));
            ext_left = min(min(ext_left, ext_to), q_offset - q_start);
        }
        if (ext_left < ext_to && (q_offset + lut) < q.concat_len) {
            ext_right = match_right(cq(q, q_offset + lut) ^ (uint32_t)__ldg(S + (s_offset + lut) / 4));
            ext_right = min(min(ext_right, s_range - (s_offset + lut)), q_range - (q_offset + lut));
            if (ext_left + ext_right < ext_to) return false;
        }
    } else {
        int32_t ext_max = min(min(ext_to, s_offset), q_offset - q_start);
        int32_t rsdl = 4 - (s_offset % 4);
        s_offset += rsdl; q_offset += rsdl; ext_max += rsdl;
        int32_t s_off = s_offset, q_off = q_offset;
        while (ext_left < ext_max) {
            int bases = match_left(cq(q, q_off - 4) ^ (uint32_t)__ldg(S + s_off / 4 - 1));
            ext_left += bases;
            if (bases < 4) break;
            q_off -= 4; s_off -= 4;
        }
        ext_left = min(ext_left, ext_max);
        s_off = s_offset; q_off = q_offset;
        ext_max = min(min(word - ext_left, s_range - s_off), q_range - q_off);
        while (ext_right < ext_max) {
            int bases = match_right(cq(q, q_off) ^ (uint32_t)__ldg(S + s_off / 4));
            ext_right += bases;
            if (bases < 4) break;
            q_off += 4; s_off += 4;
        }
        ext_right = min(ext_right, ext_max);
        if (ext_left + ext_right < word) return false;
    }
    q_out = q_offset - ext_left;
    s_out = s_offset - ext_left;
    return true;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const DevQuery q, const ScanLaunch s)
{
    __shared__ int32_t sh_first_chunk, sh_last_chunk;
    const int64_t block_pos0 = (int64_t)blockIdx.x * POS_PER_BLOCK;
    if (threadIdx.x == 0) {
        sh_first_chunk = s.block_chunk[blockIdx.x];
        sh_last_chunk = s.block_chunk[blockIdx.x + 1];
    }
    __syncthreads();
    const int32_t c_lo = sh_first_chunk, c_hi = sh_last_chunk;
    const int32_t lut = q.lut_word_length, step = q.scan_step;
    unsigned long long my_lookup_hits = 0;

#pragma unroll 1
    for (int it = 0; it < POS_PER_THREAD; it++) {
        const int64_t g = block_pos0 + (int64_t)it * SCAN_THREADS + threadIdx.x;
        if (g >= s.total_pos) break;
        // chunk that owns position g: last chunk in [c_lo, c_hi] with pos_prefix <= g
        int32_t lo = c_lo, hi = c_hi;
        while (lo < hi) {
            int32_t m = (lo + hi + 1) >> 1;
            if (__ldg(&s.chunks[m].pos_prefix) <= g) lo = m; else hi = m - 1;
        }
        const DevChunk ch = s.chunks[lo];
        const int32_t p = ch.p_first + (int32_t)(g - ch.pos_prefix) * step;
        const uint8_t *S = s.packed + ch.byte_off;
        const uint32_t window = load_window(s.packed, ch.byte_off + (p >> 2));
        const uint32_t idx = (window >> (2 * (16 - ((p & 3) + lut)))) & q.hash_mask;
        lo = ch.parent;                       // seed hits record the chunk, not the scan unit

        if (q.lut_type == 0) {
            int32_t qp = mb_cell(q, idx);
            while (qp) {
                ++my_lookup_hits;
                int32_t qo, so;
                if (s.raw_pairs) emit_hit(q, s, (uint32_t)lo, (uint32_t)p, g, qp - 1, p);
                else if (mini_extend_mb(q, S, ch.s_range, qp - 1, p, qo, so))
                    emit_hit(q, s, (uint32_t)lo, (uint32_t)p, g, qo, so);
                qp = __ldg(&q.next_pos[qp]);
            }
        } else if (q.lut_type == 2) {
            // s_BlastLookupRetrieve core/blast_nascan.c:63-85; extension as for the megablast table
            const int4 cell = __ldg(&q.na_cells[idx]);
            const int32_t nh = cell.x;
            for (int32_t i = 0; i < nh; i++) {
                const int32_t v = nh <= 3 ? (i == 0 ? cell.y : (i == 1 ? cell.z : cell.w)) : __ldg(&q.na_overflow[cell.y + i]);
                ++my_lookup_hits;
                int32_t qo, so;
                if (s.raw_pairs) emit_hit(q, s, (uint32_t)lo, (uint32_t)p, g, v, p);
                else if (mini_extend_mb(q, S, ch.s_range, v, p, qo, so))
                    emit_hit(q, s, (uint32_t)lo, (uint32_t)p, g, qo, so);
            }
        } else {
            int32_t v = __ldg(&q.backbone[idx]);
            if (v == -1) continue;
            int32_t src = 0;
            if (v < 0) { src = -v; v = __ldg(&q.overflow[src++]); }
            do {
                ++my_lookup_hits;
                int32_t qo, so;
                if (s.raw_pairs) emit_hit(q, s, (uint32_t)lo, (uint32_t)p, g, v, p);
                else if (mini_extend_small(q, S, ch.s_range, v, p, qo, so))
                    emit_hit(q, s, (uint32_t)lo, (uint32_t)p, g, qo, so);
                v = src ? (int32_t)__ldg(&q.overflow[src++]) : -1;
            } while (v >= 0);
        }
    }
    // one atomic per warp for the lookup-hit statistic (BlastUngappedStats.lookup_hits)
    for (int o = 16; o > 0; o >>= 1) my_lookup_hits += __shfl_down_sync(0xffffffffu, my_lookup_hits, o);
    if ((threadIdx.x & 31) == 0 && my_lookup_hits) atomicAdd(&s.counters[1], my_lookup_hits);
}

// ---- staged kernel (megablast tables) ------------------------------------------------------------
// candidate = occupied table cell met at a scan position: {rank of the cell, chunk delta << 11 | position in block}
// 16-base window of the staged tile starting at tile-relative base position tb (>= 0)
__device__ __forceinline__ uint32_t tile_win(const uint32_t *tile, int32_t tb)
{
    const uint32_t a = __byte_perm(tile[tb >> 4], 0, 0x0123);
    const uint32_t b = __byte_perm(tile[(tb >> 4) + 1], 0, 0x0123);
    return __funnelshift_l(b, a, (uint32_t)(tb & 15) * 2);
}


// Lookup words of 8 CONSECUTIVE scan positions of one chunk (compile-time stride STEP bases), first position at
// tile-relative base tb_first: the 10 words that hold them are loaded once (lanes 8 * STEP bases apart: for STEP 17
// and 18 an odd number of words, no bank conflict), aligned to the first position with 9 funnel shifts, and every
// word is then one more shift by a compile-time amount — ~6 instructions per position where the per-position
// form (two LDS, selector arithmetic, PRMT, two shifts on run-time amounts) takes ~25.
template <int STEP>
__device__ __forceinline__ void words8(const uint32_t *tile, int32_t tb_first, uint32_t shr, uint32_t (&idx)[8])
{
    static_assert(STEP >= 16 && 14 * STEP + 26 + 30 <= 320 && (14 * STEP) / 32 + 1 <= 8, "ten raw words cover the span");
    const uint32_t *w = tile + (tb_first >> 4);
    uint32_t R[10], N[9];
#pragma unroll
    for (int i = 0; i < 10; i++) R[i] = __byte_perm(w[i], 0, 0x0123);
    const uint32_t off0 = ((uint32_t)tb_first & 15u) * 2u;
#pragma unroll
    for (int i = 0; i < 9; i++) N[i] = __funnelshift_l(R[i + 1], R[i], off0);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int bit = j * 2 * STEP, a = bit >> 5, sh = bit & 31;
        const uint32_t v = sh ? __funnelshift_l(N[a + 1 < 9 ? a + 1 : 8], N[a], (uint32_t)sh) : N[a];
        idx[j] = v >> shr;
    }
}

// Dense-signature probe (DevQuery::psig): what a scan position contributes is its table cell and the 14 bits its psig
// entry is compared with.  v0 = 16 bases from 4 bases in front of the position, v1 = the 16 after them.
__device__ __forceinline__ void probe_key(uint32_t v0, uint32_t v1, uint32_t shr, uint32_t &idx, uint32_t &want)
{
    const uint32_t v = __funnelshift_l(v1, v0, 8);           // 16 bases from the position on
    idx = v >> shr;
    want = ((v0 >> 16) & 0x3F00u) | ((v >> (shr - 8u)) & 0xFFu);     // 3 bases left of the word | 4 bases right of it (lut <= 12)
}
// occupied cell whose first chain element cannot reach the full word: fewer than 3 matching bases on its left and fewer
// than 4 on its right, no further chain elements (see scan_candidate)
__device__ __forceinline__ bool psig_pass(uint32_t ps, uint32_t want)
{
    const uint32_t x = ps ^ want;
    return (ps & 0x4000u) || !(x & 0x3F00u) || !(x & 0xFFu);
}
// words8 for the dense-signature probe: the stream is aligned 4 bases in front of the first position and one raw word
// longer (11 words: position 7 of stride 18 ends at bit 252 + 64)
template <int STEP>
__device__ __forceinline__ void words8f(const uint32_t *tile, int32_t tb_first, uint32_t shr, uint32_t (&idx)[8], uint32_t (&want)[8])
{
    static_assert(STEP >= 16 && 14 * STEP + 64 <= 320, "eleven raw words cover the span");
    const int32_t t0 = tb_first - 4;
    const uint32_t *w = tile + (t0 >> 4);
    uint32_t R[11], N[10];
#pragma unroll
    for (int i = 0; i < 11; i++) R[i] = __byte_perm(w[i], 0, 0x0123);
    const uint32_t off0 = ((uint32_t)t0 & 15u) * 2u;
#pragma unroll
    for (int i = 0; i < 10; i++) N[i] = __funnelshift_l(R[i + 1], R[i], off0);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int bit = j * 2 * STEP, a = bit >> 5, sh = bit & 31;
        const uint32_t v0 = sh ? __funnelshift_l(N[a + 1], N[a], (uint32_t)sh) : N[a];
        const uint32_t v1 = sh ? __funnelshift_l(N[a + 2 < 10 ? a + 2 : 9], N[a + 1], (uint32_t)sh) : N[a + 1];
        probe_key(v0, v1, shr, idx[j], want[j]);
    }
}

// s_BlastNaExtend on 16-base windows.  The query's 16 bases on either side of the lookup word come
// with the chain element (qinfo), so the common case needs no further query access; tbase =
// tile-relative base index of the chunk's base 0.
__device__ __forceinline__ bool mini_extend_tile(const DevQuery &q, const uint32_t *tile, int32_t tbase,
                                                 int32_t s_range, int32_t q_offset, int32_t s_offset,
                                                 const uint4 qi, int32_t &q_out, int32_t &s_out)
{
    const int32_t lut = q.lut_word_length, ext_to = q.word_length - lut;
    int32_t ext_left = 0;
    if (ext_to > 0) {
        const int32_t lim = min(ext_to, s_offset);
        if (lim > 0) {
            const uint32_t m = mismatch_bits(qi.y, qi.w & 0x55555555u, tile_win(tile, tbase + s_offset - 16));
            ext_left = m ? ((__ffs(m) - 1) >> 1) : 16;
            if (ext_left >= lim) ext_left = lim;
            else if (!m) {                       // all 16 matched and more are wanted: continue on windows
                while (ext_left < lim) {
                    uint32_t qb, qa;
                    qwin(q, q_offset - ext_left - 16, qb, qa);
                    const uint32_t mm = mismatch_bits(qb, qa, tile_win(tile, tbase + s_offset - ext_left - 16));
                    if (mm) { ext_left = min(ext_left + ((__ffs(mm) - 1) >> 1), lim); break; }
                    ext_left = min(ext_left + 16, lim);
                }
            }
        }
        if (ext_left < ext_to) {
            const int32_t need = ext_to - ext_left;
            const int32_t sp = s_offset + lut;
            if ((uint32_t)(sp + need) > (uint32_t)s_range) return false;
            const uint32_t m = mismatch_bits(qi.z, (qi.w >> 1) & 0x55555555u, tile_win(tile, tbase + sp));
            int32_t ext_right = m ? (__clz(m) >> 1) : 16;
            if (ext_right < need) {
                if (m) return false;
                while (ext_right < need) {       // need > 16
                    uint32_t qb, qa;
                    qwin(q, q_offset + lut + ext_right, qb, qa);
                    const uint32_t mm = mismatch_bits(qb, qa, tile_win(tile, tbase + sp + ext_right));
                    if (mm) { ext_right = min(ext_right + (__clz(mm) >> 1), need); break; }
                    ext_right = min(ext_right + 16, need);
                }
                if (ext_right < need) return false;
            }
        }
    }
    q_out = q_offset - ext_left;
    s_out = s_offset - ext_left;
    return true;
}


// ---- direct filter (blastn mode: lookup word == full word, one-hit) ---------------------------------
// With lut == word every lookup hit goes straight to the diagonal logic and — unless an earlier extension on its
// diagonal covers it — to s_NuclUngappedExtend (core/na_ungapped.c:263-350): at stride 1 that is one extension per
// ~2 subject bases (C3: 4.8e8 for 1 Gb), of which ~1 % reach cutoff_score.  A hit whose extension scores below the
// cutoff leaves nothing behind but last_hit = s_off + word on its diagonal (:888-915), which can only reject hits of
// the same exact-match run, and those score the same; so hits that certainly fail are dropped right here, one
// THREAD per hit on 16-base windows of the staged tile, and only the rest goes through sort + replay (where the
// extension is recomputed bit-exactly).  The test below is the reference's own decision — approximate pass
// (4 bases per step, nucl_score_table) >= reduced cutoff, then exact pass >= cutoff_score — evaluated exactly when
// every base it reads is unambiguous; anything else (ambiguity codes, sentinels) is kept.
struct DirectCtx {
    const uint32_t *tile;      // staged slice
    int32_t tile_bases;        // bases in the slice
    int64_t tile_abs;          // absolute base index (4 * volume byte) of the slice's first base
    const uint8_t *packed;
};
__device__ __forceinline__ uint32_t direct_swin(const DirectCtx &d, int32_t tb)
{
    if (tb >= 0 && tb + 32 <= d.tile_bases) return tile_win(d.tile, tb);
    return swin(d.packed, d.tile_abs + tb);
}

// true = keep (may reach the cutoff); false = the reference's extension certainly stays below cutoff_score
__device__ __noinline__ bool direct_keep(const DevQuery &q, const DirectCtx &d, int32_t tbase, int32_t slen,
                                         int32_t q_off, int32_t s_off, int32_t x_dropoff, int32_t cutoff, int32_t reduced)
{
    const int32_t X = -x_dropoff;
    const int32_t r = q.reward, pen = q.penalty, r4 = 4 * r, dd = r - pen;
    // Keeping a hit is always allowed (the extension is recomputed behind the filter), so an evaluation stops as soon as
    // the running score says "keep": inside a true alignment that is after a few dozen bases instead of the whole
    // alignment, which a lone lane would otherwise walk while the other 31 of its warp wait
    const int32_t keep_at = max(cutoff, reduced);
    const int32_t shift = (4 - (s_off & 3)) & 3;
    const int32_t q_ext = q_off + shift, s_ext = s_off + shift;
    int32_t score = 0;
    {   // approximate pass, left: group k covers query [q_ext - 4k - 4, q_ext - 4k)
        const int32_t n = min(q_ext, s_ext) >> 2;
        int32_t sum = 0;
        bool stop = false;
        for (int32_t k = 0; k < n && !stop; k += 4) {
            uint32_t qb, qa;
            qwin(q, q_ext - 4 * k - 16, qb, qa);
            const uint32_t m = mismatch_bits(qb, 0u, direct_swin(d, tbase + s_ext - 4 * k - 16));
            const int cnt = min(4, n - k);
            for (int j = 0; j < cnt; j++) {
                if ((qa >> (8 * j)) & 0xFFu) return true;
                sum += r4 - dd * __popc((m >> (8 * j)) & 0xFFu);
                if (sum > 0) { score += sum; sum = 0; if (score >= keep_at) return true; }
                if (sum < X) { stop = true; break; }
            }
        }
    }
    {   // approximate pass, right: group k covers query [q_ext + 4k, q_ext + 4k + 4)
        const int32_t n = min(q.concat_len - q_ext, slen - s_ext) >> 2;
        int32_t sum = 0;
        bool stop = false;
        for (int32_t k = 0; k < n && !stop; k += 4) {
            uint32_t qb, qa;
            qwin(q, q_ext + 4 * k, qb, qa);
            const uint32_t m = mismatch_bits(qb, 0u, direct_swin(d, tbase + s_ext + 4 * k));
            const int cnt = min(4, n - k);
            for (int j = 0; j < cnt; j++) {
                if ((qa >> (24 - 8 * j)) & 0xFFu) return true;
                sum += r4 - dd * __popc((m >> (24 - 8 * j)) & 0xFFu);
                if (sum > 0) { score += sum; sum = 0; if (score >= keep_at) return true; }
                if (sum < X) { stop = true; break; }
            }
        }
    }
    if (score < reduced) return score >= cutoff;          // kept as computed by the approximate pass (:343-350)
    // exact pass (s_NuclUngappedExtendExact :153-245), run by run of matching bases
    score = 0;
    {   // left of q_off
        const int32_t n = min(q_off, s_off);
        int32_t sum = 0, done = 0;
        bool stop = false;
        while (done < n && !stop) {
            uint32_t qb, qa;
            qwin(q, q_off - done - 16, qb, qa);
            const int32_t rem = min(16, n - done);
            // step t reads base 15 - t of the window: its flag sits at bit 2t
            const uint32_t used = rem == 16 ? 0xFFFFFFFFu : ((1u << (2 * rem)) - 1u);
            if (qa & used) return true;
            const uint32_t m = mismatch_bits(qb, 0u, direct_swin(d, tbase + s_off - done - 16));
            int32_t t = 0;
            while (t < rem) {
                const uint32_t mm = m >> (2 * t);
                int32_t run = mm ? ((__ffs(mm) - 1) >> 1) : 16;
                run = min(run, rem - t);
                if (run > 0) { sum += run * r; if (sum > 0) { score += sum; sum = 0; if (score >= cutoff) return true; } t += run; }
                if (t >= rem) break;
                sum += pen;
                if (sum < X) { stop = true; break; }
                ++t;
            }
            done += rem;
        }
    }
    {   // right from q_off
        const int32_t n = min(q.concat_len - q_off, slen - s_off);
        int32_t sum = 0, done = 0;
        bool stop = false;
        while (done < n && !stop) {
            uint32_t qb, qa;
            qwin(q, q_off + done, qb, qa);
            const int32_t rem = min(16, n - done);
            const uint32_t used = rem == 16 ? 0xFFFFFFFFu : ~((1u << (32 - 2 * rem)) - 1u);
            if (qa & used) return true;
            const uint32_t m = mismatch_bits(qb, 0u, direct_swin(d, tbase + s_off + done));
            int32_t t = 0;
            while (t < rem) {
                const uint32_t mm = m << (2 * t);
                int32_t run = mm ? (__clz(mm) >> 1) : 16;
                run = min(run, rem - t);
                if (run > 0) { sum += run * r; if (sum > 0) { score += sum; sum = 0; if (score >= cutoff) return true; } t += run; }
                if (t >= rem) break;
                sum += pen;
                if (sum < X) { stop = true; break; }
                ++t;
            }
            done += rem;
        }
    }
    return score >= cutoff;
}

// ---- TMA (bulk async copy) + mbarrier helpers ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy executed by the TMA unit (SASS UBLKCP); src/dst 16-byte aligned, bytes % 16 == 0
// The subject streams through once per search: its sectors are marked evict-first in L2, so that they do not push
// out the lookup data (prk / sig / cinfo) every probe of the next tiles — and of the next search — gathers from there.
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int MAXC = 256;          // chunks one staged block may span (more: direct-load path)
int scan_max_block_chunks() { return MAXC; }
int scan_tile_margin() { return TILE_MARGIN; }

// 256-bit loads (LDG.E.256): a block descriptor / a cell's first two chain elements are one 32-byte sector
__device__ __forceinline__ ScanBlockDesc ld_block_desc(const ScanBlockDesc *p)
{
    uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7) : "l"(p));
    ScanBlockDesc d;
    d.tile_lo = (int64_t)(((uint64_t)a1 << 32) | a0);
    d.bytes = (int32_t)a2; d.c_lo = (int32_t)a3; d.c_hi = (int32_t)a4; d.staged = (int32_t)a5;
    d.pad0 = (int32_t)a6; d.pad1 = (int32_t)a7;
    return d;
}
__device__ __forceinline__ void ld_cinfo_pair(const uint4 *p, uint4 &e0, uint4 &e1)
{
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(e0.x), "=r"(e0.y), "=r"(e0.z), "=r"(e0.w), "=r"(e1.x), "=r"(e1.y), "=r"(e1.z), "=r"(e1.w) : "l"(p));
}

// Direct-load path for the rare block whose byte span does not fit the tile (runs of sequences
// shorter than a word) : same semantics, per-position global loads.
__device__ __noinline__ unsigned long long scan_block_direct(const DevQuery &q, const ScanLaunch &s, int64_t block_pos0,
                                                             int32_t c_lo, int32_t c_hi, int tid)
{
    unsigned long long my_lookup_hits = 0;
    const int32_t lut = q.lut_word_length, step = q.scan_step;
    for (int it = 0; it < POS_PER_THREAD; it++) {
        const int64_t g = block_pos0 + (int64_t)it * SCAN_THREADS + tid;
        if (g >= s.total_pos) break;
        int32_t lo = c_lo, hi = c_hi;
        while (lo < hi) {
            const int32_t m = (lo + hi + 1) >> 1;
            if (__ldg(&s.chunks[m].pos_prefix) <= g) lo = m; else hi = m - 1;
        }
        const DevChunk ch = s.chunks[lo];
        const int32_t p = ch.p_first + (int32_t)(g - ch.pos_prefix) * step;
        const uint32_t window = load_window(s.packed, ch.byte_off + (p >> 2));
        const uint32_t idx = (window >> (2 * (16 - ((p & 3) + lut)))) & q.hash_mask;
        lo = ch.parent;
        int32_t qp = mb_cell(q, idx);
        while (qp) {
            ++my_lookup_hits;
            int32_t qo, so;
            if (s.raw_pairs) emit_hit(q, s, (uint32_t)lo, (uint32_t)p, g, qp - 1, p);
            else if (mini_extend_mb(q, s.packed + ch.byte_off, ch.s_range, qp - 1, p, qo, so))
                emit_hit(q, s, (uint32_t)lo, (uint32_t)p, g, qo, so);
            qp = __ldg(&q.next_pos[qp]);
        }
    }
    return my_lookup_hits;
}


// One occupied table cell met at a scan position (rank = its index among the occupied cells; gl = block-relative
// position; k = index in the block's chunk table): flank-signature pre-filter, chain walk, mini-extension (or
// direct filter), emission.  Shared by the queue-driven kernel and the filtered one.
struct BlockView {
    const uint32_t *tile;
    const int32_t *ct_start, *ct_tbase, *ct_len, *ct_pfirst, *ct_parent;
    int64_t block_pos0, tile_lo;
    int32_t tile_bytes;
};
// Direct filter, dense: the lanes of a warp meet their non-follower hits at different steps of different chains, and
// a filter evaluation entered by 10 lanes out of 32 that then leave it one by one costs the warp ~700 instructions
// per call (C3: 3.7 G warp instructions for 100 Mb).  With a queue the chain walk only records {query position, block
// position | chunk << 11}; full sets of 32 records are then evaluated together.
constexpr int DQ_CAP = 128;
struct DirectQueue {
    uint2 rec[DQ_CAP];
    uint32_t count;
    uint32_t pad[3];
};
template <bool DIRECT>
__device__ __forceinline__ void scan_candidate(const DevQuery &q, const ScanLaunch &s, const BlockView &b, uint32_t rank,
                                               int32_t gl, int32_t k, bool use_sig, uint32_t &my_lookup_hits,
                                               DirectQueue *dq = nullptr)
{
    const int32_t lut = q.lut_word_length, step = q.scan_step;
    const uint32_t *tile = b.tile;
    // mini-extension pre-filter: a hit reaches the full word only if the 4 bases right of the lookup word all match
    // (when fewer than 4 match on the left, ext_to - 3 >= 4 are still owed on the right) or the 4 on the left do
    if (use_sig) {
        // 4 bytes per candidate from a table that stays in L2, instead of the 32-byte chain record from HBM:
        // nine out of ten candidates of a megablast batch end here
        const uint32_t sg = __ldg(&q.sig[rank]);
        const int32_t p0 = b.ct_pfirst[k] + (gl - b.ct_start[k]) * step, tb0 = b.ct_tbase[k];
        const uint32_t wl = tile_win(tile, tb0 + p0 - 4) >> 24, wr = tile_win(tile, tb0 + p0 + lut) >> 24;
        if (!(sg & 0x10000u) && wl != (sg & 0xFFu) && wr != ((sg >> 8) & 0xFFu)) { ++my_lookup_hits; return; }
    }
    // first two chain elements of the cell in ONE 32-byte sector: {qp | more << 31, left 16, right 16, ambiguity} x 2
    uint4 qi, qi1;
    ld_cinfo_pair(q.cinfo + 2 * (size_t)rank, qi, qi1);
    const int32_t p = b.ct_pfirst[k] + (gl - b.ct_start[k]) * step;
    const int32_t tbase = b.ct_tbase[k], len = b.ct_len[k];
    const uint32_t chunk = (uint32_t)b.ct_parent[k];
    const int64_t g = b.block_pos0 + gl;
    int32_t qp = (int32_t)(qi.x & QP_MASK);
    bool more = (qi.x >> 31) != 0, second = true;
    // base in front of the scan position (direct filter: is this hit the continuation of a run of matches?)
    const uint32_t s_prev = (DIRECT && p > 0) ? (tile_win(tile, tbase + p - 1) >> 30) : 4u;
    for (;;) {
        ++my_lookup_hits;
        int32_t qo, so;
        if (DIRECT) {
            // lut == word: no mini-extension.  A hit whose predecessor on the diagonal, (q_off - 1, s_off - 1),
            // is a lookup hit too belongs to the same run of matching bases as that hit and shares its fate:
            // rejected by the diagonal test when the run's first hit was extended successfully, below the
            // cutoff itself when that one was.  Only a run's first hit goes on, and only if its own
            // ungapped extension can reach the cutoff.
            const bool follower = (qi.x & PREV_INDEXED) && s_prev == (qi.y & 3u) && !(qi.w & 1u);
            uint32_t slot = DQ_CAP;
            if (!follower && dq) slot = atomicAdd(&dq->count, 1u);
            if (slot < (uint32_t)DQ_CAP) dq->rec[slot] = make_uint2((uint32_t)qp, (uint32_t)gl | ((uint32_t)k << 11));
            else if (!follower) {
                int32_t xd = s.uni_x, co = s.uni_cutoff, rc = s.uni_reduced;
                if (!s.uni_ok) {
                    const DevContext c = q.ctx[ctx_search(q, qp - 1)];
                    xd = c.x_dropoff; co = c.cutoff_score; rc = c.reduced_cutoff;
                }
                DirectCtx dc{tile, b.tile_bytes * 4, b.tile_lo * 4, s.packed};
                if (direct_keep(q, dc, tbase, len, qp - 1, p, xd, co, rc))
                    emit_hit(q, s, chunk, (uint32_t)p, g, qp - 1, p);
            }
        }
        else if (s.raw_pairs) emit_hit(q, s, chunk, (uint32_t)p, g, qp - 1, p);
        else if (mini_extend_tile(q, tile, tbase, len, qp - 1, p, qi, qo, so))
            emit_hit(q, s, chunk, (uint32_t)p, g, qo, so);
        if (!more) break;
        if (second) {                       // second element came with the first
            second = false;
            qi = qi1;
            qp = (int32_t)(qi.x & QP_MASK);
            more = (qi.x >> 31) != 0;
        } else {                            // third and later (rare): pointer chase
            qp = __ldg(&q.next_pos[qp]);
            qi = __ldg(&q.qinfo[qp]);       // {next | PREV_INDEXED, left 16 bases, right 16 bases, ambiguity}
            more = (qi.x & QP_MASK) != 0;
        }
    }
}

// Evaluates queued hits of the direct filter, 32 at a time (only full sets unless `all`); what is left moves to the front.
__device__ __forceinline__ void direct_drain(const DevQuery &q, const ScanLaunch &s, const BlockView &b, DirectQueue *dq, bool all,
                                             int lane)
{
    __syncwarp();
    const int32_t n = (int32_t)min(dq->count, (uint32_t)DQ_CAP);
    int32_t done = 0;
    const int32_t step = q.scan_step;
    while (n - done >= 32 || (all && done < n)) {
        const int32_t i = done + lane;
        if (i < n) {
            const uint2 r = dq->rec[i];
            const int32_t qp = (int32_t)r.x, gl = (int32_t)(r.y & 2047u), k = (int32_t)(r.y >> 11);
            const int32_t p = b.ct_pfirst[k] + (gl - b.ct_start[k]) * step;
            int32_t xd = s.uni_x, co = s.uni_cutoff, rc = s.uni_reduced;
            if (!s.uni_ok) {
                const DevContext c = q.ctx[ctx_search(q, qp - 1)];
                xd = c.x_dropoff; co = c.cutoff_score; rc = c.reduced_cutoff;
            }
            DirectCtx dc{b.tile, b.tile_bytes * 4, b.tile_lo * 4, s.packed};
            if (direct_keep(q, dc, b.ct_tbase[k], b.ct_len[k], qp - 1, p, xd, co, rc))
                emit_hit(q, s, (uint32_t)b.ct_parent[k], (uint32_t)p, b.block_pos0 + gl, qp - 1, p);
        }
        done += 32;
        __syncwarp();
    }
    done = min(done, n);
    const int32_t left = n - done;
    uint2 r = make_uint2(0u, 0u);
    if (lane < left) r = dq->rec[done + lane];
    __syncwarp();
    if (lane < left) dq->rec[lane] = r;
    if (lane == 0) dq->count = (uint32_t)left;
    __syncwarp();
}

// STEP: compile-time scan stride for the consecutive-position word loader (blocks inside one chunk), 0 = run-time stride
// DENSE: the probe is the per-cell signature table (DevQuery::psig) instead of {presence, rank}: a position becomes a
// candidate only if its cell is occupied AND the signature lets the mini-extension succeed (1 % of the positions of a
// megablast batch instead of 12 %); the rank is looked up for those alone
template <bool DIRECT, int STEP, bool DENSE = false>
__global__ void __launch_bounds__(SCAN_THREADS, DIRECT ? 4 : BN_SCAN_MIN_BLOCKS)
scan_kernel_staged(const __grid_constant__ DevQuery q, const __grid_constant__ ScanLaunch s)
{
    extern __shared__ __align__(128) uint32_t smem_dyn[];
    uint32_t *tile = smem_dyn;                                               // s.tile_cap bytes
    uint2 *cand = reinterpret_cast<uint2 *>(smem_dyn + s.tile_cap / 4);      // POS_PER_BLOCK entries {rank, where}
    // block-local chunk table: first position (block-relative, may be negative for the first chunk),
    // tile-relative base index of the chunk's base 0, chunk length
    __shared__ int32_t ct_start[MAXC + 1], ct_tbase[MAXC], ct_len[MAXC], ct_pfirst[MAXC], ct_parent[MAXC];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t block_pos0 = (int64_t)blockIdx.x * POS_PER_BLOCK;
    const int32_t npos = (int32_t)min((int64_t)POS_PER_BLOCK, s.total_pos - block_pos0);
    const int32_t lut = q.lut_word_length, step = q.scan_step;
    // the block's slice of the volume was worked out once per (volume, table shape) on the host:
    // one uniform 32-byte load instead of a binary search + two chunk loads in front of the TMA issue
    const ScanBlockDesc bd = ld_block_desc(s.block_desc + blockIdx.x);
    if (tid == 0 && bd.staged) {                // the block's slice of the packed subject: one TMA bulk copy
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, (uint32_t)bd.bytes);
        tma_bulk_g2s(tile, s.packed + bd.tile_lo, (uint32_t)bd.bytes, &bar);
    }
    const int32_t c_lo = bd.c_lo;
    const int32_t nch = bd.c_hi - c_lo + 1;
    uint32_t my_lookup_hits = 0;
    if (!bd.staged) {
        my_lookup_hits = (uint32_t)scan_block_direct(q, s, block_pos0, c_lo, bd.c_hi, tid);
    } else {
        const int64_t tile_lo = bd.tile_lo;
        for (int i = tid; i < nch; i += SCAN_THREADS) {
            const DevChunk c = s.chunks[c_lo + i];
            ct_start[i] = (int32_t)(c.pos_prefix - block_pos0);
            ct_tbase[i] = (int32_t)((c.byte_off - tile_lo) * 4);
            ct_len[i] = c.s_range;              // right bound of the unit's range (the chunk length without masks)
            ct_pfirst[i] = c.p_first;
            ct_parent[i] = c.parent;
        }
        if (tid == 0) ct_start[nch] = INT32_MAX;
        __syncthreads();
        mbar_wait(&bar, 0);

        // From here on every warp works alone on its 256 positions (block-relative position of lane l,
        // round it: it * 256 + warp * 32 + l): no block barrier between the probe and the candidate phase.
        // ---- phase A: lookup words from the tile, all 8 presence probes of the thread in flight ------
        uint2 words[DENSE ? 1 : POS_PER_THREAD];
        uint32_t bitpack[POS_PER_THREAD / 4], cpack[POS_PER_THREAD / 4];
        // DENSE: psig entry, table cell and the bits the entry is compared with, per position
        uint32_t pw[DENSE ? POS_PER_THREAD : 1], cell[DENSE ? POS_PER_THREAD : 1], want[DENSE ? POS_PER_THREAD : 1];
        const uint32_t shr = 32u - 2u * (uint32_t)lut;
        // thread -> position map: round it of thread tid handles block-relative position it * 256 + tid, or — blocks
        // inside one chunk with a compile-time stride — tid * 8 + it (consecutive positions, words8)
        const bool consec = STEP > 0 && POS_PER_THREAD == 8 && nch == 1;
        if constexpr (DENSE) {
            if constexpr (STEP > 0 && POS_PER_THREAD == 8) {
                if (consec) {
                    const int32_t tb0 = ct_tbase[0] + ct_pfirst[0] - ct_start[0] * STEP;
                    words8f<STEP>(tile, tb0 + tid * 8 * STEP, shr, cell, want);
#pragma unroll
                    for (int it = 0; it < 8; it++) pw[it] = __ldg(&q.psig[cell[it]]);
                    for (int i = 0; i < POS_PER_THREAD / 4; i++) cpack[i] = 0;
                }
            }
            if (!consec) {
                int32_t c = 0, cnext = ct_start[1];
#pragma unroll
                for (int it = 0; it < POS_PER_THREAD; it++) {
                    const int32_t gl = min(it * SCAN_THREADS + tid, npos - 1);
                    while (gl >= cnext) { ++c; cnext = ct_start[c + 1]; }          // monotone cursor over the block's chunk table
                    if ((it & 3) == 0) cpack[it >> 2] = 0;
                    cpack[it >> 2] |= (uint32_t)c << (8 * (it & 3));
                    const int32_t tb = ct_tbase[c] + ct_pfirst[c] + (gl - ct_start[c]) * step;
                    probe_key(tile_win(tile, tb - 4), tile_win(tile, tb + 12), shr, cell[it], want[it]);
                    pw[it] = __ldg(&q.psig[cell[it]]);
                }
            }
        }
        if constexpr (!DENSE && STEP > 0 && POS_PER_THREAD == 8) {
            if (consec) {
                const int32_t tb0 = ct_tbase[0] + ct_pfirst[0] - ct_start[0] * STEP;
                uint32_t idxs[8];
                words8<STEP>(tile, tb0 + tid * 8 * STEP, shr, idxs);
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    words[it] = __ldg(&q.prk[idxs[it] >> 5]);
                    if ((it & 3) == 0) bitpack[it >> 2] = 0;
                    bitpack[it >> 2] |= (idxs[it] & 31u) << (8 * (it & 3));
                }
                for (int i = 0; i < POS_PER_THREAD / 4; i++) cpack[i] = 0;
            }
        }
        if constexpr (!DENSE) {
        if (consec) {
        } else if (nch == 1) {
            // the whole block lies in one chunk: tile offsets are an arithmetic progression
            const int32_t tb0 = ct_tbase[0] + ct_pfirst[0] - ct_start[0] * step;
#pragma unroll
            for (int it = 0; it < POS_PER_THREAD; it++) {
                const int32_t gl = min(it * SCAN_THREADS + tid, npos - 1);
                const int32_t tb = tb0 + gl * step;
                const uint32_t w0 = tile[tb >> 4], w1 = tile[(tb >> 4) + 1];
                // 4 consecutive stream bytes from byte (tb >> 2) on, most significant first
                const uint32_t W = __byte_perm(w0, w1, 0x0123u + 0x1111u * ((uint32_t)(tb >> 2) & 3u));
                const uint32_t idx = (W << (2u * ((uint32_t)tb & 3u))) >> shr;
                words[it] = __ldg(&q.prk[idx >> 5]);
                if ((it & 3) == 0) bitpack[it >> 2] = 0;
                bitpack[it >> 2] |= (idx & 31u) << (8 * (it & 3));
            }
            for (int i = 0; i < POS_PER_THREAD / 4; i++) cpack[i] = 0;
        } else {
            {   // chunk of each of the thread's positions: monotone cursor over the block's chunk table
                int32_t c = 0, cnext = ct_start[1];
#pragma unroll
                for (int it = 0; it < POS_PER_THREAD; it++) {
                    const int32_t gl = min(it * SCAN_THREADS + tid, npos - 1);
                    while (gl >= cnext) { ++c; cnext = ct_start[c + 1]; }
                    if ((it & 3) == 0) cpack[it >> 2] = 0;
                    cpack[it >> 2] |= (uint32_t)c << (8 * (it & 3));
                }
            }
#pragma unroll
            for (int it = 0; it < POS_PER_THREAD; it++) {
                const int32_t gl = min(it * SCAN_THREADS + tid, npos - 1);
                const uint32_t c = (cpack[it >> 2] >> (8 * (it & 3))) & 255u;
                const int32_t tb = ct_tbase[c] + ct_pfirst[c] + (gl - ct_start[c]) * step;
                const uint32_t w0 = tile[tb >> 4], w1 = tile[(tb >> 4) + 1];
                const uint32_t W = __byte_perm(w0, w1, 0x0123u + 0x1111u * ((uint32_t)(tb >> 2) & 3u));
                const uint32_t idx = (W << (2u * ((uint32_t)tb & 3u))) >> shr;
                words[it] = __ldg(&q.prk[idx >> 5]);
                if ((it & 3) == 0) bitpack[it >> 2] = 0;
                bitpack[it >> 2] |= (idx & 31u) << (8 * (it & 3));
            }
        }
        }
        // ---- warp-local compaction of the occupied cells (ballots, no atomics) ----------------------
        uint2 *wcand = cand + (tid >> 5) * (32 * POS_PER_THREAD);
        int ncand = 0;
        {
            const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
            for (int it = 0; it < POS_PER_THREAD; it++) {
                const int32_t gl = consec ? tid * POS_PER_THREAD + it : it * SCAN_THREADS + tid;
                bool hit;
                uint32_t rec;                   // DENSE: the table cell (its rank is looked up in phase B), else the rank
                if constexpr (DENSE) {
                    const bool occupied = (gl < npos) && (pw[it] & 0x8000u);
                    hit = occupied && psig_pass(pw[it], want[it]);
                    my_lookup_hits += (occupied && !hit) ? 1u : 0u;         // the one chain element that cannot reach the word
                    rec = cell[it];
                } else {
                    const uint32_t bit = (bitpack[it >> 2] >> (8 * (it & 3))) & 31u;
                    const uint2 wd = words[it];
                    hit = (gl < npos) && ((wd.x >> bit) & 1u);
                    rec = wd.y + (uint32_t)__popc(wd.x & ((1u << bit) - 1u));
                }
                const uint32_t m = __ballot_sync(0xffffffffu, hit);
                if (hit) {
                    const uint32_t c = (cpack[it >> 2] >> (8 * (it & 3))) & 255u;
                    wcand[ncand + __popc(m & lt)] = make_uint2(rec, (c << 11) | (uint32_t)gl);
                }
                ncand += __popc(m);
            }
        }
        __syncwarp();

        // ---- phase B: the warp's candidates, one per lane ----------------------------------------------
        const bool use_sig = !DIRECT && !s.raw_pairs && q.sig != nullptr && (q.word_length - lut) >= 7;
        const BlockView bv{tile, ct_start, ct_tbase, ct_len, ct_pfirst, ct_parent, block_pos0, bd.tile_lo, bd.bytes};
        if constexpr (DIRECT) {
            DirectQueue *dq = reinterpret_cast<DirectQueue *>(cand + POS_PER_BLOCK) + (tid >> 5);
            if (lane == 0) dq->count = 0;
            __syncwarp();
            for (int c0 = 0; c0 < ncand; c0 += 32) {
                const int ci = c0 + lane;
                if (ci < ncand) {
                    const uint2 cd = wcand[ci];
                    scan_candidate<DIRECT>(q, s, bv, cd.x, (int32_t)(cd.y & 2047u), (int32_t)(cd.y >> 11), use_sig, my_lookup_hits,
                                           s.direct_dense ? dq : nullptr);
                }
                direct_drain(q, s, bv, dq, c0 + 32 >= ncand, lane);
            }
        } else {
            for (int ci = lane; ci < ncand; ci += 32) {
                const uint2 cd = wcand[ci];
                uint32_t rank = cd.x;
                if constexpr (DENSE) {
                    const uint2 wd = __ldg(&q.prk[cd.x >> 5]);
                    rank = wd.y + (uint32_t)__popc(wd.x & ((1u << (cd.x & 31u)) - 1u));
                }
                scan_candidate<DIRECT>(q, s, bv, rank, (int32_t)(cd.y & 2047u), (int32_t)(cd.y >> 11), use_sig && !DENSE, my_lookup_hits);
            }
        }
    }
    // one atomic per warp for the lookup-hit statistic (BlastUngappedStats.lookup_hits); REDUX.SUM
    const uint32_t warp_hits = __reduce_add_sync(0xffffffffu, my_lookup_hits);
    if (lane == 0 && warp_hits) atomicAdd(&s.counters[1], (unsigned long long)warp_hits);
}


// ---- filtered kernel (small megablast tables) ---------------------------------------------------------
// A small query batch (one 10 kb query = 20 k table entries in 4^11 cells) leaves almost every cell empty, yet in the
// queue-driven kernel above every scan position still costs one random L2 sector.  Here a hashed presence filter of
// the table (2^20 or 2^19 bits, DevQuery::filt) lives in SHARED memory: a probe is one LDS with a few-way bank
// conflict instead of 32 L1 wavefronts per warp, and only the positions it passes (table density) go on to the exact
// {presence, rank} word in L2.  One persistent 1024-thread CTA per SM holds the filter; its four 256-thread groups
// each walk their own sequence of 2048-position blocks (same block descriptors, same staged slice by TMA, same
// candidate routine as above), synchronised by a named barrier per group.  Candidates are rare by construction, so
// they are handled right where they are met, without a queue.
constexpr int FILT_GROUPS = 4;
__host__ __device__ constexpr int filt_ct_ints(int maxc) { return 5 * maxc + 4; }     // ct_start[maxc + 1] + 4 arrays, 16-byte multiple
__device__ __forceinline__ void group_sync(int group)
{
    asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "r"(SCAN_THREADS) : "memory");
}

// Each group double-buffers: while it works on block r, the slice of block r + 1 is in flight (TMA) and its chunk
// table is already written; the descriptor of block r + 2 is loaded a round ahead.  maxc = chunks a block may span
// and still be staged here (a block with more takes the direct-load path).
template <bool DIRECT, int STEP>
__global__ void __launch_bounds__(FILT_GROUPS * SCAN_THREADS, 1)
scan_kernel_filtered(const __grid_constant__ DevQuery q, const __grid_constant__ ScanLaunch s, const int32_t filt_log2,
                     const int32_t maxc)
{
    extern __shared__ __align__(128) uint32_t smem_dyn[];
    __shared__ __align__(8) unsigned long long bars[FILT_GROUPS][2];
    const uint32_t fwords = 1u << (filt_log2 - 5), fmask = (1u << filt_log2) - 1u;
    uint32_t *filt = smem_dyn;
    const int group = threadIdx.x / SCAN_THREADS, tid = threadIdx.x % SCAN_THREADS, lane = tid & 31;
    const int32_t stage_words = s.tile_cap / 4 + filt_ct_ints(maxc);
    uint32_t *gmem = smem_dyn + fwords + (size_t)group * 2 * stage_words;

    {   // the filter: 128 KB (or the OR of its two halves) from L2, once per CTA
        const uint4 *src = reinterpret_cast<const uint4 *>(q.filt);
        uint4 *dst = reinterpret_cast<uint4 *>(filt);
        const uint32_t n4 = fwords / 4;
        if (filt_log2 == FILT_LOG2)
            for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
        else
            for (uint32_t i = threadIdx.x; i < n4; i += blockDim.x) {
                const uint4 a = __ldg(src + i), b = __ldg(src + i + n4);
                dst[i] = make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w);
            }
    }
    if (tid == 0) { mbar_init(&bars[group][0], 1); mbar_init(&bars[group][1], 1); }
    __syncthreads();

    const int32_t lut = q.lut_word_length, step = q.scan_step;
    const uint32_t shr = 32u - 2u * (uint32_t)lut;
    const bool use_sig = !DIRECT && !s.raw_pairs && q.sig != nullptr && (q.word_length - lut) >= 7;
    const int64_t n_blocks = (s.total_pos + POS_PER_BLOCK - 1) / POS_PER_BLOCK;
    const int64_t stride = (int64_t)gridDim.x * FILT_GROUPS;
    uint32_t my_lookup_hits = 0, parity = 0;        // parity: bit st = phase of stage st's barrier

    // stage a block: TMA of its slice + its chunk table (block-relative first positions, tile-relative bases)
    auto prefetch = [&](const ScanBlockDesc &d, int64_t vb, int st) {
        const int32_t nch = d.c_hi - d.c_lo + 1;
        if (!d.staged || nch > maxc) return;
        if (tid >= 32 && tid >= nch) return;                // warps with nothing to write (the usual block spans a few chunks)
        uint32_t *tile = gmem + (size_t)st * stage_words;
        if (tid == 0) {
            mbar_expect_tx(&bars[group][st], (uint32_t)d.bytes);
            tma_bulk_g2s(tile, s.packed + d.tile_lo, (uint32_t)d.bytes, &bars[group][st]);
        }
        int32_t *ct_start = reinterpret_cast<int32_t *>(tile + s.tile_cap / 4);
        int32_t *ct_tbase = ct_start + maxc + 4, *ct_len = ct_tbase + maxc, *ct_pfirst = ct_len + maxc, *ct_parent = ct_pfirst + maxc;
        const int64_t block_pos0 = vb * POS_PER_BLOCK;
        for (int i = tid; i < nch; i += SCAN_THREADS) {
            const DevChunk c = s.chunks[d.c_lo + i];
            ct_start[i] = (int32_t)(c.pos_prefix - block_pos0);
            ct_tbase[i] = (int32_t)((c.byte_off - d.tile_lo) * 4);
            ct_len[i] = c.s_range;
            ct_pfirst[i] = c.p_first;
            ct_parent[i] = c.parent;
        }
        if (tid == 0) ct_start[nch] = INT32_MAX;
    };

    int64_t vb = (int64_t)blockIdx.x * FILT_GROUPS + group;
    ScanBlockDesc bd{}, bd1{};
    if (vb < n_blocks) { bd = ld_block_desc(s.block_desc + vb); prefetch(bd, vb, 0); }
    if (vb + stride < n_blocks) bd1 = ld_block_desc(s.block_desc + vb + stride);
    group_sync(group);
    for (int st = 0; vb < n_blocks; vb += stride, st ^= 1) {
        ScanBlockDesc bd2{};
        if (vb + stride < n_blocks) prefetch(bd1, vb + stride, st ^ 1);
        if (vb + 2 * stride < n_blocks) bd2 = ld_block_desc(s.block_desc + vb + 2 * stride);

        const int64_t block_pos0 = vb * POS_PER_BLOCK;
        const int32_t npos = (int32_t)min((int64_t)POS_PER_BLOCK, s.total_pos - block_pos0);
        const int32_t nch = bd.c_hi - bd.c_lo + 1;
        if (!bd.staged || nch > maxc) {
            my_lookup_hits += (uint32_t)scan_block_direct(q, s, block_pos0, bd.c_lo, bd.c_hi, tid);
        } else {
            const uint32_t *tile = gmem + (size_t)st * stage_words;
            const int32_t *ct_start = reinterpret_cast<const int32_t *>(tile + s.tile_cap / 4);
            const int32_t *ct_tbase = ct_start + maxc + 4, *ct_len = ct_tbase + maxc, *ct_pfirst = ct_len + maxc,
                          *ct_parent = ct_pfirst + maxc;
            mbar_wait(&bars[group][st], (parity >> st) & 1u);
            parity ^= 1u << st;
            const BlockView bv{tile, ct_start, ct_tbase, ct_len, ct_pfirst, ct_parent, block_pos0, bd.tile_lo, bd.bytes};
            // lookup words of the thread's 8 positions, filter bits gathered first (8 independent LDS), then the few that pass
            uint32_t idxs[POS_PER_THREAD];
            uint32_t cpack[POS_PER_THREAD / 4];
            uint32_t pass = 0;
            const bool consec = STEP > 0 && POS_PER_THREAD == 8 && nch == 1;      // thread -> position map, see scan_kernel_staged
            if constexpr (STEP > 0 && POS_PER_THREAD == 8) {
                if (consec) {
                    const int32_t tb0 = ct_tbase[0] + ct_pfirst[0] - ct_start[0] * STEP;
                    words8<STEP>(tile, tb0 + tid * 8 * STEP, shr, idxs);
                    for (int i = 0; i < POS_PER_THREAD / 4; i++) cpack[i] = 0;
                }
            }
            if (consec) {
            } else if (nch == 1) {
                const int32_t tb0 = ct_tbase[0] + ct_pfirst[0] - ct_start[0] * step;
#pragma unroll
                for (int it = 0; it < POS_PER_THREAD; it++) {
                    const int32_t gl = min(it * SCAN_THREADS + tid, npos - 1);
                    const int32_t tb = tb0 + gl * step;
                    const uint32_t w0 = tile[tb >> 4], w1 = tile[(tb >> 4) + 1];
                    const uint32_t W = __byte_perm(w0, w1, 0x0123u + 0x1111u * ((uint32_t)(tb >> 2) & 3u));
                    idxs[it] = (W << (2u * ((uint32_t)tb & 3u))) >> shr;
                }
                for (int i = 0; i < POS_PER_THREAD / 4; i++) cpack[i] = 0;
            } else {
                int32_t c = 0, cnext = ct_start[1];
#pragma unroll
                for (int it = 0; it < POS_PER_THREAD; it++) {
                    const int32_t gl = min(it * SCAN_THREADS + tid, npos - 1);
                    while (gl >= cnext) { ++c; cnext = ct_start[c + 1]; }
                    if ((it & 3) == 0) cpack[it >> 2] = 0;
                    cpack[it >> 2] |= (uint32_t)c << (8 * (it & 3));
                    const int32_t tb = ct_tbase[c] + ct_pfirst[c] + (gl - ct_start[c]) * step;
                    const uint32_t w0 = tile[tb >> 4], w1 = tile[(tb >> 4) + 1];
                    const uint32_t W = __byte_perm(w0, w1, 0x0123u + 0x1111u * ((uint32_t)(tb >> 2) & 3u));
                    idxs[it] = (W << (2u * ((uint32_t)tb & 3u))) >> shr;
                }
            }
#pragma unroll
            for (int it = 0; it < POS_PER_THREAD; it++) {
                const uint32_t h = (idxs[it] ^ (idxs[it] >> FILT_LOG2)) & fmask;
                const uint32_t f = filt[h >> 5];
                pass |= (__funnelshift_r(f, 0u, h) & 1u) << it;           // shift amount taken mod 32
            }
            {   // positions behind the end of the volume (last block only)
                const int32_t left = consec ? npos - tid * POS_PER_THREAD : (npos - tid + SCAN_THREADS - 1) / SCAN_THREADS;
                if (left < POS_PER_THREAD) pass &= left <= 0 ? 0u : (1u << left) - 1u;
            }
            while (pass) {
                const int it = __ffs(pass) - 1;
                pass &= pass - 1u;
                uint32_t idx = idxs[0], cp = cpack[0];
#pragma unroll
                for (int j = 1; j < POS_PER_THREAD; j++) if (j == it) idx = idxs[j];
#pragma unroll
                for (int j = 1; j < POS_PER_THREAD / 4; j++) if (j == (it >> 2)) cp = cpack[j];
                const uint2 w = __ldg(&q.prk[idx >> 5]);
                const uint32_t bit = idx & 31u;
                if ((w.x >> bit) & 1u)
                    scan_candidate<DIRECT>(q, s, bv, w.y + (uint32_t)__popc(w.x & ((1u << bit) - 1u)),
                                           consec ? tid * POS_PER_THREAD + it : it * SCAN_THREADS + tid,
                                           (int32_t)((cp >> (8 * (it & 3))) & 255u), use_sig, my_lookup_hits);
            }
        }
        group_sync(group);          // stage st is free for block r + 2; block r + 1's chunk table is complete
        bd = bd1; bd1 = bd2;
    }
    const uint32_t warp_hits = __reduce_add_sync(0xffffffffu, my_lookup_hits);
    if (lane == 0 && warp_hits) atomicAdd(&s.counters[1], (unsigned long long)warp_hits);
}

__global__ void build_filter_kernel(const uint32_t *presence, int64_t nwords, uint32_t *filt)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    uint32_t m = presence[w];
    while (m) {
        const uint32_t h = filt_hash((uint32_t)(w * 32) + (uint32_t)(__ffs(m) - 1));
        m &= m - 1u;
        atomicOr(&filt[h >> 5], 1u << (h & 31u));
    }
}
cudaError_t launch_build_filter(const uint32_t *presence, int64_t nwords, uint32_t *filt, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(filt, 0, FILT_BITS / 8, st);
    if (e != cudaSuccess) return e;
    build_filter_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(presence, nwords, filt);
    return cudaGetLastError();
}

// sig[rank]: see DevQuery::sig; from the cell's first chain record (cinfo[2 rank]: .y = 16 bases left of the word, the
// nearest in the low bits; .z = 16 bases right of it, the nearest in the high bits; .x bit 31 = chain continues)
__global__ void build_sig_kernel(const uint4 *cinfo, int64_t n_ranks, uint32_t *sig)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_ranks) return;
    const uint4 c = cinfo[2 * r];
    sig[r] = (c.y & 0xFFu) | ((c.z >> 24) << 8) | ((c.x >> 31) << 16);
}
cudaError_t launch_build_sig(const uint4 *cinfo, int64_t n_ranks, uint32_t *sig, cudaStream_t st)
{
    if (n_ranks > 0) build_sig_kernel<<<(unsigned)((n_ranks + 255) / 256), 256, 0, st>>>(cinfo, n_ranks, sig);
    return cudaGetLastError();
}
// psig[cell]: see DevQuery::psig; the cell's sig entry spread over the 4^lut cells (two cells per thread, one 4-byte store)
__global__ void build_psig_kernel(const uint2 *prk, const uint32_t *sig, int64_t n_cells, uint16_t *psig)
{
    const int64_t c0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (c0 >= n_cells) return;
    const uint2 w = prk[c0 >> 5];
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const uint32_t bit = (uint32_t)(c0 & 31) + k;
        if ((w.x >> bit) & 1u) {
            const uint32_t sg = sig[w.y + (uint32_t)__popc(w.x & ((1u << bit) - 1u))];
            out |= (0x8000u | ((sg >> 16) & 1u) << 14 | (sg & 0x3Fu) << 8 | ((sg >> 8) & 0xFFu)) << (16 * k);
        }
    }
    if (c0 + 1 < n_cells) *reinterpret_cast<uint32_t *>(psig + c0) = out;
    else psig[c0] = (uint16_t)out;
}
cudaError_t launch_build_psig(const uint2 *prk, const uint32_t *sig, int64_t n_cells, uint16_t *psig, cudaStream_t st)
{
    const int64_t threads = (n_cells + 1) / 2;
    if (threads > 0) build_psig_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(prk, sig, n_cells, psig);
    return cudaGetLastError();
}

// qinfo[qp] for every 1-based query position qp: {next_pos[qp], 16 bases left of the lookup word that
// starts at qp-1, 16 bases right of it, ambiguity flags (left in the even bits, right in the odd bits)}
// indexed[] bit qp <=> the 1-based query position qp is in the lookup table = it heads a chain or some position links to it
__global__ void mark_linked_kernel(const int32_t *next_pos, int32_t concat_len, uint32_t *indexed)
{
    const int64_t qp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qp < 1 || qp > concat_len) return;
    const int32_t n = next_pos[qp];
    if (n > 0) atomicOr(&indexed[n >> 5], 1u << (n & 31));
}
__global__ void mark_heads_kernel(const int32_t *heads, int64_t n, uint32_t *indexed)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t h = heads[i];
    if (h > 0) atomicOr(&indexed[h >> 5], 1u << (h & 31));
}
__global__ void build_qinfo_kernel(const DevQuery q, const int32_t *next_pos, int32_t concat_len, const uint32_t *indexed,
                                   uint4 *qinfo)
{
    const int64_t qp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (qp > concat_len) return;
    uint32_t lb = 0, la = 0x55555555u, rb = 0, ra = 0x55555555u;
    if (qp >= 1) {
        qwin(q, (int32_t)qp - 1 - 16, lb, la);
        qwin(q, (int32_t)qp - 1 + q.lut_word_length, rb, ra);
    }
    uint32_t x = (uint32_t)next_pos[qp];
    if (indexed && qp >= 2 && ((indexed[(qp - 1) >> 5] >> ((qp - 1) & 31)) & 1u)) x |= PREV_INDEXED;
    qinfo[qp] = make_uint4(x, lb, rb, (la & 0x55555555u) | ((ra & 0x55555555u) << 1));
}
// heads: the chain heads (hashtable[] of a host-built table, or the per-rank first positions of the device fill)
cudaError_t launch_build_qinfo(const DevQuery &q, const int32_t *next_pos, int32_t concat_len, const int32_t *heads,
                               int64_t n_heads, uint32_t *indexed_scratch, uint4 *qinfo, cudaStream_t st)
{
    const int64_t n = (int64_t)concat_len + 1;
    if (indexed_scratch) {
        cudaError_t e = cudaMemsetAsync(indexed_scratch, 0, (size_t)((n + 32) / 32 + 1) * sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
        mark_linked_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(next_pos, concat_len, indexed_scratch);
        if (n_heads > 0) mark_heads_kernel<<<(unsigned)((n_heads + 255) / 256), 256, 0, st>>>(heads, n_heads, indexed_scratch);
    }
    build_qinfo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, next_pos, concat_len, indexed_scratch, qinfo);
    return cudaGetLastError();
}

// ---- query-load helpers: derived device arrays -----------------------------------------------------
// presence[w] bit b  <=>  hashtable[32 w + b] != 0.  One warp per 1024 cells: coalesced loads + ballot.
__global__ void build_presence_kernel(const int32_t *hashtable, int64_t hashsize, uint32_t *presence)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t cell0 = warp * 1024;
    if (cell0 >= hashsize) return;
    uint32_t mine = 0;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        const int64_t c = cell0 + 32 * i + lane;
        const uint32_t bits = __ballot_sync(0xffffffffu, c < hashsize && hashtable[c] != 0);
        if (lane == i) mine = bits;
    }
    const int64_t w = (cell0 >> 5) + lane;
    if (w < (hashsize + 31) / 32) presence[w] = mine;
}

// qpk[i]: 16-base window of concatenated-query positions 16 (i-1) .. 16 (i-1) + 15 (bn_device.cuh: qwin)
__global__ void build_qpk_kernel(const uint8_t *query_start, int32_t concat_len, uint2 *qpk, int64_t nwords)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    uint32_t bases = 0, amb = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int64_t pos = 16 * (i - 1) + j;
        const uint32_t code = (pos >= -1 && pos <= concat_len) ? query_start[pos + 1] : 15u;
        bases |= (code & 3u) << (30 - 2 * j);
        amb |= (code >= 4u ? 1u : 0u) << (30 - 2 * j);
    }
    qpk[i] = make_uint2(bases, amb);
}

// prk[w] = {presence word, number of occupied cells before word w}; dense[rank] = hashtable value of
// the rank-th occupied cell.  `prefix` = exclusive scan of popcounts (prefix_sum_u32, radix_sort.cu).
__global__ void popc_kernel(const uint32_t *presence, int64_t nwords, uint32_t *counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) counts[i] = __popc(presence[i]);
}
// cinfo[2 rank], cinfo[2 rank + 1] = qinfo of the cell's first / second chain element with
// .x = {qp, bit 31 = chain continues}; the second entry is zero for single-element cells
__device__ __forceinline__ void store_cinfo_pair(uint4 *cinfo, uint32_t rank, uint32_t qp, const uint4 *qinfo)
{
    uint4 v = qinfo[qp], w = make_uint4(0, 0, 0, 0);
    const uint32_t nxt = v.x & QP_MASK;
    v.x = qp | (nxt ? 0x80000000u : 0u) | (v.x & PREV_INDEXED);
    if (nxt) {
        w = qinfo[nxt];
        w.x = nxt | ((w.x & QP_MASK) ? 0x80000000u : 0u) | (w.x & PREV_INDEXED);
    }
    cinfo[2 * (size_t)rank] = v;
    cinfo[2 * (size_t)rank + 1] = w;
}
__global__ void build_compact_kernel(const int32_t *hashtable, const uint32_t *presence, const uint32_t *prefix,
                                     int64_t nwords, uint2 *prk, const uint4 *qinfo, uint4 *cinfo)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    uint32_t bits = presence[i];
    uint32_t r = prefix[i];
    prk[i] = make_uint2(bits, r);
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        store_cinfo_pair(cinfo, r++, (uint32_t)hashtable[i * 32 + b], qinfo);
    }
}
// device-built tables: prk from the presence bitmap, cinfo[rank] from the per-rank chain heads
__global__ void build_prk_kernel(const uint32_t *presence, const uint32_t *prefix, int64_t nwords, uint2 *prk)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) prk[i] = make_uint2(presence[i], prefix[i]);
}
__global__ void build_cinfo_ranks_kernel(const int32_t *first_qp, int64_t n, const uint4 *qinfo, uint4 *cinfo)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t qp = (uint32_t)first_qp[r];
    if (qp) store_cinfo_pair(cinfo, (uint32_t)r, qp, qinfo);
}
cudaError_t launch_build_prk_cinfo(const uint32_t *presence, const uint32_t *prefix, int64_t nwords, uint2 *prk,
                                   const int32_t *first_qp, int64_t n_ranks, const uint4 *qinfo, uint4 *cinfo,
                                   cudaStream_t st)
{
    build_prk_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(presence, prefix, nwords, prk);
    if (n_ranks > 0)
        build_cinfo_ranks_kernel<<<(unsigned)((n_ranks + 255) / 256), 256, 0, st>>>(first_qp, n_ranks, qinfo, cinfo);
    return cudaGetLastError();
}
// parity tap: hashtable[] reconstructed from the compact table
__global__ void rebuild_hashtable_kernel(const DevQuery q, int64_t hashsize, int32_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hashsize) out[i] = mb_cell(q, (uint32_t)i);
}
cudaError_t launch_rebuild_hashtable(const DevQuery &q, int64_t hashsize, int32_t *out, cudaStream_t st)
{
    rebuild_hashtable_kernel<<<(unsigned)((hashsize + 255) / 256), 256, 0, st>>>(q, hashsize, out);
    return cudaGetLastError();
}

cudaError_t launch_popc(const uint32_t *presence, int64_t nwords, uint32_t *counts, cudaStream_t st)
{
    popc_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(presence, nwords, counts);
    return cudaGetLastError();
}
cudaError_t launch_build_compact(const int32_t *hashtable, const uint32_t *presence, const uint32_t *prefix,
                                 int64_t nwords, uint2 *prk, const uint4 *qinfo, uint4 *cinfo, cudaStream_t st)
{
    build_compact_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(hashtable, presence, prefix, nwords, prk, qinfo, cinfo);
    return cudaGetLastError();
}

cudaError_t launch_build_presence(const int32_t *hashtable, int64_t hashsize, uint32_t *presence, cudaStream_t st)
{
    const int64_t warps = (hashsize + 1023) / 1024;
    const int64_t blocks = (warps * 32 + 255) / 256;
    build_presence_kernel<<<(unsigned)blocks, 256, 0, st>>>(hashtable, hashsize, presence);
    return cudaGetLastError();
}
cudaError_t launch_build_qpk(const uint8_t *query_start, int32_t concat_len, uint2 *qpk, int64_t nwords, cudaStream_t st)
{
    build_qpk_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(query_start, concat_len, qpk, nwords);
    return cudaGetLastError();
}

cudaError_t launch_scan(const DevQuery &q, const ScanLaunch &s, cudaStream_t st)
{
    if (s.total_pos <= 0) return cudaSuccess;
    int64_t blocks = (s.total_pos + POS_PER_BLOCK - 1) / POS_PER_BLOCK;
    // compile-time stride of the consecutive-position loader (the megablast defaults); BN_NO_CONSEC: test switch
    const int cstep = (!getenv("BN_NO_CONSEC") && (q.scan_step == 17 || q.scan_step == 18)) ? q.scan_step : 0;
    if (q.lut_type == 0 && q.prk != nullptr && q.lut_word_length <= 13 && q.filt != nullptr) {
        // small table: shared-memory filter, one persistent CTA per SM
        static int n_sm = 0, smem_max = 0;
        if (!n_sm) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            const int lim = smem_max - 1024;
            cudaFuncSetAttribute(scan_kernel_filtered<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            cudaFuncSetAttribute(scan_kernel_filtered<false, 17>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            cudaFuncSetAttribute(scan_kernel_filtered<false, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            cudaFuncSetAttribute(scan_kernel_filtered<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
        }
        const bool half = getenv("BN_FILT_HALF") != nullptr;        // test switch: the folded 2^19-bit map
        int flog = 0, maxc = 0;
        size_t smem = 0;
        for (int fl = half ? FILT_LOG2 - 1 : FILT_LOG2; fl >= FILT_LOG2 - 1 && !flog; fl--)
            for (int mc = 128; mc >= 64 && !flog; mc >>= 1) {
                const size_t need = ((size_t)1 << (fl - 3)) +
                                    (size_t)FILT_GROUPS * 2 * ((size_t)s.tile_cap + sizeof(int32_t) * filt_ct_ints(mc));
                if (need <= (size_t)smem_max - 1024) { flog = fl; maxc = mc; smem = need; }
            }
        if (flog) {
            const int64_t ctas = (blocks + FILT_GROUPS - 1) / FILT_GROUPS;
            const unsigned grid = (unsigned)(ctas < n_sm ? ctas : n_sm);
            const unsigned nt = FILT_GROUPS * SCAN_THREADS;
            if (s.direct_filter && !s.raw_pairs) scan_kernel_filtered<true, 0><<<grid, nt, smem, st>>>(q, s, flog, maxc);
            else if (cstep == 17) scan_kernel_filtered<false, 17><<<grid, nt, smem, st>>>(q, s, flog, maxc);
            else if (cstep == 18) scan_kernel_filtered<false, 18><<<grid, nt, smem, st>>>(q, s, flog, maxc);
            else scan_kernel_filtered<false, 0><<<grid, nt, smem, st>>>(q, s, flog, maxc);
            return cudaGetLastError();
        }
    }
    if (q.lut_type == 0 && q.prk != nullptr && q.lut_word_length <= 13) {
        const size_t smem = (size_t)s.tile_cap + sizeof(uint2) * POS_PER_BLOCK;
        if (s.direct_filter && !s.raw_pairs)
            scan_kernel_staged<true, 0><<<(unsigned)blocks, SCAN_THREADS, smem + sizeof(DirectQueue) * (SCAN_THREADS / 32), st>>>(q, s);
        else if (q.psig != nullptr && !s.raw_pairs) {           // dense-signature probe
            if (cstep == 17) scan_kernel_staged<false, 17, true><<<(unsigned)blocks, SCAN_THREADS, smem, st>>>(q, s);
            else if (cstep == 18) scan_kernel_staged<false, 18, true><<<(unsigned)blocks, SCAN_THREADS, smem, st>>>(q, s);
            else scan_kernel_staged<false, 0, true><<<(unsigned)blocks, SCAN_THREADS, smem, st>>>(q, s);
        }
        else if (cstep == 17) scan_kernel_staged<false, 17><<<(unsigned)blocks, SCAN_THREADS, smem, st>>>(q, s);
        else if (cstep == 18) scan_kernel_staged<false, 18><<<(unsigned)blocks, SCAN_THREADS, smem, st>>>(q, s);
        else scan_kernel_staged<false, 0><<<(unsigned)blocks, SCAN_THREADS, smem, st>>>(q, s);
    }
    else
        scan_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(q, s);
    return cudaGetLastError();
}

}  // namespace bn
