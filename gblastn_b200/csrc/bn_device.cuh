// bn_device.cuh — device-side views shared by the kernels of the blastn hot path.
//
// Memory layout in HBM:
//   volume   : packed ncbi2na bytes of every sequence back to back (+>=16 pad bytes), resident for
//              the lifetime of the volume handle; per-sequence byte offset / length tables.
//   chunks   : per (volume, query-batch) table of subject chunks (the reference's
//              s_GetNextSubjectChunk split, core/blast_engine.c:220-301) with the prefix sum of
//              scan positions, so one launch covers a whole volume (SURVEY.md §7 hard part 6).
//   query    : blastna bytes with sentinels, context table, lookup-table arrays exactly as the
//              reference lays them out, replicated per device.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bn {

struct DevContext {
    int32_t query_offset, query_length, query_index, frame;
    int32_t x_dropoff, cutoff_score, reduced_cutoff, gapped_cutoff;
};

struct DevChunk {
    int64_t byte_off;      // byte offset of the chunk's first base in the volume
    int64_t pos_prefix;    // number of scan positions in all earlier chunks
    int32_t len;           // bases in this chunk (subject->length for the word finder)
    int32_t oid;
    int32_t chunk_off;     // base offset of the chunk inside its sequence
    int32_t npos;          // scan positions in this chunk
    int32_t diag_offset;   // BLAST_DiagHash/BLAST_DiagTable ::offset while this chunk is scanned
    int32_t diag_epoch;    // number of container resets (core/blast_extend.c:170-182) before this chunk
    // Scan unit view (database masks, core/masksubj.inl:43-59): the scan kernel walks "units" = one unmasked
    // range of one chunk; without masks a chunk is its own unit (p_first 0, s_range len, parent = its index).
    int32_t p_first;       // first scan position of the unit (range.left + word - lut)
    int32_t s_range;       // right end of the unit's range: bound of the mini-extension and of s_TypeOfWord
    int32_t parent;        // index of the chunk in the chunk table (what a seed hit records)
    // chunk view: its unmasked ranges in the range table (n_ranges == 0: not masked)
    int32_t range_first, n_ranges;
    int32_t pad;
};
static_assert(sizeof(DevChunk) == 64, "two sectors");

struct DevQuery {
    const uint8_t *query;        // query->sequence (byte before it is the leading sentinel)
    int32_t concat_len;
    const DevContext *ctx;
    int32_t num_contexts;
    int32_t lut_type, word_length, lut_word_length, scan_step;
    uint32_t hash_mask;
    const int32_t *next_pos;               // MB chain links (the 4^lut hashtable itself is not kept: mb_cell)
    const uint2 *prk;                      // MB compact table: {presence word, rank of its first occupied cell}
    const uint4 *cinfo;                    // MB compact table: per occupied cell (in cell order) TWO entries = the qinfo
                                           // of its first and second chain element, .x = {qp, bit 31: chain continues}
    const uint4 *qinfo;                    // MB: per query position {next_pos, 16 bases left, 16 right, ambiguity}
    const uint32_t *sig;                   // MB: per occupied cell (by rank) the 4 query bases on either side of its first chain
                                           // element's lookup word: bits 0-7 left, 8-15 right, bit 16 = the chain has more elements
    const uint16_t *psig;                  // MB: the same signature per table CELL (4^lut entries, dense): bit 15 = cell occupied, bit 14 =
                                           // its chain has more elements, bits 8-13 = the 3 query bases left of the first element's lookup
                                           // word, bits 0-7 = the 4 right of it.  ONE gather per scan position answers "occupied?" and
                                           // "can the mini-extension succeed?" (scan_kernel_staged<.., DENSE>); NULL: prk + sig are used
    const uint32_t *filt;                  // MB, small batches only: FILT_BITS-bit hashed presence filter (filt_hash of every occupied
                                           // cell), kept in shared memory by scan_kernel_filtered; NULL otherwise
    const int16_t *backbone, *overflow;    // SmallNa
    const int4 *na_cells;                  // eNaLookupTable: thick backbone {num_used, entries[3] | overflow_cursor}
    const int32_t *na_overflow;
    int32_t has_locations;                 // lut->masked_locations != NULL
    int32_t container_type, window_size, scan_range;
    const uint2 *qpk;                      // 2-bit packed query windows: .x bases, .y ambiguity (see qwin)
    const int32_t *score_table;            // 256
    const int32_t *matrix;                 // 16 x 16
    int32_t gap_algo, reward, penalty, gap_open, gap_extend, gap_x_dropoff;
};

// A word hit that survived the mini-extension (input of the diagonal stage).
struct SeedHit {
    uint32_t chunk;      // index into the chunk table
    uint32_t scan_pos;   // subject offset of the lookup word inside the chunk
    uint32_t q_off;      // query offset after the left shift (q_offset - ext_left)
    uint32_t s_off;      // subject offset after the left shift
};

struct DevInitHit {      // == BnInitHit + ordering info
    int32_t chunk, q_off, s_off, q_start, s_start, length, score;
    uint32_t order;      // rank of the seed in global emission order (tie-break of the stable sort)
};

struct DevGapResult {
    int32_t q_start, q_stop, s_start, s_stop, score, q_seed, s_seed, status;  // status 1 = scratch overflow, 2 = long alignment (DP row budget)
};

// NCBI2NA_UNPACK_BASE (inc-core/blast_util.h:52-55)
__device__ __forceinline__ int sbase(const uint8_t *s, int32_t pos)
{
    return (__ldg(s + (pos >> 2)) >> (6 - 2 * (pos & 3))) & 3;
}

// 32-bit big-endian window starting at byte `b` of a byte stream (unaligned).
__device__ __forceinline__ uint32_t be32(const uint8_t *p)
{
    return ((uint32_t)__ldg(p) << 24) | ((uint32_t)__ldg(p + 1) << 16) |
           ((uint32_t)__ldg(p + 2) << 8) | (uint32_t)__ldg(p + 3);
}

// ---- 16-base windows -------------------------------------------------------------------------
// Both sequences are compared 16 bases at a time.  A window is a 32-bit word with base j of the
// window in bits 31-2j..30-2j (the subject's own ncbi2na bit order, inc-core/blast_util.h:52-55).
//
// Query: qpk[i] covers concatenated-query positions 16*(i-1) .. 16*(i-1)+15 (one leading pad word
// so that the sentinel at -1 and reverse windows are addressable); .x = 2-bit bases, .y = 01 in the
// base's bit pair when the blastna code is >= 4 (ambiguity / sentinel) or the position lies outside
// the buffer.  Built on the host at bn_query_load from the bytes the reference uses.
__device__ __forceinline__ void qwin(const DevQuery &q, int32_t pos, uint32_t &bases, uint32_t &amb)
{
    const int32_t a = pos + 16;
    const uint2 w0 = __ldg(&q.qpk[a >> 4]);
    const uint2 w1 = __ldg(&q.qpk[(a >> 4) + 1]);
    const uint32_t sh = (uint32_t)(a & 15) * 2;
    bases = __funnelshift_l(w1.x, w0.x, sh);
    amb = __funnelshift_l(w1.y, w0.y, sh);
}

// Subject: `packed` is the volume base (4-byte aligned, >= 64 readable bytes in front of it),
// abs_base = 4 * byte offset of the chunk + position inside the chunk (may be negative).
__device__ __forceinline__ uint32_t swin(const uint8_t *packed, int64_t abs_base)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(packed) + (abs_base >> 4);
    const uint32_t a = __byte_perm(__ldg(w), 0, 0x0123);
    const uint32_t b = __byte_perm(__ldg(w + 1), 0, 0x0123);
    return __funnelshift_l(b, a, (uint32_t)(abs_base & 15) * 2);
}

// one bit (the low bit of each pair) per mismatching base of two windows
__device__ __forceinline__ uint32_t mismatch_bits(uint32_t qb, uint32_t qamb, uint32_t sb)
{
    const uint32_t x = qb ^ sb;
    return ((x | (x >> 1)) & 0x55555555u) | qamb;
}

// number of equal bases query[qpos + t] == subject[spos + t], t = 0 .. n-1, before the first mismatch
__device__ __forceinline__ int32_t match_run_fwd(const DevQuery &q, const uint8_t *packed, int32_t qpos,
                                                 int64_t spos, int32_t n)
{
    int32_t cnt = 0;
    while (cnt < n) {
        uint32_t qb, qa;
        qwin(q, qpos + cnt, qb, qa);
        const uint32_t m = mismatch_bits(qb, qa, swin(packed, spos + cnt));
        if (m) return min(cnt + (__clz(m) >> 1), n);
        cnt += 16;
    }
    return n;
}

// number of equal bases query[qend - 1 - t] == subject[send - 1 - t], t = 0 .. n-1
__device__ __forceinline__ int32_t match_run_rev(const DevQuery &q, const uint8_t *packed, int32_t qend,
                                                 int64_t send, int32_t n)
{
    int32_t cnt = 0;
    while (cnt < n) {
        uint32_t qb, qa;
        qwin(q, qend - cnt - 16, qb, qa);
        const uint32_t m = mismatch_bits(qb, qa, swin(packed, send - cnt - 16));
        if (m) return min(cnt + ((__ffs(m) - 1) >> 1), n);
        cnt += 16;
    }
    return n;
}

// Hashed presence filter of a small MB table (scan_kernel_filtered): cell idx -> bit filt_hash(idx) of a 2^20-bit map;
// exact for lut <= 10.  A 2^19-bit map is the OR of the two halves (hash & (2^19 - 1)).
constexpr int FILT_LOG2 = 20;
constexpr uint32_t FILT_BITS = 1u << FILT_LOG2;
__host__ __device__ __forceinline__ uint32_t filt_hash(uint32_t idx) { return (idx ^ (idx >> FILT_LOG2)) & (FILT_BITS - 1u); }

// cinfo / qinfo .x: bits 0-29 a 1-based query position (cinfo: the chain element itself; qinfo: the next one), bit 31
// (cinfo) "the chain continues", bit 30 "the query position in front of this element is in the lookup table too"
constexpr uint32_t QP_MASK = 0x3fffffffu;
constexpr uint32_t PREV_INDEXED = 0x40000000u;

// BSearchContextInfo (core/blast_query_info.c:220-236)
__device__ __forceinline__ int32_t ctx_search(const DevQuery &q, int32_t n)
{
    int32_t lo = 0, hi = q.num_contexts;
    while (lo < hi - 1) {
        int32_t m = (lo + hi) / 2;
        if (__ldg(&q.ctx[m].query_offset) > n) hi = m; else lo = m;
    }
    return lo;
}

// BSearchContextInfo (core/blast_query_info.c:220-236) with 32 pivots per round: the largest
// context index whose query_offset <= n.  Warp-uniform result.
__device__ __forceinline__ int32_t ctx_search_warp(const DevQuery &q, int32_t n, int lane)
{
    int32_t lo = 0, hi = q.num_contexts;
    while (hi - lo > 1) {
        const int32_t step = (hi - lo + 31) >> 5;
        const int32_t piv = lo + lane * step;
        const bool ok = piv < hi && __ldg(&q.ctx[piv].query_offset) <= n;
        const unsigned m = __ballot_sync(0xffffffffu, ok) | 1u;      // pivot 0 (= lo) always qualifies
        const int top = 31 - __clz(m);
        const int32_t nlo = lo + top * step;
        hi = min(hi, nlo + step);
        lo = nlo;
    }
    return lo;
}

// hashtable[idx] of BlastMBLookupTable (inc-core/blast_nalookup.h:236) recovered from the compact
// table: 0 when the cell is empty, else the 1-based query position that heads the cell's chain.
__device__ __forceinline__ int32_t mb_cell(const DevQuery &q, uint32_t idx)
{
    const uint2 w = __ldg(&q.prk[idx >> 5]);
    const uint32_t bit = idx & 31u;
    if (!((w.x >> bit) & 1u)) return 0;
    return (int32_t)(__ldg(&q.cinfo[2 * (size_t)(w.y + (uint32_t)__popc(w.x & ((1u << bit) - 1u)))].x) & QP_MASK);
}

// ---- launchers implemented in the .cu files ----------------------------------------------------
// Per scan block (scan_positions_per_block() consecutive positions): the slice of the volume it stages
// and the chunks it spans; built on the host with the chunk table.  32 bytes = one sector.
struct ScanBlockDesc {
    int64_t tile_lo;      // first byte of the staged slice (16-byte aligned, may reach into the front pad)
    int32_t bytes;        // slice size (multiple of 16)
    int32_t c_lo, c_hi;   // first / last chunk with a position in the block
    int32_t staged;       // 0: slice does not fit the tile (or too many chunks) -> direct-load path
    int32_t pad0, pad1;
};
static_assert(sizeof(ScanBlockDesc) == 32, "one sector");

struct ScanLaunch {
    const uint8_t *packed;
    const DevChunk *chunks;
    int32_t n_chunks;
    int64_t total_pos;
    SeedHit *hits;            // capacity entries
    uint64_t *keys;           // sort key per hit
    unsigned long long *counters;   // [0] = #survivors, [1] = #lookup hits
    int64_t capacity;
    const int32_t *block_chunk;   // first chunk of every BLOCK_POS-sized slice of positions (generic kernel)
    const ScanBlockDesc *block_desc;  // staged kernel
    int32_t raw_pairs;            // 1: emit every lookup hit (q_off, scan_pos) without mini-extension (scan tap)
    int32_t gbits;                // bits of the global position in the sort key
    int32_t diag_array_length;    // eDiagArray: cells (power of two)
    int32_t tile_cap;             // staged kernel: bytes of shared memory for the subject slice
    uint32_t *bucket_count;       // optional: survivors per diagonal-hash bucket (group_sort.cu)
    uint64_t *bucket_keys;        // with bucket_count: bucket_cap keys (position << 24 | emission slot) per bucket
    int32_t bucket_cap;
    int32_t one_group;            // 1: every survivor gets group 0 (serial replay, off-diagonal two-hit search)
    // direct filter (lut == word, one-hit mode, unmasked volume): drop hits whose ungapped extension certainly
    // stays below the cutoff; uni_*: the cutoffs when every context has the same ones (uni_ok), else per context
    int32_t direct_filter;
    int32_t direct_dense;         // queue the filter evaluations and run them 32 at a time (test switch BN_NO_DIRECT_DENSE: inline)
    int32_t uni_ok, uni_x, uni_cutoff, uni_reduced;
};
cudaError_t launch_scan(const DevQuery &q, const ScanLaunch &s, cudaStream_t st);
int scan_positions_per_block();
int scan_tile_cap(int scan_step, int word_length);
int scan_max_block_chunks();
int scan_tile_margin();
// radix_sort.cu: the device-wide sort and prefix sums of the path (no library kernels)
size_t prefix_sum_temp_bytes(int64_t n);
cudaError_t prefix_sum_u32(const uint32_t *in, uint32_t *out, int64_t n, bool inclusive, void *temp, cudaStream_t st);
size_t radix_sort_temp_bytes(int64_t n);
cudaError_t radix_sort_hits(uint64_t *ka, uint64_t *kb, SeedHit *va, SeedHit *vb, int64_t n, int end_bit, void *temp, bool *in_b,
                            int64_t *n_launches, cudaStream_t st);
cudaError_t radix_sort_u32(uint32_t *ka, uint32_t *kb, uint32_t *va, uint32_t *vb, int64_t n, int end_bit, void *temp, bool *in_b,
                           int64_t *n_launches, cudaStream_t st);
cudaError_t launch_build_filter(const uint32_t *presence, int64_t nwords, uint32_t *filt, cudaStream_t st);
cudaError_t launch_build_sig(const uint4 *cinfo, int64_t n_ranks, uint32_t *sig, cudaStream_t st);
cudaError_t launch_build_psig(const uint2 *prk, const uint32_t *sig, int64_t n_cells, uint16_t *psig, cudaStream_t st);
cudaError_t launch_build_qinfo(const DevQuery &q, const int32_t *next_pos, int32_t concat_len, const int32_t *heads,
                               int64_t n_heads, uint32_t *indexed_scratch, uint4 *qinfo, cudaStream_t st);

// Outcome of s_TypeOfWord + ungapped extension of one word hit (extend_kernel.cu), 32 bytes.
struct SpecResult {
    int32_t status;          // SPEC_* (0 = not computed)
    int32_t q_off, s_off;    // after s_TypeOfWord's shift
    int32_t extended;
    int32_t q_start, s_start, length, score;
};

struct ExtendLaunch {
    const uint8_t *packed;
    const DevChunk *chunks;
    const int2 *ranges;           // unmasked [left, right) ranges of masked chunks (DevChunk::range_first)
    const SeedHit *hits;          // sorted by (group, global scan position)
    int32_t *cells;               // hash: 4 ints per cell, one region per group (same offsets as hits)
    DevInitHit *init;
    SpecResult *spec;             // per hit: outcome of the speculative extension
    uint32_t *leaders;            // indices of the hits extended speculatively
    unsigned long long *counters; // [2] = #init hits, [3] = #extended, [4] = #groups, [5] = #leaders, [6] = fast path refused
    int64_t init_capacity;
    int32_t n_from_device;        // 1: the number of hits is counters[0] (no host round trip), kernels idle if counters[6]
    // blastn mode with plain reward / penalty tables (Query::direct_ok): the speculative pass extends one leader per LANE
    // with the closed-form scores (extend_leaders_kernel); uni_*: the cutoffs when every context has the same ones
    int32_t scalar_ok, uni_ok, uni_x, uni_cutoff, uni_reduced;
};
cudaError_t launch_extend_groups(const DevQuery &q, const ExtendLaunch &e, const uint64_t *keys,
                                 uint32_t *heads, int64_t n_hits, int gbits, cudaStream_t st);
cudaError_t launch_extend_grouped(const DevQuery &q, const ExtendLaunch &e, const uint64_t *keys,
                                  const uint32_t *heads, int gbits, cudaStream_t st);
// two-hit mode with scan_range > 0: one warp replays ALL hits in emission order (hits sorted by global position)
cudaError_t launch_extend_serial(const DevQuery &q, const ExtendLaunch &e, const uint64_t *keys, int64_t n_hits,
                                 int gbits, int32_t diag_array_length, cudaStream_t st);
int64_t extend_serial_cells(int64_t n_hits, bool is_hash, int32_t diag_array_length);

// Device-side grouping of the seed hits by diagonal-hash bucket (group_sort.cu).
struct BucketLaunch {
    const SeedHit *hits_in;       // scan output, emission order by slot
    const uint32_t *bucket_count; // 512, filled by the scan kernel
    const uint64_t *keys_tmp;     // 512 regions of group_sort_bucket_cap() keys, filled by the scan kernel
    SeedHit *hits_out;            // grouped + ordered hits
    uint64_t *keys_out;           // (bucket << gbits) | global scan position, ordered
    uint32_t *heads, *leaders;
    SpecResult *spec;
    unsigned long long *counters; // [0] = #hits (in), [4] = #groups, [5] = #leaders, [6] = fast path refused (out)
    int64_t n_limit;              // most hits the fast path takes (buffer sizes, <= 2^24)
    int32_t gbits, spec_enabled;
};
cudaError_t launch_bucket_group(const BucketLaunch &L, cudaStream_t st);
cudaError_t launch_mirror_results(const DevInitHit *init, const DevGapResult *gap, const unsigned long long *counters,
                                  int64_t cap, DevInitHit *h_init, DevGapResult *h_gap, unsigned long long *h_counters,
                                  cudaStream_t st);
int group_sort_buckets();
int group_sort_bucket_cap();

// Derived costs of BLAST_AffineGreedyAlign (core/greedy_align.c:792-842): odd rewards double every
// score, the three operation costs are divided by their gcd (BLAST_Gdb3 core/ncbi_math.c:427).
struct AffineCosts {
    int32_t match, mismatch, xdrop;           // after the doubling
    int32_t op_cost, gap_open, gap_extend;    // after the gcd division
    int32_t common_factor, max_penalty, xdrop_offset;
};
__host__ __device__ inline int32_t bn_gcd(int32_t a, int32_t b)
{
    if (b < 0) b = -b;
    if (b > a) { const int32_t c = a; a = b; b = c; }
    while (b != 0) { const int32_t c = a % b; a = b; b = c; }
    return a;
}
__host__ __device__ inline AffineCosts affine_costs(int32_t reward, int32_t penalty, int32_t gap_open,
                                                    int32_t gap_extend, int32_t xdrop)
{
    AffineCosts c;
    c.match = reward; c.mismatch = -penalty; c.xdrop = xdrop;
    int32_t go = gap_open, ge = gap_extend;
    if (c.match % 2 == 1) { c.match *= 2; c.mismatch *= 2; c.xdrop *= 2; go *= 2; ge *= 2; }
    c.op_cost = c.match + c.mismatch;
    c.gap_open = go;
    c.gap_extend = ge + c.match / 2;
    const int32_t g = (c.gap_open == 0) ? bn_gcd(c.op_cost, c.gap_extend)
                                        : bn_gcd(c.op_cost, bn_gcd(c.gap_open, c.gap_extend));
    if (g > 1) { c.op_cost /= g; c.gap_open /= g; c.gap_extend /= g; }
    c.common_factor = g;
    const int32_t goe = c.gap_open + c.gap_extend;
    c.max_penalty = c.op_cost > goe ? c.op_cost : goe;
    c.xdrop_offset = (c.xdrop + c.match / 2) / g + 1;
    return c;
}
// ints of per-thread scratch the affine greedy needs for diagonal half-width D
__host__ __device__ inline int64_t affine_scratch_ints(const AffineCosts &c, int32_t D)
{
    const int64_t dmax = (int64_t)D * c.gap_extend;
    return 3 * (int64_t)(c.max_penalty + 1) * (2 * (int64_t)D + 6) + 2 * (dmax + 1 + c.max_penalty) +
           (dmax + 2 + c.xdrop_offset) + 8;
}

struct GappedLaunch {
    const uint8_t *packed;
    const DevChunk *chunks;
    const DevInitHit *init;
    const unsigned long long *n_init;   // device counter
    int64_t max_init;
    DevGapResult *out;
    int32_t *scratch;             // per-thread scratch
    int64_t scratch_ints_per_thread;
    int32_t tier_d;               // greedy: max distance this tier can hold; dp: ring capacity
    const int32_t *todo;          // optional list of init indices (tier 2); nullptr = all
    int32_t n_todo;
    int32_t grid_blocks;          // 0 = default persistent grid
    int32_t dp_max_rows;          // DP tier 1: rows (per direction) a single thread may walk before it reports status 2 (0 = no limit)
    int32_t dp_smem_ring;         // DP: 1 = tier-1 rings in shared memory (no global scratch), 2 = the same with 16-bit cells
                                  // (needs dp_max_rows), 0 = global ring of tier_d cells (power of two)
    unsigned long long *work_counter;   // optional, zeroed by the caller: extensions are handed out one by one
};
cudaError_t launch_gapped(const DevQuery &q, const GappedLaunch &g, cudaStream_t st);
cudaError_t launch_greedy_warp(const DevQuery &q, const GappedLaunch &g, int warps_per_block, int blocks,
                               bool use_smem, cudaStream_t st);
int gapped_threads();
int gapped_threads_per_block();
int gapped_dp_smem_blocks();
int gapped_dp_ring16_blocks();
cudaError_t launch_gapped_warp(const DevQuery &q, const GappedLaunch &g, int blocks, cudaStream_t st);
// long alignments, one thread per (extension, direction): g.todo / g.n_todo list them; halves = 2 * n_todo int2 scratch
cudaError_t launch_gapped_long(const DevQuery &q, const GappedLaunch &g, int2 *halves, cudaStream_t st);
int gapped_warp_per_block();

// ---- triage of the speculative gapped extensions (triage_kernel.cu) ----------------------------------
struct TriageLaunch {
    const DevInitHit *init;
    const DevGapResult *gap;
    const unsigned long long *n_init;   // device counter
    int64_t max_init;
    int32_t *ctx_of;                    // scratch: context per init-HSP, bit 31 = winner
    DevInitHit *sel_init;               // winners, then the losers the host has to replay in order
    DevGapResult *sel_gap;
    int32_t *sel_idx;                   // index of the selected record among all init-HSPs
    int32_t *sel_ctx;                   // per winner: the next winner (+1) of its (chunk, context), 0 = end of the chain
    int64_t sel_cap;
    unsigned long long *tcount;         // [0] winners, [1] undecided losers, [2] counted losers (zeroed by the caller)
    uint2 *table;                       // per (chunk, context), zeroed by the caller: .x bit 31 = has a winner, .y = newest
                                        // winner + 1 (head of the chain through sel_ctx)
    int32_t n_ctx;
};
cudaError_t launch_triage(const DevQuery &q, const TriageLaunch &t, cudaStream_t st);
// Long gapped extensions made in rounds (triage_kernel.cu): per (chunk, context) the best pending one first, pending
// ones inside a box made so far set aside (DevGapResult::status 3 = not computed).
struct LongRounds {
    const DevInitHit *init;
    DevGapResult *gap;
    const int32_t *todo;                // the long extensions (indices of init-HSPs); w = position in this list
    int32_t n_todo;
    int32_t *state, *ctx_w, *chain_next;        // per w: 0 pending / 1 made / 2 in this round / 3 set aside; context; chain link
    unsigned long long *best;           // per (chunk, context): best pending key of the round (all ones = none)
    uint32_t *chain_head;               // per (chunk, context): newest saved box (w + 1), zeroed by the caller
    int32_t *round_list, *round_w;      // this round's extensions: init index, w
    unsigned long long *round_count;
    int32_t n_ctx, min_diag_separation;
    int32_t set_aside_all;              // test switch: once a (chunk, context) has a box, set every other pending one aside
};
cudaError_t launch_long_prepare(const DevQuery &q, const LongRounds &r, cudaStream_t st);
cudaError_t launch_long_select(const DevQuery &q, const LongRounds &r, cudaStream_t st);
cudaError_t launch_long_commit(const DevQuery &q, const LongRounds &r, int32_t n_round, cudaStream_t st);
cudaError_t launch_collect_status(const DevGapResult *gap, const unsigned long long *n_init, int64_t max_init, int32_t want,
                                  int32_t *todo, unsigned long long *count, cudaStream_t st);

// ---- gapped alignment with traceback (traceback_kernel.cu) -----------------------------------------
struct DevTracebackItem {
    int64_t byte_off;            // byte offset of the subject sequence in the volume
    int32_t context;             // query context
    int32_t s_shift, s_length;   // AdjustSubjectRange: subject = sequence + s_shift, s_length bases
    int32_t q_start, s_start;    // start point (s_start relative to s_shift)
    int32_t pad;
    int32_t amb_first, amb_n;    // ambiguity runs of the subject sequence in the volume's run table (amb_n == 0: none)
    int32_t pad2;
};
struct DevTracebackDir {         // one direction of one alignment
    int32_t score, a_off, b_off; // best score and the query / subject extents of ALIGN_EX
    int32_t n_ops;               // run-length edit operations in walk order (far end first)
    long long ops_off;           // index of the first one in TracebackLaunch::ops
    int32_t status;              // 0 ok, 1 band wider than the ring, 3 arena exhausted, 4 ops buffer exhausted
    int32_t ran;                 // 0: this direction is not run (start point on the last base)
    int32_t pad;
};
struct TracebackLaunch {
    const uint8_t *packed;
    const DevTracebackItem *items;
    int64_t n;
    int32_t x_dropoff;           // gap_x_dropoff_final (raw)
    uint8_t *arena;              // script rows, row tables, temporary run lists
    long long arena_bytes;
    unsigned long long *arena_used;
    int2 *ops;                   // {EGapAlignOpType, num}
    long long ops_cap;
    unsigned long long *ops_used;
    DevTracebackDir *out;        // 2 n entries: left, right
    const int4 *amb_runs;        // the volume's ambiguity runs {first base, end, blastna code, 0} (nullptr: none)
    const uint8_t *todo;         // greedy: optional per-item flags (retry of the items whose arena overflowed); nullptr = all
    int2 *wide_ring;             // DP, wide variant: traceback_wide_cells() band cells per warp of the grid in global memory
    uint8_t *wide_pf;            // (nullptr: the shared-memory ring)
};
int traceback_wide_cells();
cudaError_t launch_traceback_dp(const DevQuery &q, const TracebackLaunch &L, int blocks, cudaStream_t st);
struct DevTracebackHsp {         // a preliminary HSP (absolute subject coordinates) about to be traced back
    int64_t byte_off;            // of its subject sequence
    int32_t seq_len, context, q_off, q_end, s_off, s_end, q_gapped_start, s_gapped_start;
    int32_t amb_first, amb_n;
};
cudaError_t launch_traceback_start(const DevQuery &q, const uint8_t *packed, const int4 *amb_runs, const DevTracebackHsp *hsps,
                                   int64_t n, DevTracebackItem *items, cudaStream_t st);
struct DevTracebackPost {        // an HSP after the list logic (absolute subject coordinates), edit script in ops[esp_off ..)
    int64_t byte_off, esp_off;
    int32_t seq_len, context, q_off, q_end, s_off, s_end, score, esp_n, reevaluate, pad;
    int32_t amb_first, amb_n;
};
struct DevTracebackPostOut {
    int32_t deleted, q_off, q_end, s_off, s_end, score, first, last, num_ident, align_length;
};
cudaError_t launch_traceback_reevaluate(const DevQuery &q, const uint8_t *packed, const int4 *amb_runs, const DevTracebackPost *items, int64_t n,
                                        int2 *ops, DevTracebackPostOut *out, cudaStream_t st);
cudaError_t launch_traceback_greedy_warp(const DevQuery &q, const TracebackLaunch &L, int blocks, cudaStream_t st);
cudaError_t launch_traceback_greedy_affine(const DevQuery &q, const TracebackLaunch &L, int blocks, int threads_per_block, cudaStream_t st);
int traceback_warps_per_block();

}  // namespace bn
