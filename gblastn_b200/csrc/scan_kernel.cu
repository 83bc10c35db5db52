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
// positions of the volume-wide position space (prefix sums in the chunk table); each thread forms its
// lookup words from two aligned 32-bit loads + a funnel shift, probes an exact presence bitmap
// (L2-resident, 1 bit per table cell), walks the chain and runs the mini-extension in registers.
// Survivors (a tiny fraction of lookup hits) are appended with warp-aggregated atomics together
// with a 64-bit key = (global position, chain rank) that restores the reference's emission order.
#include "bn_device.cuh"

namespace bn {

constexpr int SCAN_THREADS = 256;
constexpr int POS_PER_THREAD = 4;
constexpr int POS_PER_BLOCK = SCAN_THREADS * POS_PER_THREAD;

int scan_positions_per_block() { return POS_PER_BLOCK; }

__device__ __forceinline__ uint32_t load_window(const uint8_t *packed, int64_t byte)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(packed + (byte & ~int64_t(3)));
    uint32_t a = __byte_perm(__ldg(w), 0, 0x0123);       // big-endian view of the first word
    uint32_t b = __byte_perm(__ldg(w + 1), 0, 0x0123);
    return __funnelshift_l(b, a, (uint32_t)(byte & 3) * 8);
}

__device__ __forceinline__ void emit_hit(const ScanLaunch &s, uint32_t chunk, uint32_t p, int64_t g,
                                         uint32_t rank, int32_t q_off, int32_t s_off)
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
        s.hits[slot] = h;
        s.keys[slot] = ((uint64_t)g << 24) | (uint64_t)(rank & 0xFFFFFFu);
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
        const int32_t p = (int32_t)(g - ch.pos_prefix) * step;
        const uint8_t *S = s.packed + ch.byte_off;
        const uint32_t window = load_window(s.packed, ch.byte_off + (p >> 2));
        const uint32_t idx = (window >> (2 * (16 - ((p & 3) + lut)))) & q.hash_mask;

        if (q.lut_type == 0) {
            if (!((__ldg(&q.presence[idx >> 5]) >> (idx & 31)) & 1u)) continue;
            int32_t qp = __ldg(&q.hashtable[idx]);
            uint32_t rank = 0;
            while (qp) {
                ++my_lookup_hits;
                int32_t qo, so;
                if (s.raw_pairs) emit_hit(s, (uint32_t)lo, (uint32_t)p, g, rank, qp - 1, p);
                else if (mini_extend_mb(q, S, ch.len, qp - 1, p, qo, so))
                    emit_hit(s, (uint32_t)lo, (uint32_t)p, g, rank, qo, so);
                ++rank;
                qp = __ldg(&q.next_pos[qp]);
            }
        } else {
            int32_t v = __ldg(&q.backbone[idx]);
            if (v == -1) continue;
            uint32_t rank = 0;
            int32_t src = 0;
            if (v < 0) { src = -v; v = __ldg(&q.overflow[src++]); }
            do {
                ++my_lookup_hits;
                int32_t qo, so;
                if (s.raw_pairs) emit_hit(s, (uint32_t)lo, (uint32_t)p, g, rank, v, p);
                else if (mini_extend_small(q, S, ch.len, v, p, qo, so))
                    emit_hit(s, (uint32_t)lo, (uint32_t)p, g, rank, qo, so);
                ++rank;
                v = src ? (int32_t)__ldg(&q.overflow[src++]) : -1;
            } while (v >= 0);
        }
    }
    // one atomic per warp for the lookup-hit statistic (BlastUngappedStats.lookup_hits)
    for (int o = 16; o > 0; o >>= 1) my_lookup_hits += __shfl_down_sync(0xffffffffu, my_lookup_hits, o);
    if ((threadIdx.x & 31) == 0 && my_lookup_hits) atomicAdd(&s.counters[1], my_lookup_hits);
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
    scan_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(q, s);
    return cudaGetLastError();
}

}  // namespace bn
