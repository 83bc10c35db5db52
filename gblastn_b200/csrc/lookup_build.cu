// lookup_build.cu — megablast lookup-table construction on the device.
//
// Replaces (semantics, not code) the host fill the reference runs once per query batch:
//   BlastMBLookupTableNew     core/blast_nalookup.c:939-1044
//   s_FillContigMBTable       core/blast_nalookup.c:832-937
// Input is what LookupTableWrapInit receives: the concatenated blastna query and the
// `lookup_segments` list (unmasked [left, right] intervals, core/lookup_wrap.c:49-122).
//
// The reference walks the segments in order and, for every word start `index` (1-based) whose
// lut_word_length bases lie inside one segment (of at least word_length bases) and are all
// unambiguous, does   next_pos[index] = hashtable[word]; hashtable[word] = index;
// so a cell's chain lists its positions in DESCENDING order.  Equivalent parallel formulation:
//   1. every position computes (word | invalid marker)                       mb_words_kernel
//   2. one stable radix sort of (word, position)                            radix_sort.cu
//   3. inside a run of equal words the predecessor is next_pos, the last element is the cell's
//      hashtable value; runs in ascending word order ARE the occupied cells in cell order, i.e.
//      the rank space of the scan kernel's compact table                     mb_link_kernel
// The 4^lut-entry hashtable itself is never materialised on this path: the scan kernel works from
// {presence word, rank} + the per-rank first chain element (scan_kernel.cu).
#include <algorithm>

#include "bn_device.cuh"
#include "devmem.h"

namespace bn {

// segmark[i]: bit 0 = base i lies inside a segment that is long enough to be indexed,
//             bit 1 = base i is the first base of such a segment (words must not span two segments)
__global__ void mark_segments_kernel(const int32_t *segs, int32_t n_segs, int32_t word_length, int32_t concat_len,
                                     uint8_t *segmark)
{
    const int32_t sg = blockIdx.x;
    if (sg >= n_segs) return;
    const int32_t left = segs[2 * sg], right = min(segs[2 * sg + 1], concat_len - 1);
    if (left < 0 || word_length > right - left + 1) return;
    for (int32_t i = left + threadIdx.x; i <= right; i += blockDim.x) segmark[i] = (i == left) ? 3 : 1;
}

__global__ void mb_words_kernel(const uint8_t *query, int32_t concat_len, int32_t lut, const uint8_t *segmark,
                                uint32_t *keys, uint32_t *vals)
{
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= concat_len) return;
    bool valid = i + lut <= concat_len;
    uint32_t w = 0;
    if (valid) {
        for (int k = 0; k < lut; k++) {
            const uint32_t b = query[i + k];
            const uint32_t m = segmark[i + k];
            valid = valid && b < 4u && (m & 1u) && (k == 0 || !(m & 2u));
            w = (w << 2) | (b & 3u);
        }
    }
    keys[i] = valid ? w : (1u << (2 * lut));
    vals[i] = (uint32_t)i;
}

// flags[j] = 1 when sorted element j opens a run of equal (valid) words
__global__ void mb_heads_kernel(const uint32_t *keys, int32_t n, uint32_t invalid, uint32_t *flags)
{
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t w = keys[j];
    flags[j] = (w != invalid && (j == 0 || keys[j - 1] != w)) ? 1u : 0u;
}

// runs[j] = inclusive scan of flags => rank of j's word = runs[j] - 1
__global__ void mb_link_kernel(const uint32_t *keys, const uint32_t *vals, const uint32_t *runs, int32_t n,
                               uint32_t invalid, int32_t *next_pos, uint32_t *presence, int32_t *first_qp)
{
    const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t w = keys[j];
    if (w == invalid) return;
    const int32_t qp = (int32_t)vals[j] + 1;                       // 1-based word start
    next_pos[qp] = (j > 0 && keys[j - 1] == w) ? (int32_t)vals[j - 1] + 1 : 0;
    if (j == n - 1 || keys[j + 1] != w) {                          // last of the run = hashtable[w]
        first_qp[runs[j] - 1] = qp;
        atomicOr(&presence[w >> 5], 1u << (w & 31u));
    }
}

struct LookupBuildTemp {
    uint8_t *segmark;
    uint32_t *keys_a, *keys_b, *vals_a, *vals_b, *flags;
};

// Builds next_pos (concat_len + 1 ints, zero-filled here), the presence bitmap (zero-filled here) and
// first_qp[rank].  All temporaries come from the stream-ordered pool.
cudaError_t build_mb_lookup_device(const uint8_t *d_query /* base 0 */, int32_t concat_len, int32_t word_length,
                                   int32_t lut, const int32_t *d_segs, int32_t n_segs, int32_t *d_next_pos,
                                   uint32_t *d_presence, int32_t *d_first_qp, int64_t *n_launches, cudaStream_t st)
{
    if (concat_len <= 0) return cudaSuccess;
    const int64_t hashsize = (int64_t)1 << (2 * lut);
    const int32_t n = concat_len;
    cudaError_t e;
    LookupBuildTemp t{};
    void *cub_tmp = nullptr;
#define LB_TRY(x) do { e = (x); if (e != cudaSuccess) goto done; } while (0)
    LB_TRY(dev_malloc((void **)&t.segmark, (size_t)n, st));
    LB_TRY(dev_malloc((void **)&t.keys_a, (size_t)n * 4, st));
    LB_TRY(dev_malloc((void **)&t.keys_b, (size_t)n * 4, st));
    LB_TRY(dev_malloc((void **)&t.vals_a, (size_t)n * 4, st));
    LB_TRY(dev_malloc((void **)&t.vals_b, (size_t)n * 4, st));
    LB_TRY(dev_malloc((void **)&t.flags, (size_t)n * 4, st));
    LB_TRY(cudaMemsetAsync(t.segmark, 0, (size_t)n, st));
    LB_TRY(cudaMemsetAsync(d_next_pos, 0, ((size_t)concat_len + 1) * 4, st));
    LB_TRY(cudaMemsetAsync(d_presence, 0, (size_t)((hashsize + 31) / 32) * 4, st));
    if (n_segs > 0) {
        mark_segments_kernel<<<n_segs, 128, 0, st>>>(d_segs, n_segs, word_length, concat_len, t.segmark);
        LB_TRY(cudaGetLastError());
    }
    {
        const unsigned blocks = (unsigned)((n + 255) / 256);
        const uint32_t invalid = 1u << (2 * lut);
        mb_words_kernel<<<blocks, 256, 0, st>>>(d_query, concat_len, lut, t.segmark, t.keys_a, t.vals_a);
        LB_TRY(cudaGetLastError());
        // one stable sort of (word, position) = every chain in position order (radix_sort.cu), run numbers by a prefix sum
        LB_TRY(dev_malloc(&cub_tmp, std::max(radix_sort_temp_bytes(n), prefix_sum_temp_bytes(n)), st));
        bool in_b = false;
        int64_t launches = 0;
        LB_TRY(radix_sort_u32(t.keys_a, t.keys_b, t.vals_a, t.vals_b, n, 2 * lut + 1, cub_tmp, &in_b, &launches, st));
        if (!in_b) { std::swap(t.keys_a, t.keys_b); std::swap(t.vals_a, t.vals_b); }
        mb_heads_kernel<<<blocks, 256, 0, st>>>(t.keys_b, n, invalid, t.flags);
        LB_TRY(cudaGetLastError());
        LB_TRY(prefix_sum_u32(t.flags, t.flags, n, true, cub_tmp, st));
        mb_link_kernel<<<blocks, 256, 0, st>>>(t.keys_b, t.vals_b, t.flags, n, invalid, d_next_pos, d_presence, d_first_qp);
        LB_TRY(cudaGetLastError());
        if (n_launches) *n_launches += 4 + launches + 3;
    }
#undef LB_TRY
done:
    {
        void *ptrs[] = {t.segmark, t.keys_a, t.keys_b, t.vals_a, t.vals_b, t.flags, cub_tmp};
        for (void *p : ptrs) if (p) dev_free(p, st);
    }
    return e;
}

}  // namespace bn
