// radix_sort.cu — the device-wide primitives of the path, written here instead of taken from a library:
//   * a stable LSD radix sort of (key, value) pairs, 8 bits per pass: the seed hits of the general word-finder path by
//     (diagonal group << gbits | global scan position) — blastn-mode batches bring millions of them (C3: 5.6 M per
//     pass) —, and the (lookup word, query position) pairs of the device-side table fill (s_FillContigMBTable,
//     core/blast_nalookup.c:1060-1140, restated as a sort in lookup_build.cu);
//   * exclusive / inclusive prefix sums of 32-bit counts (ranks of the occupied table cells, run numbers).
//
// Sort, one pass: every WARP owns RS_WARP_ITEMS consecutive items.
//   histogram   the warp counts its items per digit in shared memory -> hist[digit][warp]
//   scan        exclusive prefix over hist in (digit, warp) order = where each warp's items of each digit go
//   scatter     the warp walks its items again, 32 at a time and in order: lanes with equal digits find each other with
//               MATCH.ANY, rank = number of lower lanes with the digit, the lowest such lane advances the warp's running
//               offset of the digit.  Order inside a digit is the input order: stable.
// Keys and values are read twice and written once per pass (C3's 5.6 M hits of 24 bytes, 5 passes: ~1 ms).
#include "bn_device.cuh"

namespace bn {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WARP_ITEMS = 2048;
constexpr int RS_DIGITS = 256;

// ---- prefix sums ------------------------------------------------------------------------------------------------------
constexpr int PS_THREADS = 1024;
constexpr int PS_PER_THREAD = 4;
constexpr int PS_TILE = PS_THREADS * PS_PER_THREAD;

// exclusive scan of one 32-bit value per thread over the block; total = sum over the block
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t *warp_sums, uint32_t &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += y;
        }
        warp_sums[lane] = s;                       // inclusive over the warps
    }
    __syncthreads();
    total = warp_sums[(blockDim.x >> 5) - 1];
    const uint32_t before = wid ? warp_sums[wid - 1] : 0u;
    __syncthreads();
    return before + x - v;
}

// tile-local scan; tile_sums[b] = the tile's total.  inclusive != 0: out[i] includes in[i].  in == out allowed.
__global__ void __launch_bounds__(PS_THREADS)
ps_tiles_kernel(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *tile_sums, int inclusive)
{
    __shared__ uint32_t warp_sums[32];
    const int64_t base = (int64_t)blockIdx.x * PS_TILE + (int64_t)threadIdx.x * PS_PER_THREAD;
    uint32_t v[PS_PER_THREAD], sum = 0;
#pragma unroll
    for (int k = 0; k < PS_PER_THREAD; k++) { v[k] = base + k < n ? in[base + k] : 0u; sum += v[k]; }
    uint32_t total;
    uint32_t run = block_exclusive(sum, warp_sums, total);
#pragma unroll
    for (int k = 0; k < PS_PER_THREAD; k++) {
        if (base + k < n) out[base + k] = inclusive ? run + v[k] : run;
        run += v[k];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// exclusive scan of the tile totals in place, one block
__global__ void __launch_bounds__(PS_THREADS)
ps_sums_kernel(uint32_t *tile_sums, int64_t n_tiles)
{
    __shared__ uint32_t warp_sums[32];
    uint32_t carry = 0;
    for (int64_t b0 = 0; b0 < n_tiles; b0 += PS_THREADS) {
        const int64_t i = b0 + threadIdx.x;
        const uint32_t v = i < n_tiles ? tile_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive(v, warp_sums, total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(PS_THREADS)
ps_add_kernel(uint32_t *out, int64_t n, const uint32_t *tile_sums)
{
    const uint32_t add = tile_sums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * PS_TILE + (int64_t)threadIdx.x * PS_PER_THREAD;
#pragma unroll
    for (int k = 0; k < PS_PER_THREAD; k++) if (base + k < n) out[base + k] += add;
}

size_t prefix_sum_temp_bytes(int64_t n) { return (size_t)((n + PS_TILE - 1) / PS_TILE + 1) * sizeof(uint32_t); }

// out[i] = sum of in[0 .. i) (exclusive) or in[0 .. i] (inclusive); in == out allowed; temp: prefix_sum_temp_bytes(n)
cudaError_t prefix_sum_u32(const uint32_t *in, uint32_t *out, int64_t n, bool inclusive, void *temp, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const int64_t tiles = (n + PS_TILE - 1) / PS_TILE;
    uint32_t *sums = static_cast<uint32_t *>(temp);
    ps_tiles_kernel<<<(unsigned)tiles, PS_THREADS, 0, st>>>(in, out, n, sums, inclusive ? 1 : 0);
    if (tiles > 1) {
        ps_sums_kernel<<<1, PS_THREADS, 0, st>>>(sums, tiles);
        ps_add_kernel<<<(unsigned)tiles, PS_THREADS, 0, st>>>(out, n, sums);
    }
    return cudaGetLastError();
}

// ---- radix sort -------------------------------------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const K *keys, int64_t n, int shift, int64_t n_wt, uint32_t *hist)
{
    __shared__ uint32_t cnt[RS_WARPS][RS_DIGITS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * RS_WARPS + wib;
    for (int d = lane; d < RS_DIGITS; d += 32) cnt[wib][d] = 0;
    __syncwarp();
    if (w < n_wt) {
        const int64_t i0 = w * RS_WARP_ITEMS;
#pragma unroll 4
        for (int s = 0; s < RS_WARP_ITEMS; s += 32) {
            const int64_t i = i0 + s + lane;
            if (i < n) atomicAdd(&cnt[wib][(uint32_t)(keys[i] >> shift) & 255u], 1u);
        }
        __syncwarp();
        for (int d = lane; d < RS_DIGITS; d += 32) hist[(int64_t)d * n_wt + w] = cnt[wib][d];
    }
}

template <typename K, typename V>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const K *kin, const V *vin, K *kout, V *vout, int64_t n, int shift, int64_t n_wt, const uint32_t *offs)
{
    __shared__ uint32_t off[RS_WARPS][RS_DIGITS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * RS_WARPS + wib;
    if (w >= n_wt) return;
    for (int d = lane; d < RS_DIGITS; d += 32) off[wib][d] = offs[(int64_t)d * n_wt + w];
    __syncwarp();
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t i0 = w * RS_WARP_ITEMS;
    for (int s = 0; s < RS_WARP_ITEMS; s += 32) {
        const int64_t i = i0 + s + lane;
        if (i0 + s >= n) break;
        const bool valid = i < n;
        K key{};
        V val{};
        if (valid) { key = kin[i]; val = vin[i]; }
        const uint32_t d = valid ? ((uint32_t)(key >> shift) & 255u) : (256u + (uint32_t)lane);     // idle lanes match nobody
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(m & lt);
        uint32_t base = 0;
        if (valid) base = off[wib][d];
        __syncwarp();
        if (valid && rank == 0) off[wib][d] = base + __popc(m);
        __syncwarp();
        if (valid) { kout[base + rank] = key; vout[base + rank] = val; }
    }
}

size_t radix_sort_temp_bytes(int64_t n)
{
    const int64_t n_wt = (n + RS_WARP_ITEMS - 1) / RS_WARP_ITEMS;
    const int64_t nh = n_wt * RS_DIGITS;
    return (size_t)nh * sizeof(uint32_t) + prefix_sum_temp_bytes(nh) + 256;
}

// Stable sort of n pairs by bits [begin_bit, end_bit) of the key.  The pairs ping-pong between (ka, va) and (kb, vb);
// returns in *in_b whether the result is in the b buffers.  n < 2^32.
template <typename K, typename V>
cudaError_t radix_sort_pairs(K *ka, K *kb, V *va, V *vb, int64_t n, int begin_bit, int end_bit, void *temp, bool *in_b,
                             int64_t *n_launches, cudaStream_t st)
{
    *in_b = false;
    if (n <= 0 || end_bit <= begin_bit) return cudaSuccess;
    const int64_t n_wt = (n + RS_WARP_ITEMS - 1) / RS_WARP_ITEMS;
    const int64_t nh = n_wt * RS_DIGITS;
    uint32_t *hist = static_cast<uint32_t *>(temp);
    void *ps_temp = static_cast<uint8_t *>(temp) + (((size_t)nh * sizeof(uint32_t) + 255) & ~(size_t)255);
    const unsigned blocks = (unsigned)((n_wt + RS_WARPS - 1) / RS_WARPS);
    K *kin = ka, *kout = kb;
    V *vin = va, *vout = vb;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        rs_hist_kernel<K><<<blocks, RS_THREADS, 0, st>>>(kin, n, shift, n_wt, hist);
        cudaError_t e = prefix_sum_u32(hist, hist, nh, false, ps_temp, st);
        if (e != cudaSuccess) return e;
        rs_scatter_kernel<K, V><<<blocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, n_wt, hist);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        std::swap(kin, kout);
        std::swap(vin, vout);
        *in_b = !*in_b;
        if (n_launches) *n_launches += 2 + (nh > PS_TILE ? 3 : 1);
    }
    return cudaSuccess;
}

cudaError_t radix_sort_hits(uint64_t *ka, uint64_t *kb, SeedHit *va, SeedHit *vb, int64_t n, int end_bit, void *temp, bool *in_b,
                            int64_t *n_launches, cudaStream_t st)
{
    static_assert(sizeof(SeedHit) == sizeof(uint4), "a seed hit moves as one 16-byte word");
    return radix_sort_pairs<uint64_t, uint4>(ka, kb, reinterpret_cast<uint4 *>(va), reinterpret_cast<uint4 *>(vb), n, 0, end_bit,
                                             temp, in_b, n_launches, st);
}

cudaError_t radix_sort_u32(uint32_t *ka, uint32_t *kb, uint32_t *va, uint32_t *vb, int64_t n, int end_bit, void *temp, bool *in_b,
                           int64_t *n_launches, cudaStream_t st)
{
    return radix_sort_pairs<uint32_t, uint32_t>(ka, kb, va, vb, n, 0, end_bit, temp, in_b, n_launches, st);
}

}  // namespace bn
