// group_sort.cu — ordering of the seed hits for the diagonal stage without a host round trip.
//
// The diagonal-hash stage (extend_kernel.cu) needs the survivors of the scan grouped by hash bucket
// (512 groups, DIAGHASH_NUM_BUCKETS core/blast_extend.h) and, inside a bucket, in the reference's
// emission order (subject offset ascending, lookup chain order: core/blast_nascan.c:1413-1427 emits a
// cell's chain in chain order and the scanners walk the subject left to right).
//
// The general path is one radix sort (radix_sort.cu) on (bucket, global scan position), which needs the number
// of survivors on the host (a stream synchronisation) and 7 launches.  For the usual case - at most
// a few hundred thousand survivors, no bucket larger than BUCKET_MAX - the same ordering comes out of
// a counting sort whose sizes never leave the device, in ONE launch after the scan:
//   scan kernel        its per-bucket atomic counter (needed anyway) also gives every survivor a slot in
//                      its bucket's fixed-capacity region, where it drops the key
//                      (position << 24 | emission slot)                                   (scan_kernel.cu)
//   bucket_sort        one block per bucket: exclusive prefix of the 512 counts (every block sums them
//                      itself: 2 KB of L2 reads), bitonic sort of the bucket's keys in shared memory, gather
//                      of the hits, group head, speculative-extension leaders (what group_heads_kernel does
//                      on the general path)
// "emission slot" = index in the scan's output buffer.  One thread emits all hits of a scan position
// and its atomic slot reservations are ordered in time, so (position, slot) ascending IS the
// reference's order, and the key is unique, which makes the (unstable) bitonic network deterministic.
// When the fast path does not apply the kernel does nothing but raise counters[6]; the host sees it
// at its next synchronisation and re-runs the general path on the untouched scan output.
#include "bn_device.cuh"

namespace bn {

constexpr int NBUCKETS = 512;
constexpr int BUCKET_MAX = 2048;          // largest bucket the shared-memory sort takes
constexpr int SORT_THREADS = 256;

int group_sort_buckets() { return NBUCKETS; }

int group_sort_bucket_cap() { return BUCKET_MAX; }
static_assert(NBUCKETS == 2 * SORT_THREADS, "every thread sums two bucket counts");

__global__ void __launch_bounds__(SORT_THREADS)
bucket_sort_kernel(const BucketLaunch L)
{
    __shared__ uint64_t sk[BUCKET_MAX];
    __shared__ uint32_t red_sum[SORT_THREADS / 32], red_max[SORT_THREADS / 32];
    const uint32_t b = blockIdx.x;
    uint32_t start = 0;
    {   // position of this bucket in the compact output = sum of the counts before it; refusal test
        const uint32_t t = threadIdx.x;
        const uint32_t c0 = L.bucket_count[t], c1 = L.bucket_count[t + SORT_THREADS];
        const uint32_t before = (t < b ? c0 : 0u) + (t + SORT_THREADS < b ? c1 : 0u);
        const uint32_t ws = __reduce_add_sync(0xffffffffu, before);
        const uint32_t wm = __reduce_max_sync(0xffffffffu, max(c0, c1));
        if ((t & 31) == 0) { red_sum[t >> 5] = ws; red_max[t >> 5] = wm; }
        __syncthreads();
        uint32_t mx = 0;
#pragma unroll
        for (int i = 0; i < SORT_THREADS / 32; i++) { start += red_sum[i]; mx = max(mx, red_max[i]); }
        if (L.counters[0] > (unsigned long long)L.n_limit || mx > (uint32_t)BUCKET_MAX) {
            if (b == 0 && t == 0) L.counters[6] = 1ull;
            return;
        }
    }
    const int cnt = (int)L.bucket_count[b];
    if (cnt == 0) return;
    const uint64_t *mine = L.keys_tmp + (size_t)b * BUCKET_MAX;
    const int tid = threadIdx.x;
    int P = 32;
    while (P < cnt) P <<= 1;
    for (int i = tid; i < P; i += SORT_THREADS) sk[i] = i < cnt ? mine[i] : ~0ull;
    __syncthreads();
    if (cnt > 1) {
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P; i += SORT_THREADS) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint64_t a = sk[i], c = sk[ixj];
                        const bool up = (i & k) == 0;
                        if ((a > c) == up) { sk[i] = c; sk[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
    }
    if (tid == 0) L.heads[atomicAdd(&L.counters[4], 1ull)] = start;
    for (int r0 = 0; r0 < cnt; r0 += SORT_THREADS) {
        const int r = r0 + tid;
        bool leader = false;
        if (r < cnt) {
            const uint64_t key = sk[r];
            const SeedHit h = L.hits_in[key & 0xFFFFFFull];
            L.hits_out[start + r] = h;
            L.keys_out[start + r] = ((uint64_t)b << L.gbits) | (key >> 24);
            L.spec[start + r].status = 0;                    // SPEC_NONE
            leader = L.spec_enabled != 0;
            if (leader && r > 0) {
                const SeedHit p = L.hits_in[sk[r - 1] & 0xFFFFFFull];
                leader = p.chunk != h.chunk || (p.s_off - p.q_off) != (h.s_off - h.q_off);
            }
        }
        // warp-aggregated append of the leaders
        const unsigned m = __ballot_sync(0xffffffffu, leader);
        if (m) {
            const int lane = tid & 31, ldr = __ffs(m) - 1;
            unsigned long long base = 0;
            if (lane == ldr) base = atomicAdd(&L.counters[5], (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, ldr);
            if (leader) L.leaders[base + __popc(m & ((1u << lane) - 1u))] = start + (uint32_t)r;
        }
    }
}

cudaError_t launch_bucket_group(const BucketLaunch &L, cudaStream_t st)
{
    bucket_sort_kernel<<<NBUCKETS, SORT_THREADS, 0, st>>>(L);
    return cudaGetLastError();
}

}  // namespace bn
