// group_sort.cu — ordering of the seed hits for the diagonal stage without a host round trip.
//
// The diagonal-hash stage (extend_kernel.cu) needs the survivors of the scan grouped by hash bucket
// (512 groups, DIAGHASH_NUM_BUCKETS core/blast_extend.h) and, inside a bucket, in the reference's
// emission order (subject offset ascending, lookup chain order: core/blast_nascan.c:1413-1427 emits a
// cell's chain in chain order and the scanners walk the subject left to right).
//
// The general path is one cub radix sort on (bucket, global scan position), which needs the number
// of survivors on the host (a stream synchronisation) and 7 launches.  For the usual case — at most
// a few hundred thousand survivors, no bucket larger than BUCKET_MAX — this file does the same
// ordering as a counting sort whose sizes never leave the device:
//   scan kernel        counts survivors per bucket while it emits them            (scan_kernel.cu)
//   bucket_plan        exclusive scan of the 512 counts; decides whether the fast path applies
//   bucket_scatter     key (position << 24 | emission slot) of every survivor into its bucket's range
//   bucket_sort        one block per bucket: bitonic sort of the keys in shared memory, gather of the
//                      hits, group head, speculative-extension leaders (what group_heads_kernel does
//                      on the general path)
// "emission slot" = index in the scan's output buffer.  One thread emits all hits of a scan position
// and its atomic slot reservations are ordered in time, so (position, slot) ascending IS the
// reference's order, and the key is unique, which makes the (unstable) bitonic network deterministic.
// When the fast path does not apply the kernels do nothing but raise counters[6]; the host sees it
// at its next synchronisation and re-runs the general path on the untouched scan output.
#include "bn_device.cuh"

namespace bn {

constexpr int NBUCKETS = 512;
constexpr int BUCKET_MAX = 2048;          // largest bucket the shared-memory sort takes
constexpr int SORT_THREADS = 256;

int group_sort_buckets() { return NBUCKETS; }

__global__ void __launch_bounds__(NBUCKETS)
bucket_plan_kernel(const BucketLaunch L)
{
    __shared__ uint32_t warp_sum[NBUCKETS / 32];
    __shared__ uint32_t warp_max[NBUCKETS / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t c = L.bucket_count[tid];
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    const uint32_t m = __reduce_max_sync(0xffffffffu, c);
    if (lane == 31) warp_sum[w] = x;
    if (lane == 0) warp_max[w] = m;
    __syncthreads();
    uint32_t base = 0, mx = 0;
    for (int i = 0; i < NBUCKETS / 32; i++) {
        if (i < w) base += warp_sum[i];
        mx = max(mx, warp_max[i]);
    }
    L.bucket_start[tid] = base + x - c;
    if (tid == NBUCKETS - 1) L.bucket_start[NBUCKETS] = base + x;
    L.cursor[tid] = 0;
    if (tid == 0) {
        const unsigned long long n = L.counters[0];
        if (n > (unsigned long long)L.n_limit || mx > (uint32_t)BUCKET_MAX) L.counters[6] = 1ull;
    }
}

__global__ void bucket_scatter_kernel(const BucketLaunch L)
{
    if (L.counters[6]) return;
    const int64_t n = (int64_t)L.counters[0];
    const uint64_t gmask = (1ull << L.gbits) - 1ull;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t key = L.keys_in[j];
        const uint32_t b = (uint32_t)(key >> L.gbits);
        const uint32_t pos = L.bucket_start[b] + atomicAdd(&L.cursor[b], 1u);
        L.keys_tmp[pos] = ((key & gmask) << 24) | (uint64_t)j;
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
bucket_sort_kernel(const BucketLaunch L)
{
    __shared__ uint64_t sk[BUCKET_MAX];
    if (L.counters[6]) return;
    const uint32_t b = blockIdx.x;
    const uint32_t start = L.bucket_start[b];
    const int cnt = (int)(L.bucket_start[b + 1] - start);
    if (cnt == 0) return;
    const int tid = threadIdx.x;
    int P = 32;
    while (P < cnt) P <<= 1;
    for (int i = tid; i < P; i += SORT_THREADS) sk[i] = i < cnt ? L.keys_tmp[start + i] : ~0ull;
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
    bucket_plan_kernel<<<1, NBUCKETS, 0, st>>>(L);
    bucket_scatter_kernel<<<148 * 2, 256, 0, st>>>(L);
    bucket_sort_kernel<<<NBUCKETS, SORT_THREADS, 0, st>>>(L);
    return cudaGetLastError();
}

}  // namespace bn
