// triage_kernel.cu — device-side triage of the speculative gapped extensions.
//
// BLAST_GetGappedScore (core/blast_gapalign.c:3351-3547) walks the init-HSPs of a subject chunk in score order; an
// init-HSP that is not contained in an HSP saved before it is extended (gapped_stats->extensions), and saved only when
// the extension reaches hit_params->cutoffs[context].cutoff_score.  A LOSER — an init-HSP whose extension stays below
// the cutoff — therefore never changes the tree or the list: all it can contribute is one count, and whether it counts
// depends only on the HSPs saved for its own query strand in its own chunk.  In blastn mode nearly every init-HSP is
// a loser (C3: 4.4 million of them around ~800 saved HSPs), so instead of shipping every init-HSP and every
// extension to the host replay, the device
//   1. classifies: winner (extension >= cutoff) or loser, and copies the winners out densely;
//   2. looks at every loser of a (chunk, context) that has winners: if a winner's alignment box could contain its
//      ungapped box (the necessary condition of s_HSPIsContained, core/blast_itree.c:815-852: both end points inside
//      the box, score not above) the loser is copied out too and the host replays it in order with the winners,
//      exactly; every other loser is certain to be extended and is only counted, per (chunk, context).
// The host replay then runs over winners + undecided losers and adds the counted ones to the statistic.
#include "bn_device.cuh"

namespace bn {

constexpr uint32_t HAS_WINNER = 0x80000000u;

__device__ __forceinline__ uint64_t warp_append(unsigned long long *counter)
{
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

__global__ void __launch_bounds__(256)
triage_classify_kernel(const DevQuery q, const TriageLaunch t)
{
    const int64_t n = min((int64_t)*t.n_init, t.max_init);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const DevInitHit h = t.init[i];
        const DevGapResult g = t.gap[i];
        const int32_t ctx = ctx_search(q, h.q_off);
        const bool winner = g.score >= __ldg(&q.ctx[ctx].gapped_cutoff);
        t.ctx_of[i] = winner ? (ctx | (int32_t)HAS_WINNER) : ctx;
        if (winner) {
            const uint64_t slot = warp_append(&t.tcount[0]);
            uint2 *cell = &t.table[(size_t)h.chunk * (size_t)t.n_ctx + (size_t)ctx];
            if ((int64_t)slot < t.sel_cap) {
                // the winners of a (chunk, context) form a chain through sel_ctx: cell.y = newest winner + 1
                t.sel_init[slot] = h; t.sel_gap[slot] = g;
                t.sel_ctx[slot] = (int32_t)atomicExch(&cell->y, (uint32_t)slot + 1u);
            }
            atomicOr(&cell->x, HAS_WINNER);
        }
    }
}

__global__ void __launch_bounds__(256)
triage_losers_kernel(const DevQuery q, const TriageLaunch t)
{
    const int64_t n = min((int64_t)*t.n_init, t.max_init);
    const int64_t n_w = (int64_t)t.tcount[0];      // > sel_cap: the caller grows the buffers and runs the triage again
    unsigned long long counted = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = t.ctx_of[i];
        if (c < 0) continue;                                   // winner
        const DevInitHit h = t.init[i];
        const uint2 *cell = &t.table[(size_t)h.chunk * (size_t)t.n_ctx + (size_t)c];
        bool undecided = false;
        if (__ldg(&cell->x) & HAS_WINNER) {
            const int32_t q0 = h.q_start - __ldg(&q.ctx[c].query_offset), q1 = q0 + h.length;
            const int32_t s0 = h.s_start, s1 = s0 + h.length;
            for (uint32_t k1 = __ldg(&cell->y); k1 && !undecided; k1 = (uint32_t)t.sel_ctx[k1 - 1]) {
                const DevGapResult w = t.sel_gap[k1 - 1];
                undecided = h.score <= w.score && w.q_start <= q0 && q0 <= w.q_stop && w.s_start <= s0 && s0 <= w.s_stop &&
                            w.q_start <= q1 && q1 <= w.q_stop && w.s_start <= s1 && s1 <= w.s_stop;
            }
        }
        if (undecided) {
            const uint64_t slot = (uint64_t)n_w + warp_append(&t.tcount[1]);
            if ((int64_t)slot < t.sel_cap) { t.sel_init[slot] = h; t.sel_gap[slot] = t.gap[i]; t.sel_ctx[slot] = 0; }
        } else ++counted;
    }
    // one atomic per warp: the counted losers only matter as a total (the triage runs when no low_score bound can
    // move, so every one of them is an extension the reference makes)
    for (int o = 16; o > 0; o >>= 1) counted += __shfl_down_sync(0xffffffffu, counted, o);
    if ((threadIdx.x & 31) == 0 && counted) atomicAdd(&t.tcount[2], counted);
}

cudaError_t launch_triage(const DevQuery &q, const TriageLaunch &t, cudaStream_t st)
{
    const int blocks = 148 * 8;
    triage_classify_kernel<<<blocks, 256, 0, st>>>(q, t);
    triage_losers_kernel<<<blocks, 256, 0, st>>>(q, t);
    return cudaGetLastError();
}

// indices of the extensions with a given status (tier hand-over lists built on the device)
__global__ void collect_status_kernel(const DevGapResult *gap, const unsigned long long *n_init, int64_t max_init,
                                      int32_t want, int32_t *todo, unsigned long long *count)
{
    const int64_t n = min((int64_t)*n_init, max_init);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (gap[i].status == want) todo[warp_append(count)] = (int32_t)i;
}
cudaError_t launch_collect_status(const DevGapResult *gap, const unsigned long long *n_init, int64_t max_init, int32_t want,
                                  int32_t *todo, unsigned long long *count, cudaStream_t st)
{
    collect_status_kernel<<<148 * 4, 256, 0, st>>>(gap, n_init, max_init, want, todo, count);
    return cudaGetLastError();
}

}  // namespace bn
