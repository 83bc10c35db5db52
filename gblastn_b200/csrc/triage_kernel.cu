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
        const bool winner = g.status == 0 && g.score >= __ldg(&q.ctx[ctx].gapped_cutoff);
        t.ctx_of[i] = winner ? (ctx | (int32_t)HAS_WINNER) : ctx;
        if (winner) {
            const uint64_t slot = warp_append(&t.tcount[0]);
            uint2 *cell = &t.table[(size_t)h.chunk * (size_t)t.n_ctx + (size_t)ctx];
            if ((int64_t)slot < t.sel_cap) {
                // the winners of a (chunk, context) form a chain through sel_ctx: cell.y = newest winner + 1
                t.sel_init[slot] = h; t.sel_gap[slot] = g; t.sel_idx[slot] = (int32_t)i;
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
        // an extension that was not computed (status 3: a long alignment expected to be contained in one already made,
        // long_rounds below) is for the host replay to judge
        bool undecided = t.gap[i].status == 3;
        if (!undecided && (__ldg(&cell->x) & HAS_WINNER)) {
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
            if ((int64_t)slot < t.sel_cap) { t.sel_init[slot] = h; t.sel_gap[slot] = t.gap[i]; t.sel_ctx[slot] = 0; t.sel_idx[slot] = (int32_t)i; }
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

// ---- long alignments in rounds ----------------------------------------------------------------------------
// The speculative gapped stage extends EVERY init-HSP.  A real 10 kb alignment contains hundreds of init-HSPs, and
// each of them would repeat the same 10 kb extension, where the reference makes it once and finds the others
// contained in it (core/blast_gapalign.c:3400-3404).  So the long extensions (the ones tier 1 handed over) are made in
// rounds: per (chunk, context) the best-scoring pending one is extended, then every pending one that lies inside a
// box made so far — the reference's own containment condition, s_HSPIsContained core/blast_itree.c:815-852 — is
// set aside (status 3, never computed) and the next best of what remains goes into the next round.  This is only a
// prediction of the host replay's decisions: the replay decides, and asks for any set-aside extension it turns out
// to need (engine.cu: resolve_set_aside).
__global__ void long_prepare_kernel(const DevQuery q, const LongRounds r)
{
    const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= r.n_todo) return;
    r.ctx_w[w] = ctx_search(q, r.init[r.todo[w]].q_off);
    r.state[w] = 0;
    r.chain_next[w] = 0;
}
__device__ __forceinline__ unsigned long long long_key(int32_t score, int32_t w)
{
    return ((unsigned long long)(uint32_t)(INT32_MAX - score) << 32) | (uint32_t)w;
}
__global__ void long_select_kernel(const DevQuery q, const LongRounds r)
{
    const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= r.n_todo || r.state[w] != 0) return;
    const int32_t i = r.todo[w], c = r.ctx_w[w];
    const DevInitHit h = r.init[i];
    const size_t cell = (size_t)h.chunk * (size_t)r.n_ctx + (size_t)c;
    const int32_t q0 = h.q_start - __ldg(&q.ctx[c].query_offset), q1 = q0 + h.length, s0 = h.s_start, s1 = s0 + h.length;
    if (r.set_aside_all && r.chain_head[cell]) {       // test switch: a deliberately wrong prediction
        r.state[w] = 3;
        r.gap[i].status = 3;
        return;
    }
    for (uint32_t k1 = r.chain_head[cell]; k1; k1 = (uint32_t)r.chain_next[k1 - 1]) {
        const DevGapResult W = r.gap[r.todo[k1 - 1]];
        if (h.score <= W.score && W.q_start <= q0 && q0 <= W.q_stop && W.s_start <= s0 && s0 <= W.s_stop &&
            W.q_start <= q1 && q1 <= W.q_stop && W.s_start <= s1 && s1 <= W.s_stop) {
            const int32_t dw0 = W.q_start - W.s_start, dw1 = W.q_stop - W.s_stop, d0 = q0 - s0, d1 = q1 - s1;
            if (r.min_diag_separation == 0 || abs(dw0 - d0) < r.min_diag_separation || abs(dw1 - d1) < r.min_diag_separation) {
                r.state[w] = 3;
                r.gap[i].status = 3;
                return;
            }
        }
    }
    atomicMin(&r.best[cell], long_key(h.score, w));
}
__global__ void long_pick_kernel(const LongRounds r)
{
    const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= r.n_todo || r.state[w] != 0) return;
    const int32_t i = r.todo[w];
    const DevInitHit h = r.init[i];
    const size_t cell = (size_t)h.chunk * (size_t)r.n_ctx + (size_t)r.ctx_w[w];
    if (r.best[cell] != long_key(h.score, w)) return;
    const unsigned long long slot = warp_append(r.round_count);
    r.round_list[slot] = i;
    r.round_w[slot] = w;
    r.state[w] = 2;
}
__global__ void long_commit_kernel(const DevQuery q, const LongRounds r, int32_t n_round)
{
    const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_round) return;
    const int32_t w = r.round_w[k], i = r.round_list[k], c = r.ctx_w[w];
    r.state[w] = 1;
    const DevGapResult g = r.gap[i];
    if (g.status == 0 && g.score >= __ldg(&q.ctx[c].gapped_cutoff)) {
        const size_t cell = (size_t)r.init[i].chunk * (size_t)r.n_ctx + (size_t)c;
        r.chain_next[w] = (int32_t)atomicExch(&r.chain_head[cell], (uint32_t)w + 1u);
    }
}
cudaError_t launch_long_prepare(const DevQuery &q, const LongRounds &r, cudaStream_t st)
{
    if (r.n_todo > 0) long_prepare_kernel<<<(r.n_todo + 255) / 256, 256, 0, st>>>(q, r);
    return cudaGetLastError();
}
cudaError_t launch_long_select(const DevQuery &q, const LongRounds &r, cudaStream_t st)
{
    if (r.n_todo > 0) {
        long_select_kernel<<<(r.n_todo + 255) / 256, 256, 0, st>>>(q, r);
        long_pick_kernel<<<(r.n_todo + 255) / 256, 256, 0, st>>>(r);
    }
    return cudaGetLastError();
}
cudaError_t launch_long_commit(const DevQuery &q, const LongRounds &r, int32_t n_round, cudaStream_t st)
{
    if (n_round > 0) long_commit_kernel<<<(n_round + 255) / 256, 256, 0, st>>>(q, r, n_round);
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
