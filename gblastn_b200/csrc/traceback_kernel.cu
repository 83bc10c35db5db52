// traceback_kernel.cu — gapped X-drop alignment WITH traceback on the device.
//
// Reference: BLAST_GappedAlignmentWithTraceback (core/blast_gapalign.c:3994-4155) -> Blast_SemiGappedAlign
// with score_only == FALSE (:711-745) -> ALIGN_EX (:350-709), as Blast_TracebackFromHSPList calls it for
// blastn with eDynProgTbck (core/blast_traceback.c:565-571) on the blastna subject (one base per byte).
//
// Here the subject stays in its packed ncbi2na form in HBM (the volume has no ambiguity data, so the
// blastna byte of base p is its 2-bit code) and one WARP computes one direction of one alignment with the
// row-parallel formulation of gapped_kernel.cu (lane per band cell, prefix maxima for the horizontal gap
// and the running best, prune flags by fixed-point iteration — exact).  On top of the score-only kernel
// every visited cell stores the reference's script byte (operation | "gap in A continues" 0x10 |
// "gap in B continues" 0x40) in a global arena that is laid out like the reference's GapStateArrayStruct:
// rows are appended back to back, row a starts at column row_first[a].  Lane 0 then walks the script from
// the best cell back to the origin (same state machine as core/blast_gapalign.c:660-702) and the warp
// writes the run-length edit operations.  The host joins the two directions exactly like
// Blast_PrelimEditBlockToGapEditScript (core/blast_gapalign.c:2455-2517).
#include "bn_device.cuh"

namespace bn {

namespace {

constexpr unsigned FULLW = 0xffffffffu;
constexpr int TB_WARPS = 4;               // warps per block
constexpr int TB_CELLS = 1024;            // band cells per warp kept in shared memory (power of two)
constexpr int32_t MININT = INT32_MIN / 2;
constexpr int32_t NEGINF = INT32_MIN / 2 - (1 << 24);

constexpr uint8_t SCRIPT_SUB = 3, SCRIPT_GAP_IN_A = 0, SCRIPT_GAP_IN_B = 6;      // eGapAlignSub / Del / Ins
constexpr uint8_t SCRIPT_OP_MASK = 0x07, SCRIPT_EXTEND_GAP_A = 0x10, SCRIPT_EXTEND_GAP_B = 0x40;

__device__ __forceinline__ int32_t excl_prefix_max(int32_t x, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(FULLW, x, o);
        if (lane >= o) x = max(x, y);
    }
    const int32_t e = __shfl_up_sync(FULLW, x, 1);
    return lane == 0 ? NEGINF : e;
}

__device__ __forceinline__ int sbase64(const uint8_t *s, int64_t pos)      // NCBI2NA_UNPACK_BASE
{
    return (__ldg(s + (pos >> 2)) >> (6 - 2 * (int)(pos & 3))) & 3;
}

// bump allocation from the launch's arena; returns -1 when it is exhausted
__device__ __forceinline__ long long arena_alloc(const TracebackLaunch &L, long long bytes)
{
    bytes = (bytes + 15) & ~15ll;
    const long long at = (long long)atomicAdd(L.arena_used, (unsigned long long)bytes);
    return (at + bytes <= L.arena_bytes) ? at : -1;
}

struct RowStore {               // where the script rows of the extension in progress go
    long long cur, end;         // current chunk of the arena
    long long *row_off;         // script byte of (a, b) = arena[row_off[a] + b - row_first[a]]
    int32_t *row_first;
};
constexpr long long ROW_CHUNK = 128 << 10;

// ALIGN_EX for one direction.  M rows (query), N columns (subject).  qrow(a) = query byte of row a,
// sub(b) = subject base that the diagonal step INTO column b consumes (b >= 1).
// status: 0 ok, 1 band wider than the shared-memory ring, 3 arena exhausted.
__device__ int32_t align_ex_warp(const TracebackLaunch &L, const uint8_t *qp, int q_inc, const uint8_t *S, int64_t s0, int s_inc,
                                 int32_t M, int32_t N, const int32_t *matrix, int32_t gap_open, int32_t gap_extend,
                                 int32_t x_dropoff, int2 *ring, RowStore &rs, int32_t &a_offset, int32_t &b_offset,
                                 int &status, int lane)
{
    const int32_t goe = gap_open + gap_extend, ge = gap_extend;
    constexpr int32_t C = TB_CELLS, MASK = TB_CELLS - 1;
    a_offset = 0; b_offset = 0;
    if (x_dropoff < goe) x_dropoff = goe;
    if (N <= 0 || M <= 0) return 0;
    const uint32_t lt = (1u << lane) - 1u;
    const int32_t num_extra = x_dropoff / ge + 3;
    uint8_t *arena = L.arena;

    // row 0: cell 0 = (0, -goe); cells i >= 1 = (-goe - (i-1) ge, that - goe) while the score is >= -X; script GAP_IN_A
    int32_t b_size;
    {
        int32_t k = min((x_dropoff - goe) / ge + 1, N);
        if (k + 2 >= C) { status = 1; return 0; }
        if (rs.cur + k + 2 > rs.end) {
            long long at = 0;
            if (lane == 0) at = arena_alloc(L, ROW_CHUNK);
            at = __shfl_sync(FULLW, at, 0);
            if (at < 0) { status = 3; return 0; }
            rs.cur = at; rs.end = at + ROW_CHUNK;
        }
        for (int32_t i = lane; i <= k; i += 32) {
            const int32_t sc = (i == 0) ? 0 : -goe - (i - 1) * ge;
            ring[i & MASK] = make_int2(sc, sc - goe);
            arena[rs.cur + i] = SCRIPT_GAP_IN_A;
        }
        if (lane == 0) { rs.row_off[0] = rs.cur; rs.row_first[0] = 0; }
        rs.cur += k + 2;
        b_size = k + 1;
        __syncwarp();
    }
    int32_t best_score = 0, first_b = 0;

    for (int32_t a_index = 1; a_index <= M; a_index++) {
        const int32_t *mrow = matrix + 16 * (int)__ldg(qp + (int64_t)a_index * q_inc);
        const int32_t row_first = first_b;
        // space for this row: it can reach b_size + num_extra columns at most (core/blast_gapalign.c:478-484)
        {
            const long long need = (long long)(b_size - row_first) + num_extra + 3;
            if (rs.cur + need > rs.end) {
                long long at = 0;
                const long long sz = need > ROW_CHUNK ? need : ROW_CHUNK;
                if (lane == 0) at = arena_alloc(L, sz);
                at = __shfl_sync(FULLW, at, 0);
                if (at < 0) { status = 3; return 0; }
                rs.cur = at; rs.end = at + sz;
            }
            if (lane == 0) { rs.row_off[a_index] = rs.cur; rs.row_first[a_index] = row_first; }
        }
        uint8_t *srow = arena + rs.cur - row_first;          // srow[b] = script of column b
        int32_t r_in = MININT, best_in = best_score, prev_old_best = MININT;
        int32_t last_b = first_b, new_first = first_b;
        bool lead = true;
        const int32_t row_end = b_size;

        for (int32_t seg = row_first; seg < row_end; seg += 32) {
            const int32_t b = seg + lane;
            const bool active = b < row_end;
            const uint32_t amask = __ballot_sync(FULLW, active);
            const int2 cell = active ? ring[b & MASK] : make_int2(MININT, MININT);
            int32_t up = __shfl_up_sync(FULLW, cell.x, 1);
            if (lane == 0) up = prev_old_best;
            prev_old_best = __shfl_sync(FULLW, cell.x, 31);
            int32_t v = NEGINF, d = MININT;
            if (active) {
                if (b != row_first) d = up + mrow[sbase64(S, s0 + (int64_t)b * s_inc)];
                v = max(d, cell.y);
            }
            // ---- fixed point over the prune flags (see gapped_kernel.cu) ---------------------------------
            uint32_t p = __ballot_sync(FULLW, active && (best_in - v > x_dropoff));
            int32_t s = v, R = r_in;
            for (;;) {
                const uint32_t um = amask & ~p;
                const bool unpruned = (um >> lane) & 1u;
                const int32_t u = __popc(um & lt);
                const int32_t w = unpruned ? v - goe + ge * (u + 1) : NEGINF;
                R = max(r_in, excl_prefix_max(w, lane)) - ge * u;
                s = max(v, R);
                const int32_t best_b = max(best_in, excl_prefix_max(unpruned ? s : NEGINF, lane));
                const uint32_t pn = __ballot_sync(FULLW, active && (best_b - s > x_dropoff));
                if (pn == p) break;
                p = pn;
            }
            const uint32_t um = amask & ~p;
            const bool unpruned = (um >> lane) & 1u;
            // ---- script byte of the cell (core/blast_gapalign.c:560-612) ---------------------------------
            if (active) {
                // script = SUB; if (score < gap_col) GAP_IN_B; if (score < gap_row) GAP_IN_A
                uint8_t op = (v < R) ? SCRIPT_GAP_IN_A : ((d < cell.y) ? SCRIPT_GAP_IN_B : SCRIPT_SUB);
                if (unpruned) {
                    if (cell.y - ge >= s - goe) op += SCRIPT_EXTEND_GAP_B;
                    if (R - ge >= s - goe) op += SCRIPT_EXTEND_GAP_A;
                }
                srow[b] = op;
            }
            // ---- commit the segment ----------------------------------------------------------------------
            const int nact = __popc(amask);
            int dropped = 0;
            if (lead) {
                dropped = um ? (__ffs(um) - 1) : nact;
                new_first += dropped;
                lead = (dropped == nact);
            }
            if (active) {
                if (unpruned) ring[b & MASK] = make_int2(s, max(s - goe, cell.y - ge));
                else if (lane >= dropped) ring[b & MASK] = make_int2(MININT, cell.y);
            }
            if (um) {
                last_b = seg + (31 - __clz(um));
                const int32_t m = __reduce_max_sync(FULLW, unpruned ? s : NEGINF);
                if (m > best_in) {
                    const uint32_t at = __ballot_sync(FULLW, unpruned && s == m);
                    best_in = m; a_offset = a_index; b_offset = seg + (__ffs(at) - 1);
                }
            }
            const int32_t r_next = unpruned ? max(s - goe, R - ge) : R;
            r_in = __shfl_sync(FULLW, r_next, nact - 1);
        }
        __syncwarp();
        best_score = best_in;
        first_b = new_first;
        if (first_b == b_size) break;
        if (last_b < b_size - 1) b_size = last_b + 1;
        else {
            int32_t k = 0;
            if (r_in >= best_score - x_dropoff) k = (r_in - (best_score - x_dropoff)) / ge + 1;
            k = max(0, min(k, N - b_size + 1));
            if (b_size + k - first_b + 2 >= C) { status = 1; return 0; }
            for (int32_t i = lane; i < k; i += 32) {
                const int32_t sc = r_in - i * ge;
                ring[(b_size + i) & MASK] = make_int2(sc, sc - goe);
                srow[b_size + i] = SCRIPT_GAP_IN_A;
            }
            b_size += k;
        }
        rs.cur += (long long)(max(row_end, b_size) - row_first) + 1;
        if (b_size <= N) {
            if (b_size - first_b + 2 >= C) { status = 1; return 0; }
            if (lane == 0) ring[b_size & MASK] = make_int2(MININT, MININT);
            b_size++;
        }
        __syncwarp();
    }
    return best_score;
}

}  // namespace

// One warp per (item, direction): warp 2 i = left extension of item i, warp 2 i + 1 = right extension.
__global__ void __launch_bounds__(TB_WARPS * 32)
traceback_dp_kernel(const DevQuery q, const TracebackLaunch L)
{
    __shared__ int2 rings[TB_WARPS][TB_CELLS];
    __shared__ int32_t s_matrix[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_matrix[i] = q.matrix[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * TB_WARPS + wib, nwarps = (int64_t)gridDim.x * TB_WARPS;
    int2 *ring = rings[wib];
    for (int64_t w = warp0; w < 2 * L.n; w += nwarps) {
        const DevTracebackItem it = L.items[w >> 1];
        const bool right = (w & 1) != 0;
        const DevContext c = q.ctx[it.context];
        const uint8_t *query = q.query + c.query_offset;
        const uint8_t *S = L.packed + it.byte_off;
        const int32_t q_length = c.query_length, s_length = it.s_length;
        DevTracebackDir out;
        out.score = 0; out.a_off = 0; out.b_off = 0; out.ops_off = 0; out.n_ops = 0; out.status = 0; out.ran = 0; out.pad = 0;
        int32_t M, N;
        const uint8_t *qp; int q_inc, s_inc; int64_t s0;
        const int64_t sabs = (int64_t)it.s_shift + it.s_start;
        if (!right) {   // Blast_SemiGappedAlign(query, subject, q_start + 1, s_start + 1, ..., reverse_sequence = TRUE)
            M = it.q_start + 1; N = it.s_start + 1;
            qp = query + it.q_start + 1; q_inc = -1;          // row a: query[q_start + 1 - a]
            s0 = sabs + 1; s_inc = -1;                        // column b consumes subject[s_start + 1 - b]
            out.ran = 1;
        } else {        // Blast_SemiGappedAlign(query + q_start, subject + s_start, q_length - q_start - 1, s_length - s_start - 1, ...)
            M = q_length - it.q_start - 1; N = s_length - it.s_start - 1;
            qp = query + it.q_start; q_inc = 1;               // row a: query[q_start + a]
            s0 = sabs; s_inc = 1;                             // column b consumes subject[s_start + b]
            out.ran = (it.q_start < q_length && it.s_start < s_length) ? 1 : 0;
        }
        int status = 0;
        if (out.ran && M > 0 && N > 0) {
            RowStore rs;
            rs.cur = rs.end = 0;
            long long tab = 0;
            if (lane == 0) tab = arena_alloc(L, (long long)(M + 1) * 12);
            tab = __shfl_sync(FULLW, tab, 0);
            if (tab < 0) status = 3;
            else {
                rs.row_off = reinterpret_cast<long long *>(L.arena + tab);
                rs.row_first = reinterpret_cast<int32_t *>(L.arena + tab + (long long)(M + 1) * 8);
                int32_t a_off, b_off;
                out.score = align_ex_warp(L, qp, q_inc, S, s0, s_inc, M, N, s_matrix, q.gap_open, q.gap_extend,
                                          L.x_dropoff, ring, rs, a_off, b_off, status, lane);
                out.a_off = a_off; out.b_off = b_off;
                __syncwarp();
                __threadfence_block();
                if (status == 0 && (a_off > 0 || b_off > 0)) {
                    // ---- walk the script back to the origin (core/blast_gapalign.c:660-702); runs go to a
                    // temporary list in the arena first (their number is not known in advance)
                    long long tmp = 0;
                    if (lane == 0) tmp = arena_alloc(L, (long long)(a_off + b_off) * 8);
                    tmp = __shfl_sync(FULLW, tmp, 0);
                    if (tmp < 0) status = 3;
                    else {
                        int2 *runs = reinterpret_cast<int2 *>(L.arena + tmp);
                        int32_t n_runs = 0;
                        if (lane == 0) {
                            int32_t a = a_off, b = b_off, run_op = -1, run_n = 0;
                            uint8_t script = SCRIPT_SUB;
                            const volatile uint8_t *ar = L.arena;
                            while (a > 0 || b > 0) {
                                const uint8_t next = ar[rs.row_off[a] + (b - rs.row_first[a])];
                                switch (script) {
                                case SCRIPT_GAP_IN_A:
                                    script = next & SCRIPT_OP_MASK;
                                    if (next & SCRIPT_EXTEND_GAP_A) script = SCRIPT_GAP_IN_A;
                                    break;
                                case SCRIPT_GAP_IN_B:
                                    script = next & SCRIPT_OP_MASK;
                                    if (next & SCRIPT_EXTEND_GAP_B) script = SCRIPT_GAP_IN_B;
                                    break;
                                default:
                                    script = next & SCRIPT_OP_MASK;
                                    break;
                                }
                                if (script == SCRIPT_GAP_IN_A) b--;
                                else if (script == SCRIPT_GAP_IN_B) a--;
                                else { a--; b--; }
                                if ((int32_t)script == run_op) run_n++;
                                else {
                                    if (run_n) runs[n_runs++] = make_int2(run_op, run_n);
                                    run_op = script; run_n = 1;
                                }
                            }
                            if (run_n) runs[n_runs++] = make_int2(run_op, run_n);
                        }
                        n_runs = __shfl_sync(FULLW, n_runs, 0);
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(L.ops_used, (unsigned long long)n_runs);
                        base = __shfl_sync(FULLW, base, 0);
                        if ((long long)(base + n_runs) > L.ops_cap) status = 4;
                        else {
                            __syncwarp();
                            for (int32_t i = lane; i < n_runs; i += 32) L.ops[base + i] = runs[i];
                            out.ops_off = (long long)base; out.n_ops = n_runs;
                        }
                    }
                }
            }
        }
        out.status = status;
        if (lane == 0) L.out[w] = out;
        __syncwarp();
    }
}

cudaError_t launch_traceback_dp(const DevQuery &q, const TracebackLaunch &L, int blocks, cudaStream_t st)
{
    traceback_dp_kernel<<<blocks, TB_WARPS * 32, 0, st>>>(q, L);
    return cudaGetLastError();
}
int traceback_warps_per_block() { return TB_WARPS; }

}  // namespace bn
