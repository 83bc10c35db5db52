// traceback_kernel.cu — gapped X-drop alignment WITH traceback on the device.
//
// Reference: BLAST_GappedAlignmentWithTraceback (core/blast_gapalign.c:3994-4155) -> Blast_SemiGappedAlign
// with score_only == FALSE (:711-745) -> ALIGN_EX (:350-709), as Blast_TracebackFromHSPList calls it for
// blastn with eDynProgTbck (core/blast_traceback.c:565-571) on the blastna subject (one base per byte).
//
// Here the subject stays in its packed ncbi2na form in HBM (ambiguity runs of a BLAST DB volume are laid over it where a base is read: SubjAmb; otherwise the
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
constexpr int TBK = 4;                    // band cells per lane in a row segment
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

// ---- subject ambiguity ------------------------------------------------------------------------------------
// The traceback stage of the reference reads its subjects in blastna: the 2-bit bases with the volume's ambiguity
// runs laid over them (CSeqDBVol::x_GetAmbigSeq, objtools/blast/seqdb_reader/seqdbvol.cpp:832-870, 1565-1640).  Here
// the volume stays packed and the runs of the item's sequence ({first base, end, blastna code}, sorted, disjoint)
// are consulted where a base is read; sequences without runs (nearly all) take the packed path untouched.
struct SubjAmb {
    const int4 *runs;
    int32_t n;
    int64_t origin;          // index (in the coordinates of the reads) of the sequence's base 0
};
__device__ __forceinline__ SubjAmb subj_amb(const int4 *all_runs, int32_t first, int32_t n, int64_t origin)
{
    SubjAmb a;
    a.runs = (n > 0 && all_runs) ? all_runs + first : nullptr; a.n = a.runs ? n : 0; a.origin = origin;
    return a;
}
// blastna code of an ambiguous base, -1 for an ordinary one
__device__ int amb_code(const SubjAmb &A, int64_t at)
{
    const int32_t pos = (int32_t)(at - A.origin);
    int32_t lo = 0, hi = A.n;                       // first run that ends behind pos
    while (lo < hi) {
        const int32_t m = (lo + hi) >> 1;
        if (__ldg(&A.runs[m].y) > pos) hi = m; else lo = m + 1;
    }
    if (lo < A.n) {
        const int4 r = __ldg(&A.runs[lo]);
        if (r.x <= pos) return r.z;
    }
    return -1;
}
__device__ __forceinline__ int subj_code(const uint8_t *s, const SubjAmb &A, int64_t at)
{
    if (A.n) { const int c = amb_code(A, at); if (c >= 0) return c; }
    return sbase64(s, at);
}
// Mismatch flags of a 16-base comparison (mismatch_bits) corrected for ambiguous subject bases: the reference compares
// blastna bytes, so an ambiguous subject base differs from A/C/G/T and equals the same code in the query.
__device__ uint32_t amb_fix16(const DevQuery &q, const SubjAmb &A, int32_t qpos, int64_t at, uint32_t m, bool equal_codes_match = true)
{
    const int32_t pos = (int32_t)(at - A.origin);
    int32_t lo = 0, hi = A.n;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (__ldg(&A.runs[mid].y) > pos) hi = mid; else lo = mid + 1;
    }
    for (; lo < A.n; lo++) {
        const int4 r = __ldg(&A.runs[lo]);
        if (r.x >= pos + 16) break;
        for (int32_t b = max(r.x, pos); b < min(r.y, pos + 16); b++) {
            const int j = b - pos;
            const int32_t qp = qpos + j;
            const int qc = (qp >= -1 && qp <= q.concat_len) ? (int)__ldg(q.query + qp) : 15;
            const uint32_t bit = 1u << (30 - 2 * j);
            if (qc == r.z && equal_codes_match) m &= ~bit; else m |= bit;
        }
    }
    return m;
}
__device__ __forceinline__ uint32_t subj_mismatch16(const DevQuery &q, const uint8_t *packed, const SubjAmb &A, int32_t qpos,
                                                    int64_t at, uint32_t qb, uint32_t qa, bool equal_codes_match = true)
{
    uint32_t m = mismatch_bits(qb, qa, swin(packed, at));
    if (A.n) m = amb_fix16(q, A, qpos, at, m, equal_codes_match);
    return m;
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
// status: 0 ok, 1 band wider than the ring of C cells (shared memory; the wide variant's is in global memory), 3 arena exhausted.
template <int C>
__device__ int32_t align_ex_warp(const TracebackLaunch &L, const uint8_t *qp, int q_inc, const uint8_t *S, const SubjAmb &amb, int64_t s0, int s_inc,
                                 int32_t M, int32_t N, const int32_t *matrix, int32_t gap_open, int32_t gap_extend,
                                 int32_t x_dropoff, int2 *ring, uint8_t *pf, RowStore &rs, int32_t &a_offset, int32_t &b_offset,
                                 int &status, int lane)
{
    const int32_t goe = gap_open + gap_extend, ge = gap_extend;
    constexpr int32_t MASK = C - 1;
    a_offset = 0; b_offset = 0;
    if (x_dropoff < goe) x_dropoff = goe;
    if (N <= 0 || M <= 0) return 0;
    const uint32_t lt = (1u << lane) - 1u;
    const int32_t num_extra = x_dropoff / ge + 3;
    uint8_t *arena = L.arena;
    // rows are taken from the arena in chunks: about what the whole extension needs (M rows of a band of ~2 X / ge
    // cells), between 2 KB (a short read) and ROW_CHUNK
    const long long chunk = min(ROW_CHUNK, max(2048ll, ((long long)M + 2) * (2 * num_extra + 40)));

    // row 0: cell 0 = (0, -goe); cells i >= 1 = (-goe - (i-1) ge, that - goe) while the score is >= -X; script GAP_IN_A
    int32_t b_size;
    {
        int32_t k = min((x_dropoff - goe) / ge + 1, N);
        if (k + 2 >= C) { status = 1; return 0; }
        if (rs.cur + k + 2 > rs.end) {
            long long at = 0;
            if (lane == 0) at = arena_alloc(L, chunk);
            at = __shfl_sync(FULLW, at, 0);
            if (at < 0) { status = 3; return 0; }
            rs.cur = at; rs.end = at + chunk;
        }
        for (int32_t i = lane; i <= k; i += 32) {
            const int32_t sc = (i == 0) ? 0 : -goe - (i - 1) * ge;
            ring[i & MASK] = make_int2(sc, sc - goe);
            arena[rs.cur + i] = SCRIPT_GAP_IN_A;
        }
        if (lane == 0) { rs.row_off[0] = rs.cur; rs.row_first[0] = 0; }
        for (int32_t i = lane; i < C; i += 32) pf[i] = 0;      // first guess of the prune flags: the previous row's
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
                const long long sz = need > chunk ? need : chunk;
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

        // TBK consecutive cells per lane: a segment is 32 TBK cells (the usual band fits in one), the two prefix
        // maxima run over per-lane partial results, the cells of a lane are chained in registers
        for (int32_t seg = row_first; seg < row_end; seg += 32 * TBK) {
            const int32_t b0 = seg + lane * TBK;
            int2 cell[TBK];
            bool act[TBK];
            uint32_t amask[TBK], p[TBK];
            int32_t v[TBK], dg[TBK];
#pragma unroll
            for (int j = 0; j < TBK; j++) {
                act[j] = b0 + j < row_end;
                cell[j] = act[j] ? ring[(b0 + j) & MASK] : make_int2(MININT, MININT);
                amask[j] = __ballot_sync(FULLW, act[j]);
            }
            int32_t up0 = __shfl_up_sync(FULLW, cell[TBK - 1].x, 1);
            if (lane == 0) up0 = prev_old_best;
            prev_old_best = __shfl_sync(FULLW, cell[TBK - 1].x, 31);
#pragma unroll
            for (int j = 0; j < TBK; j++) {
                const int32_t b = b0 + j;
                const int32_t up = j == 0 ? up0 : cell[j > 0 ? j - 1 : 0].x;
                v[j] = NEGINF; dg[j] = MININT;
                if (act[j]) {
                    if (b != row_first) dg[j] = up + mrow[subj_code(S, amb, s0 + (int64_t)b * s_inc)];
                    v[j] = max(dg[j], cell[j].y);
                }
                // first guess: what became of the cell's diagonal predecessor in the previous row (the band follows
                // the diagonal); any guess converges to the same flags, a good one in one or two passes
                p[j] = __ballot_sync(FULLW, act[j] && pf[(b - 1) & MASK] != 0);
            }
            // ---- fixed point over the prune flags (cell order = lane-major) -------------------------------
            bool un[TBK];
            int32_t R[TBK], sc[TBK];
            for (;;) {
                int32_t ubase = 0;
#pragma unroll
                for (int j = 0; j < TBK; j++) ubase += __popc(amask[j] & ~p[j] & lt);
                int32_t u[TBK], w[TBK];
                int32_t running = ubase, lw = NEGINF;
#pragma unroll
                for (int j = 0; j < TBK; j++) {
                    un[j] = ((amask[j] & ~p[j]) >> lane) & 1u;
                    u[j] = running;
                    running += un[j] ? 1 : 0;
                    w[j] = un[j] ? v[j] - goe + ge * (u[j] + 1) : NEGINF;
                    lw = max(lw, w[j]);
                }
                int32_t run = max(r_in, excl_prefix_max(lw, lane));
                int32_t ls = NEGINF;
#pragma unroll
                for (int j = 0; j < TBK; j++) {
                    R[j] = run - ge * u[j];
                    sc[j] = max(v[j], R[j]);
                    run = max(run, w[j]);
                    if (un[j]) ls = max(ls, sc[j]);
                }
                int32_t runb = max(best_in, excl_prefix_max(ls, lane));
                bool same = true;
#pragma unroll
                for (int j = 0; j < TBK; j++) {
                    const uint32_t pn = __ballot_sync(FULLW, act[j] && (runb - sc[j] > x_dropoff));
                    if (un[j]) runb = max(runb, sc[j]);
                    same = same && (pn == p[j]);
                    p[j] = pn;
                }
                if (same) break;
            }
            // ---- script bytes (core/blast_gapalign.c:560-612) --------------------------------------------
            uint32_t mine = 0;                                  // my unpruned cells, bit j
            int32_t ls = NEGINF;
            uint8_t pfn[TBK];
#pragma unroll
            for (int j = 0; j < TBK; j++) {
                if (act[j]) {
                    uint8_t op = (v[j] < R[j]) ? SCRIPT_GAP_IN_A : ((dg[j] < cell[j].y) ? SCRIPT_GAP_IN_B : SCRIPT_SUB);
                    if (un[j]) {
                        if (cell[j].y - ge >= sc[j] - goe) op += SCRIPT_EXTEND_GAP_B;
                        if (R[j] - ge >= sc[j] - goe) op += SCRIPT_EXTEND_GAP_A;
                    }
                    srow[b0 + j] = op;
                }
                if (un[j]) { mine |= 1u << j; ls = max(ls, sc[j]); }
                pfn[j] = un[j] ? 0 : 1;
            }
            // the guesses of this segment were all read before the first ballot; the next segment of this row reads
            // pf[seg + 32 TBK - 1 ...], the flag of THIS row's last cell here instead of the previous row's: still a guess
#pragma unroll
            for (int j = 0; j < TBK; j++) if (act[j]) pf[(b0 + j) & MASK] = pfn[j];
            // ---- commit the segment ----------------------------------------------------------------------
            const int nact = min(32 * TBK, row_end - seg);
            const uint32_t anyu = __ballot_sync(FULLW, mine != 0);
            int dropped = 0;
            if (lead) {
                if (anyu) {
                    const int l1 = __ffs(anyu) - 1;
                    const uint32_t m1 = __shfl_sync(FULLW, mine, l1);
                    dropped = l1 * TBK + (__ffs(m1) - 1);
                } else dropped = nact;
                new_first += dropped;
                lead = (dropped == nact);
            }
#pragma unroll
            for (int j = 0; j < TBK; j++) {
                if (act[j]) {
                    if (un[j]) ring[(b0 + j) & MASK] = make_int2(sc[j], max(sc[j] - goe, cell[j].y - ge));
                    else if (lane * TBK + j >= dropped) ring[(b0 + j) & MASK] = make_int2(MININT, cell[j].y);
                }
            }
            if (anyu) {
                const int l2 = 31 - __clz(anyu);
                const uint32_t m2 = __shfl_sync(FULLW, mine, l2);
                last_b = seg + l2 * TBK + (31 - __clz(m2));
                const int32_t m = __reduce_max_sync(FULLW, ls);
                if (m > best_in) {
                    int cand = -1;
#pragma unroll
                    for (int j = TBK - 1; j >= 0; j--) if (un[j] && sc[j] == m) cand = j;
                    const uint32_t at = __ballot_sync(FULLW, cand >= 0);
                    const int l3 = __ffs(at) - 1;
                    const int j3 = __shfl_sync(FULLW, cand, l3);
                    best_in = m; a_offset = a_index; b_offset = seg + l3 * TBK + j3;
                }
            }
            {   // gap_row leaving the segment: state after its last active cell
                const int lr = (nact - 1) / TBK, jr = (nact - 1) % TBK;
                int32_t rn = 0;
#pragma unroll
                for (int j = 0; j < TBK; j++)
                    if (j == jr) rn = un[j] ? max(sc[j] - goe, R[j] - ge) : R[j];
                r_in = __shfl_sync(FULLW, rn, lr);
            }
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
// CELLS = band cells a warp can hold: TB_CELLS in shared memory, or (WIDE) TB_WIDE_CELLS in global memory for the rare
// batch with a band beyond that (2 X_final / gap_extend above ~1000: gap_extend 1 with a large final X-drop).
constexpr int TB_WIDE_CELLS = 32768;
template <int CELLS, bool WIDE>
__global__ void __launch_bounds__(TB_WARPS * 32)
traceback_dp_kernel(const DevQuery q, const TracebackLaunch L)
{
    __shared__ int2 rings[WIDE ? 1 : TB_WARPS][WIDE ? 1 : CELLS];
    __shared__ uint8_t s_pf[WIDE ? 1 : TB_WARPS][WIDE ? 1 : CELLS];
    __shared__ int32_t s_matrix[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_matrix[i] = q.matrix[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * TB_WARPS + wib, nwarps = (int64_t)gridDim.x * TB_WARPS;
    int2 *ring = WIDE ? L.wide_ring + warp0 * CELLS : rings[wib];
    uint8_t *pf = WIDE ? L.wide_pf + warp0 * CELLS : s_pf[wib];
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
                const SubjAmb amb = subj_amb(L.amb_runs, it.amb_first, it.amb_n, 0);      // S-relative reads
                out.score = align_ex_warp<CELLS>(L, qp, q_inc, S, amb, s0, s_inc, M, N, s_matrix, q.gap_open, q.gap_extend,
                                                 L.x_dropoff, ring, pf, rs, a_off, b_off, status, lane);
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
                        {
                            // The walk is a chain of dependent loads (row table, then the script byte).  All lanes keep
                            // the walk state; lane j prefetches the byte the walk needs after j diagonal steps from the
                            // batch's first cell, (a0 - j, b0 - j) — an alignment mostly moves diagonally — and the steps are
                            // fed from registers by shuffles until the path leaves that diagonal (a gap) or 32 steps are done.
                            int32_t a = a_off, b = b_off, run_op = -1, run_n = 0;
                            uint8_t script = SCRIPT_SUB;
                            const volatile uint8_t *ar = L.arena;
                            while (a > 0 || b > 0) {
                                const int32_t a0 = a, b0 = b;
                                const int32_t aj = a0 - lane, bj = b0 - lane;
                                uint32_t pre = 0x100u;                        // bit 8 = not available
                                if (aj >= 0 && bj >= 0) {
                                    const long long off = rs.row_off[aj];
                                    const int32_t first = rs.row_first[aj];
                                    if (bj >= first && bj - first < CELLS + 64 && off + (bj - first) < L.arena_bytes)
                                        pre = ar[off + (bj - first)];
                                }
                                for (;;) {
                                    const int j = a0 - a;
                                    if (j >= 32 || b0 - b != j) break;        // off the prefetched diagonal
                                    uint32_t nx = __shfl_sync(FULLW, pre, j);
                                    if (nx & 0x100u) nx = ar[rs.row_off[a] + (b - rs.row_first[a])];
                                    const uint8_t next = (uint8_t)nx;
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
                                        if (run_n && lane == 0) runs[n_runs] = make_int2(run_op, run_n);
                                        if (run_n) n_runs++;
                                        run_op = script; run_n = 1;
                                    }
                                    if (!(a > 0 || b > 0)) break;
                                }
                            }
                            if (run_n && lane == 0) runs[n_runs] = make_int2(run_op, run_n);
                            if (run_n) n_runs++;
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

// ================================================================================================
// Greedy alignment with traceback: BLAST_GreedyGappedAlignment with do_traceback (core/blast_gapalign.c:2620-2751)
// -> BLAST_AffineGreedyAlign prologue (core/greedy_align.c:799-814) -> BLAST_GreedyAlign (:385-752) with an edit
// block, then Blast_PrelimEditBlockToGapEditScript (core/blast_gapalign.c:2455-2517) and s_ReduceGaps (:2545-2617).
// One WARP per alignment, one lane per diagonal (see below).  Unlike the
// score-only kernel every row last_seq2_off[d][] is kept — including the sentinel overwrites of row d-1 the
// reference makes before it fills row d, which the traceback reads back — in a per-warp arena that is reset for
// every alignment.  Arena layout (ints): [row table][max_score][rows ...   ... rev list][fwd list].
// ================================================================================================
namespace {

constexpr int32_t GREEDY_MAX_COST = 10000;
constexpr int32_t GREEDY_INVALID = -2;          // kInvalidOffset

struct GreedyTbSeq {
    const DevQuery *q;
    const uint8_t *packed;
    int32_t qbase;      // concatenated-query position of seq1[0] (both directions: position of the window start)
    int64_t sbase;      // absolute volume base of seq2[0]
    int32_t len1, len2;
    bool reverse;
    SubjAmb amb;        // ambiguity runs of the subject sequence (origin = its absolute base index)
};
__device__ __forceinline__ int32_t tb_first_mismatch(const GreedyTbSeq &p, int32_t i1, int32_t i2)
{
    const int32_t n = min(p.len1 - i1, p.len2 - i2);
    if (n <= 0) return 0;
    if (p.amb.n) {          // s_FindFirstMismatch on blastna bytes (core/greedy_align.c:318-380), 16 bases at a time; an
                            // ambiguity code in the query never matches there, not even the same code in the subject
        int32_t cnt = 0;
        while (cnt < n) {
            uint32_t qb, qa;
            if (p.reverse) {
                const int32_t qpos = p.qbase + p.len1 - i1 - cnt - 16;
                qwin(*p.q, qpos, qb, qa);
                const uint32_t m = subj_mismatch16(*p.q, p.packed, p.amb, qpos, p.sbase + p.len2 - i2 - cnt - 16, qb, qa, false);
                if (m) return min(cnt + ((__ffs(m) - 1) >> 1), n);
            } else {
                const int32_t qpos = p.qbase + i1 + cnt;
                qwin(*p.q, qpos, qb, qa);
                const uint32_t m = subj_mismatch16(*p.q, p.packed, p.amb, qpos, p.sbase + i2 + cnt, qb, qa, false);
                if (m) return min(cnt + (__clz(m) >> 1), n);
            }
            cnt += 16;
        }
        return n;
    }
    if (p.reverse) return match_run_rev(*p.q, p.packed, p.qbase + p.len1 - i1, p.sbase + p.len2 - i2, n);
    return match_run_fwd(*p.q, p.packed, p.qbase + i1, p.sbase + i2, n);
}

struct OpList {                 // GapPrelimEditBlock: list grows DOWNWARD from `top` (entry i at top - 2 (i + 1))
    int32_t *top;
    int32_t n, last_op;
    __device__ __forceinline__ void add(int32_t op, int32_t num)    // GapPrelimEditBlockAdd (core/gapinfo.c:179-190)
    {
        if (num == 0) return;
        if (last_op == op) top[-2 * n + 1] += num;
        else { top[-2 * (n + 1)] = op; top[-2 * (n + 1) + 1] = num; n++; last_op = op; }
    }
    __device__ __forceinline__ int32_t op(int32_t i) const { return top[-2 * (i + 1)]; }
    __device__ __forceinline__ int32_t num(int32_t i) const { return top[-2 * (i + 1) + 1]; }
};

// s_ReduceGaps (core/blast_gapalign.c:2545-2617) on {op[], num[]}; returns the new size
__device__ int32_t reduce_gaps(const DevQuery &q, const uint8_t *packed, const SubjAmb &amb, int32_t ctx_off, int32_t qi, int64_t si,
                               int32_t *op, int32_t *num, int32_t size)
{
    const uint8_t *Qb = q.query + ctx_off;
    auto eq = [&](int32_t a, int64_t b) -> bool { return (int)__ldg(Qb + a) == subj_code(packed, amb, b); };
    for (int32_t i = 0; i < size; i++) {
        if (op[i] == 3) { qi += num[i]; si += num[i]; continue; }
        if (i > 1 && op[i] != op[i - 2] && num[i - 2] > 0) {
            int32_t d = num[i] + num[i - 1] + num[i - 2];
            if (d == 3) {
                num[i - 2] = 0; num[i - 1] = 2; num[i] = 0;
                if (op[i] == 6) ++qi; else ++si;
            } else if (d < 12) {
                int32_t nm1 = 0, nm2 = 0;
                d = min(num[i], num[i - 2]);
                qi -= num[i - 1]; si -= num[i - 1];
                int32_t q1 = qi; int64_t s1 = si;
                if (op[i] == 6) si -= d; else qi -= d;
                for (int32_t j = 0; j < num[i - 1]; ++j, ++q1, ++s1, ++qi, ++si) {
                    if (eq(q1, s1)) nm1++;
                    if (eq(qi, si)) nm2++;
                }
                for (int32_t j = 0; j < d; ++j, ++qi, ++si) if (eq(qi, si)) nm2++;
                if (nm2 >= nm1 - d) { num[i - 2] -= d; num[i - 1] += d; num[i] -= d; }
                else { qi = q1; si = s1; }
            }
        }
        if (op[i] == 6) qi += num[i]; else si += num[i];
    }
    int32_t j = 0;
    for (int32_t i = 0; i < size; i++) {
        if (num[i] > 0) { num[j] = num[i]; op[j] = op[i]; ++j; }
        else if (++i < size && j > 0) num[j - 1] += num[i];
    }
    return j;
}

}  // namespace

// ---- warp-parallel greedy with traceback -----------------------------------------------------------------------
// One warp per alignment, one lane per diagonal of the current distance (the formulation of greedy_align_warp in
// gapped_kernel.cu: every diagonal's new offset depends on row d-1 only; the order-dependent bookkeeping is replayed
// from ballots in ascending k).  Rows are kept in the warp's arena; the sentinel overwrites of row d-1 are made
// literally, so the walk reads back exactly what the reference's s_GetNextNonAffineTback reads.
namespace {

__device__ int32_t tb_first_mismatch_warp(const GreedyTbSeq &p, int32_t i1, int32_t i2, int lane)
{
    const int32_t n = min(p.len1 - i1, p.len2 - i2);
    if (n <= 0) return 0;
    for (int32_t base = 0; base < n; base += 512) {
        const int32_t off = base + 16 * lane;
        int32_t c = 16;
        if (off < n) {
            uint32_t qb, qa, m;
            if (p.reverse) {
                qwin(*p.q, p.qbase + p.len1 - i1 - off - 16, qb, qa);
                m = subj_mismatch16(*p.q, p.packed, p.amb, p.qbase + p.len1 - i1 - off - 16, p.sbase + p.len2 - i2 - off - 16, qb, qa, false);
                if (m) c = (__ffs(m) - 1) >> 1;
            } else {
                qwin(*p.q, p.qbase + i1 + off, qb, qa);
                m = subj_mismatch16(*p.q, p.packed, p.amb, p.qbase + i1 + off, p.sbase + i2 + off, qb, qa, false);
                if (m) c = __clz(m) >> 1;
            }
        }
        const unsigned stop = __ballot_sync(FULLW, off < n && c < 16);
        if (stop) {
            const int src = __ffs(stop) - 1;
            const int32_t cc = __shfl_sync(FULLW, c, src);
            return min(base + 16 * src + cc, n);
        }
    }
    return n;
}

// A[0 .. cap): the warp's free arena (ints).  ed: lane 0's list.  Warp-uniform return value and status.
__device__ int32_t greedy_align_tb_warp(const GreedyTbSeq &sp, int32_t xdrop_threshold, int32_t match_cost, int32_t mismatch_cost,
                                        int32_t &seq1_len, int32_t &seq2_len, int32_t *A, int64_t cap, OpList &ed, int &status,
                                        int lane)
{
    const int32_t len1 = sp.len1, len2 = sp.len2;
    int32_t best_dist = 0, best_diag = 0;
    const int32_t max_dist = min(GREEDY_MAX_COST, len2 / 2 + 1);
    const int32_t origin = max_dist + 2;
    const int32_t xdrop_offset = (xdrop_threshold + match_cost / 2) / (match_cost + mismatch_cost) + 1;

    const int32_t index0 = tb_first_mismatch_warp(sp, 0, 0, lane);
    seq1_len = index0; seq2_len = index0;
    if (index0 == len1 || index0 == len2) { if (lane == 0) ed.add(3, index0); return 0; }

    const int64_t tcap = min((int64_t)max_dist + 2, cap / 8);
    const int64_t mcap = min((int64_t)max_dist + 1 + xdrop_offset, cap / 8);
    if (cap < 64 || mcap <= xdrop_offset + 1) { status = 3; return 0; }
    int32_t *rowtab = A;
    int32_t *max_score_mem = A + tcap;
    int64_t top = tcap + mcap;
    int32_t *max_score = max_score_mem + xdrop_offset;
    for (int32_t i = lane; i < xdrop_offset; i += 32) max_score_mem[i] = 0;
    // rows 0 and 1
    if (2 >= tcap || top + 9 + 11 > cap) { status = 3; return 0; }
    if (lane == 0) {
        rowtab[0] = (int32_t)(top - (origin - 4));
        rowtab[1] = (int32_t)(top + 9 - (origin - 5));
        (A + rowtab[0])[origin] = index0;
        max_score[0] = index0 * match_cost;
    }
    top += 9 + 11;
    __syncwarp();
    int32_t diag_lower = origin - 1, diag_upper = origin + 1;
    bool end1_reached = false, end2_reached = false;

    for (int32_t d = 1; d <= max_dist; d++) {
        if (d + xdrop_offset >= mcap) { status = 3; return 0; }
        int32_t curr_extent = 0, curr_seq2_index = 0, curr_diag = 0;
        const int32_t tmp_lower = diag_lower, tmp_upper = diag_upper;
        int32_t *prev = A + rowtab[d - 1];
        int32_t *cur = A + rowtab[d];
        if (lane == 0) {
            prev[diag_lower - 1] = GREEDY_INVALID;
            prev[diag_lower] = GREEDY_INVALID;
            prev[diag_upper] = GREEDY_INVALID;
            prev[diag_upper + 1] = GREEDY_INVALID;
        }
        __syncwarp();
        int32_t xdrop_score = max_score[d - xdrop_offset] + (match_cost + mismatch_cost) * d - xdrop_threshold;
        {
            const int32_t h = match_cost / 2;
            int32_t qd = xdrop_score / h;
            if (xdrop_score % h > 0) ++qd;
            xdrop_score = qd;
        }
        for (int32_t kb = tmp_lower; kb <= tmp_upper; kb += 32) {
            const int32_t k = kb + lane;
            const bool active = k <= tmp_upper;
            bool ok = false;
            int32_t seq1_index = 0, seq2_index = 0;
            if (active) {
                seq2_index = max(prev[k + 1], prev[k]) + 1;
                seq2_index = max(seq2_index, prev[k - 1]);
                seq1_index = seq2_index + k - origin;
                ok = !(seq2_index < 0 || seq1_index + seq2_index < xdrop_score);
                if (ok) {
                    const int32_t run = tb_first_mismatch(sp, seq1_index, seq2_index);
                    seq1_index += run; seq2_index += run;
                }
            }
            const unsigned act = __ballot_sync(FULLW, active);
            const unsigned succ = __ballot_sync(FULLW, ok);
            const unsigned e2 = __ballot_sync(FULLW, ok && seq2_index == len2);
            const unsigned e1 = __ballot_sync(FULLW, ok && seq1_index == len1);
            unsigned inv = 0;
            if (e2 == 0) {
                // no diagonal reached the end of seq2 in this round (the usual case): diag_lower only moves over the
                // leading run of failures (and only if it still sits on the round's first diagonal), every later
                // failure is marked invalid, diag_upper follows the last success (same closed form as greedy_kernel)
                const unsigned fail = act & ~succ;
                unsigned lead = 0;
                if (diag_lower == kb) {
                    lead = (unsigned)__ffs(~fail) - 1u;
                    if (fail == 0xffffffffu) lead = 32;
                    diag_lower += (int32_t)lead;
                }
                inv = lead >= 32 ? 0u : (fail & ~((1u << lead) - 1u));
                if (succ) {
                    const int top = 31 - __clz(succ);
                    diag_upper = kb + top;
                    if (e1) {
                        end1_reached = true;
                        if ((e1 >> top) & 1u) diag_upper = kb + top - 1;
                    }
                }
            } else {
                for (unsigned rem = act; rem; rem &= rem - 1) {    // the bookkeeping in ascending k
                    const int b = __ffs(rem) - 1;
                    const unsigned bit = 1u << b;
                    const int32_t kk = kb + b;
                    if (!(succ & bit)) {
                        if (kk == diag_lower) diag_lower++;
                        else inv |= bit;
                    } else {
                        diag_upper = kk;
                        if (e2 & bit) { diag_lower = kk + 1; end2_reached = true; }
                        if (e1 & bit) { diag_upper = kk - 1; end1_reached = true; }
                    }
                }
            }
            if (ok) cur[k] = seq2_index;
            else if (inv & (1u << lane)) cur[k] = GREEDY_INVALID;
            if (succ) {     // extent: first diagonal with the strict maximum of seq1 + seq2
                const int32_t ext = ok ? seq1_index + seq2_index : -1;
                const int32_t best_ext = __reduce_max_sync(FULLW, ext);
                if (best_ext > curr_extent) {
                    const int src = __ffs(__ballot_sync(FULLW, ok && ext == best_ext)) - 1;
                    curr_extent = best_ext;
                    curr_seq2_index = __shfl_sync(FULLW, seq2_index, src);
                    curr_diag = kb + src;
                }
            }
        }
        const int32_t curr_score = curr_extent * (match_cost / 2) - d * (match_cost + mismatch_cost);
        const int32_t prev_best = max_score[d - 1];
        __syncwarp();
        if (curr_score > prev_best) {
            if (lane == 0) max_score[d] = curr_score;
            best_dist = d;
            best_diag = curr_diag;
            seq2_len = curr_seq2_index;
            seq1_len = curr_seq2_index + curr_diag - origin;
        } else if (lane == 0) max_score[d] = prev_best;
        if (diag_lower > diag_upper) { __syncwarp(); break; }
        if (!end2_reached) diag_lower--;
        if (!end1_reached) diag_upper++;
        {   // row d + 1 addressable on [diag_lower - 2, diag_upper + 4]
            const int64_t w = (int64_t)diag_upper - diag_lower + 7;
            if (d + 1 >= tcap || top + w > cap) { status = 3; return 0; }
            if (lane == 0) rowtab[d + 1] = (int32_t)(top - (diag_lower - 2));
            top += w;
        }
        __syncwarp();
    }
    // ---- traceback by lane 0 (core/greedy_align.c:687-751) ----------------------------------------------------
    int bad = 0;
    if (lane == 0) {
        int32_t d = best_dist;
        int32_t seq2_index = seq2_len;
        if ((int64_t)(ed.top - A) - 2 * (int64_t)(ed.n + 2 * d + 2) < top) bad = 1;
        else {
            while (d > 0) {
                const int32_t *pr = A + rowtab[d - 1];
                int32_t new_diag, new_seq2_index;
                if (pr[best_diag - 1] > max(pr[best_diag], pr[best_diag + 1])) { new_seq2_index = pr[best_diag - 1]; new_diag = best_diag - 1; }
                else if (pr[best_diag] > pr[best_diag + 1]) { new_seq2_index = pr[best_diag]; new_diag = best_diag; }
                else { new_seq2_index = pr[best_diag + 1]; new_diag = best_diag + 1; }
                if (new_diag == best_diag) {
                    if (seq2_index - new_seq2_index > 0) ed.add(3, seq2_index - new_seq2_index);
                } else if (new_diag < best_diag) {
                    if (seq2_index - new_seq2_index > 0) ed.add(3, seq2_index - new_seq2_index);
                    ed.add(6, 1);
                } else {
                    if (seq2_index - new_seq2_index - 1 > 0) ed.add(3, seq2_index - new_seq2_index - 1);
                    ed.add(0, 1);
                }
                d--;
                best_diag = new_diag;
                seq2_index = new_seq2_index;
            }
            ed.add(3, index0);
        }
    }
    bad = __shfl_sync(FULLW, bad, 0);
    if (bad) { status = 3; return 0; }
    return best_dist;
}

}  // namespace

__global__ void __launch_bounds__(128)
traceback_greedy_warp_kernel(const DevQuery q, const TracebackLaunch L)
{
    const int lane = threadIdx.x & 31;
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t cap = L.arena_bytes / 4 / nw;                 // ints per warp
    int32_t *A = reinterpret_cast<int32_t *>(L.arena) + w0 * cap;
    int32_t match = q.reward, mismatch = -q.penalty, xd = L.x_dropoff;
    if (match % 2 == 1) { match *= 2; mismatch *= 2; xd *= 2; }
    for (int64_t w = w0; w < L.n; w += nw) {
        if (L.todo && !L.todo[w]) continue;
        const DevTracebackItem it = L.items[w];
        const DevContext c = q.ctx[it.context];
        const int32_t q_length = c.query_length, s_length = it.s_length;
        const int32_t q_off = it.q_start, s_off = it.s_start;
        const int64_t seq_base = it.byte_off * 4 + it.s_shift;
        DevTracebackDir out, out2;
        out.score = 0; out.a_off = 0; out.b_off = 0; out.ops_off = 0; out.n_ops = 0; out.status = 0; out.ran = 1; out.pad = 0;
        out2 = out;
        int status = 0;
        int32_t q_ext_r = 0, s_ext_r = 0, q_ext_l = 0, s_ext_l = 0;
        OpList fwd, rev;
        fwd.top = A + cap; fwd.n = 0; fwd.last_op = 8;
        rev = fwd;
        GreedyTbSeq sp;
        sp.q = &q; sp.packed = L.packed;
        sp.amb = subj_amb(L.amb_runs, it.amb_first, it.amb_n, it.byte_off * 4);
        sp.qbase = c.query_offset + q_off; sp.sbase = seq_base + s_off;
        sp.len1 = q_length - q_off; sp.len2 = s_length - s_off; sp.reverse = false;
        __syncwarp();
        int32_t dist = greedy_align_tb_warp(sp, xd, match, mismatch, q_ext_r, s_ext_r, A, cap, fwd, status, lane);
        if (!status) {
            const int32_t fn = __shfl_sync(FULLW, fwd.n, 0);
            rev.top = fwd.top - 2 * fn; rev.n = 0; rev.last_op = 8;
            sp.qbase = c.query_offset; sp.sbase = seq_base; sp.len1 = q_off; sp.len2 = s_off; sp.reverse = true;
            __syncwarp();
            dist += greedy_align_tb_warp(sp, xd, match, mismatch, q_ext_l, s_ext_l, A, (int64_t)(rev.top - A), rev, status, lane);
        }
        __syncwarp();
        if (!status && lane == 0) {
            const int32_t score = (q_ext_r + s_ext_r + q_ext_l + s_ext_l) * q.reward / 2 - dist * (q.reward - q.penalty);
            int32_t size = fwd.n + rev.n;
            const bool merge = fwd.n > 0 && rev.n > 0 && fwd.op(fwd.n - 1) == rev.op(rev.n - 1);
            if (merge) size--;
            if (2 * (int64_t)size + 2 > (int64_t)(rev.top - 2 * rev.n - A)) status = 3;
            else {
                int32_t *op = A, *num = A + size + 1;
                int32_t idx = 0;
                for (int32_t i = 0; i < rev.n; i++) { op[idx] = rev.op(i); num[idx] = rev.num(i); idx++; }
                if (fwd.n > 0) {
                    int32_t i = fwd.n - 1;
                    if (merge) { num[idx - 1] += fwd.num(fwd.n - 1); i = fwd.n - 2; }
                    for (; i >= 0; i--) { op[idx] = fwd.op(i); num[idx] = fwd.num(i); idx++; }
                }
                size = reduce_gaps(q, L.packed, sp.amb, c.query_offset, q_off - q_ext_l, seq_base + s_off - s_ext_l, op, num, size);
                const unsigned long long base = atomicAdd(L.ops_used, (unsigned long long)size);
                if ((long long)(base + size) > L.ops_cap) status = 4;
                else {
                    for (int32_t i = 0; i < size; i++) L.ops[base + i] = make_int2(op[i], num[i]);
                    out.ops_off = (long long)base; out.n_ops = size;
                }
                out.score = score;
                out.a_off = q_ext_l; out.b_off = s_ext_l;
                out2.a_off = q_ext_r; out2.b_off = s_ext_r;
            }
        }
        if (lane == 0) {
            out.status = status; out2.status = status;
            L.out[2 * w] = out;
            L.out[2 * w + 1] = out2;
        }
        __syncwarp();
    }
}

// ---- affine greedy with traceback (BLAST_AffineGreedyAlign body with an edit block, core/greedy_align.c:817-1237;
// s_GetNextAffineTbackFromMatch :150-182, s_GetNextAffineTbackFromIndel :201-262).  One THREAD per alignment (the
// option is off in every blastn task default); every row {insert, match, delete} x diagonal is kept. ------------------
namespace {

__device__ int32_t greedy_align_affine_tb(const GreedyTbSeq &sp, const AffineCosts &ac, int32_t &seq1_len, int32_t &seq2_len,
                                          int32_t *A, int64_t cap, OpList &ed, int &status)
{
    const int32_t kInvalidDiag = 100000000;
    const int32_t len1 = sp.len1, len2 = sp.len2;
    const int32_t match_half = ac.match / 2;
    const int32_t op_cost = ac.op_cost, gap_open = ac.gap_open, gap_extend = ac.gap_extend, goe = ac.gap_open + ac.gap_extend;
    const int32_t max_penalty = ac.max_penalty;
    const int32_t max_dist = min(GREEDY_MAX_COST, len2 / 2 + 1);
    const int64_t scaled_max_dist = (int64_t)max_dist * gap_extend;
    const int32_t origin = max_dist + 2;

    int32_t index = tb_first_mismatch(sp, 0, 0);
    seq1_len = index; seq2_len = index;
    int32_t seq1_index = index, seq2_index;
    const int32_t first_run = index;
    if (index == len1 || index == len2) { ed.add(3, index); return index * ac.match; }

    // tables for distances 0 .. Tn-1: row table, diagonal bounds (max_penalty negative slots in front), best scores
    const int64_t Tmin = (int64_t)max_penalty + ac.xdrop_offset + 4;       // what the shortest extension already needs
    const int64_t Tn = min(scaled_max_dist + Tmin, cap / 16);
    if (Tn < Tmin || cap < 256) { status = 3; return 0; }
    int32_t *rowtab = A;
    int32_t *diag_lower = A + Tn + max_penalty;
    int32_t *diag_upper = diag_lower + Tn + max_penalty;
    int32_t *max_score_mem = diag_upper + Tn;
    int32_t *max_score = max_score_mem + ac.xdrop_offset;
    int64_t top = 4 * Tn + 2 * max_penalty + ac.xdrop_offset + 1;
    auto alloc_row = [&](int64_t d, int32_t lo, int32_t hi) -> bool {      // row d addressable on diagonals [lo, hi]
        const int64_t w = hi >= lo ? 3 * ((int64_t)hi - lo + 1) : 0;
        if (d >= Tn || top + w > cap) return false;
        rowtab[d] = (int32_t)(top - 3 * (int64_t)lo);
        top += w;
        return true;
    };
#define AFF(dd, kk, f) A[(int64_t)rowtab[(dd)] + 3 * (int64_t)(kk) + (f)]       // f: 0 insert, 1 match, 2 delete
    if (top >= cap) { status = 3; return 0; }
    for (int32_t i = 0; i < ac.xdrop_offset; i++) max_score_mem[i] = 0;
    for (int32_t i = 1; i <= max_penalty; i++) { diag_lower[-i] = kInvalidDiag; diag_upper[-i] = -kInvalidDiag; }
    for (int32_t dd = 0; dd <= max_penalty; dd++)
        if (!alloc_row(dd, origin - max_penalty - 3, origin + max_penalty + 3)) { status = 3; return 0; }
    AFF(0, origin, 1) = seq1_index;
    AFF(0, origin, 0) = GREEDY_INVALID;
    AFF(0, origin, 2) = GREEDY_INVALID;
    max_score[0] = seq1_index * ac.match;
    diag_lower[0] = origin; diag_upper[0] = origin;
    int32_t curr_lower = origin - 1, curr_upper = origin + 1;
    int32_t end1_diag = 0, end2_diag = 0, num_nonempty = 1;
    int32_t best_dist = 0, best_diag = 0;
    int64_t d = 1;

    while (d <= scaled_max_dist) {
        if (d + ac.xdrop_offset + 1 >= Tn) { status = 3; return 0; }
        int32_t curr_extent = 0, curr_seq2_index = 0, curr_diag = 0;
        const int32_t tmp_lower = curr_lower, tmp_upper = curr_upper;
        int32_t xdrop_score = max_score[d - ac.xdrop_offset] + ac.common_factor * (int32_t)d - ac.xdrop;
        {
            int32_t qd = xdrop_score / match_half;
            if (xdrop_score % match_half > 0) ++qd;
            xdrop_score = qd < 0 ? 0 : qd;
        }
        const int32_t lo_goe = diag_lower[d - goe], up_goe = diag_upper[d - goe];
        const int32_t lo_ge = diag_lower[d - gap_extend], up_ge = diag_upper[d - gap_extend];
        const int32_t lo_op = diag_lower[d - op_cost], up_op = diag_upper[d - op_cost];
        for (int32_t k = tmp_lower; k <= tmp_upper; k++) {
            seq2_index = GREEDY_INVALID;
            if (k + 1 <= up_goe && k + 1 >= lo_goe) seq2_index = AFF(d - goe, k + 1, 1);
            if (k + 1 <= up_ge && k + 1 >= lo_ge) {
                const int32_t v = AFF(d - gap_extend, k + 1, 2);
                if (seq2_index < v) seq2_index = v;
            }
            const int32_t del = (seq2_index == GREEDY_INVALID) ? GREEDY_INVALID : seq2_index + 1;
            AFF(d, k, 2) = del;
            seq2_index = GREEDY_INVALID;
            if (k - 1 <= up_goe && k - 1 >= lo_goe) seq2_index = AFF(d - goe, k - 1, 1);
            if (k - 1 <= up_ge && k - 1 >= lo_ge) {
                const int32_t v = AFF(d - gap_extend, k - 1, 0);
                if (seq2_index < v) seq2_index = v;
            }
            AFF(d, k, 0) = seq2_index;
            seq2_index = max(seq2_index, del);
            if (k <= up_op && k >= lo_op) seq2_index = max(seq2_index, AFF(d - op_cost, k, 1) + 1);
            seq1_index = seq2_index + k - origin;
            if (seq2_index < 0 || seq1_index + seq2_index < xdrop_score) {
                if (k == curr_lower) curr_lower++;
                else AFF(d, k, 1) = GREEDY_INVALID;
                continue;
            }
            curr_upper = k;
            index = tb_first_mismatch(sp, seq1_index, seq2_index);
            seq1_index += index; seq2_index += index;
            AFF(d, k, 1) = seq2_index;
            if (seq1_index + seq2_index > curr_extent) {
                curr_extent = seq1_index + seq2_index;
                curr_seq2_index = seq2_index;
                curr_diag = k;
            }
            if (seq1_index == len1) { curr_upper = k; end1_diag = k - 1; }
            if (seq2_index == len2) { curr_lower = k; end2_diag = k + 1; }
        }
        const int32_t curr_score = curr_extent * match_half - (int32_t)d * ac.common_factor;
        if (curr_score > max_score[d - 1]) {
            max_score[d] = curr_score;
            best_dist = (int32_t)d;
            best_diag = curr_diag;
            seq2_len = curr_seq2_index;
            seq1_len = curr_seq2_index + curr_diag - origin;
        } else max_score[d] = max_score[d - 1];
        if (curr_lower <= curr_upper) {
            num_nonempty++;
            diag_lower[d] = curr_lower; diag_upper[d] = curr_upper;
        } else { diag_lower[d] = kInvalidDiag; diag_upper[d] = -kInvalidDiag; }
        if (diag_lower[d - max_penalty] <= diag_upper[d - max_penalty]) num_nonempty--;
        if (num_nonempty == 0) break;
        d++;
        if (d >= Tn) { if (d <= scaled_max_dist) { status = 3; return 0; } break; }
        curr_lower = min(diag_lower[d - goe], diag_lower[d - gap_extend]) - 1;
        curr_lower = min(curr_lower, diag_lower[d - op_cost]);
        if (end2_diag > 0) curr_lower = max(curr_lower, end2_diag);
        curr_upper = max(diag_upper[d - goe], diag_upper[d - gap_extend]) + 1;
        curr_upper = max(curr_upper, diag_upper[d - op_cost]);
        if (end1_diag > 0) curr_upper = min(curr_upper, end1_diag);
        if (d > max_penalty && !alloc_row(d, curr_lower, curr_upper)) { status = 3; return 0; }
    }
    // ---- traceback (:1178-1232) ---------------------------------------------------------------------------------
    {
        int32_t dd = best_dist;
        seq2_index = seq2_len;
        int32_t state = 3;
        if ((int64_t)(ed.top - A) - 2 * (int64_t)(ed.n + 2 * (int64_t)dd + 4) < top) { status = 3; return 0; }
        while (dd > 0) {
            if (state == 3) {       // s_GetNextAffineTbackFromMatch
                int32_t new_seq2_index;
                bool took = false;
                if (best_diag >= diag_lower[dd - op_cost] && best_diag <= diag_upper[dd - op_cost]) {
                    new_seq2_index = AFF(dd - op_cost, best_diag, 1);
                    if (new_seq2_index >= max(AFF(dd, best_diag, 0), AFF(dd, best_diag, 2))) {
                        dd -= op_cost;
                        state = 3;
                        took = true;
                    }
                }
                if (!took) {
                    if (AFF(dd, best_diag, 0) > AFF(dd, best_diag, 2)) { new_seq2_index = AFF(dd, best_diag, 0); state = 6; }
                    else { new_seq2_index = AFF(dd, best_diag, 2); state = 0; }
                }
                ed.add(3, seq2_index - new_seq2_index);
                seq2_index = new_seq2_index;
            } else {                // s_GetNextAffineTbackFromIndel
                const int32_t IorD = state;
                ed.add(IorD, 1);
                const int32_t new_diag = (IorD == 6) ? best_diag - 1 : best_diag + 1;
                int32_t last_d = dd - gap_extend, new_seq2_index;
                if (new_diag >= diag_lower[last_d] && new_diag <= diag_upper[last_d])
                    new_seq2_index = (IorD == 6) ? AFF(last_d, new_diag, 0) : AFF(last_d, new_diag, 2);
                else new_seq2_index = GREEDY_INVALID;
                last_d = dd - goe;
                if (new_diag >= diag_lower[last_d] && new_diag <= diag_upper[last_d] &&
                    new_seq2_index < AFF(last_d, new_diag, 1)) { dd -= goe; state = 3; }
                else { dd -= gap_extend; state = IorD; }
                if (IorD == 6) best_diag--;
                else { best_diag++; seq2_index--; }
            }
        }
        ed.add(3, first_run);       // last_seq2_off[0][diag_origin].match_off
    }
#undef AFF
    (void)gap_open;
    return max_score[best_dist];
}

}  // namespace

__global__ void __launch_bounds__(32)
traceback_greedy_affine_kernel(const DevQuery q, const TracebackLaunch L)
{
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    const int64_t cap = L.arena_bytes / 4 / nt;
    int32_t *A = reinterpret_cast<int32_t *>(L.arena) + t0 * cap;
    const AffineCosts ac = affine_costs(q.reward, q.penalty, q.gap_open, q.gap_extend, L.x_dropoff);
    for (int64_t w = t0; w < L.n; w += nt) {
        if (L.todo && !L.todo[w]) continue;
        const DevTracebackItem it = L.items[w];
        const DevContext c = q.ctx[it.context];
        const int32_t q_length = c.query_length, s_length = it.s_length;
        const int32_t q_off = it.q_start, s_off = it.s_start;
        const int64_t seq_base = it.byte_off * 4 + it.s_shift;
        DevTracebackDir out, out2;
        out.score = 0; out.a_off = 0; out.b_off = 0; out.ops_off = 0; out.n_ops = 0; out.status = 0; out.ran = 1; out.pad = 0;
        out2 = out;
        int status = 0;
        int32_t q_ext_r = 0, s_ext_r = 0, q_ext_l = 0, s_ext_l = 0;
        OpList fwd, rev;
        fwd.top = A + cap; fwd.n = 0; fwd.last_op = 8;
        rev = fwd;
        GreedyTbSeq sp;
        sp.q = &q; sp.packed = L.packed;
        sp.amb = subj_amb(L.amb_runs, it.amb_first, it.amb_n, it.byte_off * 4);
        sp.qbase = c.query_offset + q_off; sp.sbase = seq_base + s_off;
        sp.len1 = q_length - q_off; sp.len2 = s_length - s_off; sp.reverse = false;
        int32_t score = greedy_align_affine_tb(sp, ac, q_ext_r, s_ext_r, A, cap, fwd, status);
        if (!status) {
            rev.top = fwd.top - 2 * fwd.n; rev.n = 0; rev.last_op = 8;
            sp.qbase = c.query_offset; sp.sbase = seq_base; sp.len1 = q_off; sp.len2 = s_off; sp.reverse = true;
            score += greedy_align_affine_tb(sp, ac, q_ext_l, s_ext_l, A, (int64_t)(rev.top - A), rev, status);
        }
        if (!status) {
            if (q.reward % 2 == 1) score /= 2;
            int32_t size = fwd.n + rev.n;
            const bool merge = fwd.n > 0 && rev.n > 0 && fwd.op(fwd.n - 1) == rev.op(rev.n - 1);
            if (merge) size--;
            if (2 * (int64_t)size + 2 > (int64_t)(rev.top - 2 * rev.n - A)) status = 3;
            else {
                int32_t *op = A, *num = A + size + 1;
                int32_t idx = 0;
                for (int32_t i = 0; i < rev.n; i++) { op[idx] = rev.op(i); num[idx] = rev.num(i); idx++; }
                if (fwd.n > 0) {
                    int32_t i = fwd.n - 1;
                    if (merge) { num[idx - 1] += fwd.num(fwd.n - 1); i = fwd.n - 2; }
                    for (; i >= 0; i--) { op[idx] = fwd.op(i); num[idx] = fwd.num(i); idx++; }
                }
                size = reduce_gaps(q, L.packed, sp.amb, c.query_offset, q_off - q_ext_l, seq_base + s_off - s_ext_l, op, num, size);
                const unsigned long long base = atomicAdd(L.ops_used, (unsigned long long)size);
                if ((long long)(base + size) > L.ops_cap) status = 4;
                else {
                    for (int32_t i = 0; i < size; i++) L.ops[base + i] = make_int2(op[i], num[i]);
                    out.ops_off = (long long)base; out.n_ops = size;
                }
                out.score = score;
                out.a_off = q_ext_l; out.b_off = s_ext_l;
                out2.a_off = q_ext_r; out2.b_off = s_ext_r;
            }
        }
        out.status = status; out2.status = status;
        L.out[2 * w] = out;
        L.out[2 * w + 1] = out2;
    }
}

cudaError_t launch_traceback_greedy_affine(const DevQuery &q, const TracebackLaunch &L, int blocks, int threads_per_block, cudaStream_t st)
{
    traceback_greedy_affine_kernel<<<blocks, threads_per_block, 0, st>>>(q, L);
    return cudaGetLastError();
}

cudaError_t launch_traceback_greedy_warp(const DevQuery &q, const TracebackLaunch &L, int blocks, cudaStream_t st)
{
    traceback_greedy_warp_kernel<<<blocks, 128, 0, st>>>(q, L);
    return cudaGetLastError();
}


// ================================================================================================
// Start point of the traceback alignment, one thread per HSP: the head of Blast_TracebackFromHSPList's loop body
// (core/blast_traceback.c:506-545): BLAST_CheckStartForGappedAlignment (:97-153), else
// BlastGetOffsetsForGappedAlignment (core/blast_gapalign.c:3059-3131); when the stored start is good,
// BlastGetStartForGappedAlignmentNucl (:3134-3182) moves it into the longest run of identities; then
// AdjustSubjectRange (:3608-3636).  items[i].pad = 1 when a start point exists (0: the reference drops the HSP).
// ================================================================================================
__global__ void traceback_start_kernel(const DevQuery q, const uint8_t *packed, const int4 *amb_runs, const DevTracebackHsp *hsps,
                                       int64_t n, DevTracebackItem *items)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int32_t HSP_MAX_WINDOW = 11, MAX_SUBJECT_OFFSET = 90000, MAX_TOTAL_GAPS = 3000;
    const DevTracebackHsp h = hsps[i];
    const DevContext c = q.ctx[h.context];
    const uint8_t *Q = q.query + c.query_offset;
    const int64_t sb = h.byte_off * 4;
    const int32_t *M = q.matrix;
    const SubjAmb amb = subj_amb(amb_runs, h.amb_first, h.amb_n, sb);
    auto sc = [&](int32_t qi, int32_t si) -> int32_t { return __ldg(M + 16 * (int)__ldg(Q + qi) + subj_code(packed, amb, sb + si)); };
    int32_t qg = h.q_gapped_start, sg = h.s_gapped_start;
    int32_t q_start = 0, s_start = 0;
    bool found = true;
    bool good = !(qg == 0 && sg == 0);
    if (good) {     // BLAST_CheckStartForGappedAlignment
        int32_t left = -HSP_MAX_WINDOW / 2;
        left = max(left, h.q_off - qg); left = max(left, h.s_off - sg);
        int32_t right = HSP_MAX_WINDOW / 2 + 1;
        right = min(right, h.q_end - qg); right = min(right, h.s_end - sg);
        int32_t score = 0;
        for (int32_t k = left; k < right; k++) score += sc(qg + k, sg + k);
        good = score > 0;
    }
    if (!good) {    // BlastGetOffsetsForGappedAlignment
        const int32_t q_length = h.q_end - h.q_off, s_length = h.s_end - h.s_off;
        if (q_length <= HSP_MAX_WINDOW) { q_start = h.q_off + q_length / 2; s_start = h.s_off + q_length / 2; }
        else {
            int32_t score = 0;
            for (int32_t k = 0; k < HSP_MAX_WINDOW; k++) score += sc(h.q_off + k, h.s_off + k);
            int32_t max_score = score, max_offset = h.q_off + HSP_MAX_WINDOW - 1;
            const int32_t hsp_end = h.q_off + min(q_length, s_length);
            for (int32_t idx = h.q_off + HSP_MAX_WINDOW; idx < hsp_end; idx++) {
                const int32_t k = idx - h.q_off;
                score -= sc(idx - HSP_MAX_WINDOW, h.s_off + k - HSP_MAX_WINDOW);
                score += sc(idx, h.s_off + k);
                if (score > max_score) { max_score = score; max_offset = idx; }
            }
            if (max_score > 0) { q_start = max_offset; s_start = (max_offset - h.q_off) + h.s_off; }
            else {
                score = 0;
                for (int32_t k = 0; k < HSP_MAX_WINDOW; k++) score += sc(h.q_end - HSP_MAX_WINDOW + k, h.s_end - HSP_MAX_WINDOW + k);
                if (score > 0) { q_start = h.q_end - HSP_MAX_WINDOW / 2; s_start = h.s_end - HSP_MAX_WINDOW / 2; }
                else found = false;
            }
        }
    } else {        // BlastGetStartForGappedAlignmentNucl
        const int32_t HSP_MAX_IDENT_RUN = 20;
        const int32_t offset = min(sg - h.s_off, qg - h.q_off);
        const int32_t q0 = qg - offset, s0 = sg - offset;
        const int32_t q_len = min(h.s_end - s0, h.q_end - q0);
        int32_t max_score = 0, max_offset = q0, score = 0, index;
        bool match = false, prev_match = false, done = false;
        for (index = q0; index < q0 + q_len; index++) {
            match = ((int)__ldg(Q + index) == subj_code(packed, amb, sb + s0 + (index - q0)));
            if (match != prev_match) {
                prev_match = match;
                if (match) score = 1;
                else if (score > max_score) { max_score = score; max_offset = index - score / 2; }
            } else if (match) {
                ++score;
                if (score > HSP_MAX_IDENT_RUN) {
                    max_offset = index - HSP_MAX_IDENT_RUN / 2;
                    qg = max_offset; sg = max_offset + s0 - q0;
                    done = true;
                    break;
                }
            }
        }
        if (!done) {
            if (match && score > max_score) { max_score = score; max_offset = index - score / 2; }
            if (max_score > 0) { qg = max_offset; sg = max_offset + s0 - q0; }
        }
        q_start = qg; s_start = sg;
    }
    // AdjustSubjectRange(&s_start, &adjusted_s_length, q_start, query_length, &start_shift)
    int32_t shift = 0, adj_len = h.seq_len;
    if (found && h.seq_len >= MAX_SUBJECT_OFFSET) {
        const int32_t s_offset = s_start;
        const int32_t max_left = q_start + MAX_TOTAL_GAPS, max_right = c.query_length - q_start + MAX_TOTAL_GAPS;
        if (s_offset > max_left) { shift = s_offset - max_left; s_start = max_left; }
        adj_len = min(h.seq_len, s_offset + max_right) - shift;
    }
    DevTracebackItem it;
    it.byte_off = h.byte_off; it.context = h.context; it.s_shift = shift; it.s_length = adj_len;
    it.q_start = q_start; it.s_start = s_start; it.pad = found ? 1 : 0;
    it.amb_first = h.amb_first; it.amb_n = h.amb_n;
    items[i] = it;
}

cudaError_t launch_traceback_start(const DevQuery &q, const uint8_t *packed, const int4 *amb_runs, const DevTracebackHsp *hsps,
                                   int64_t n, DevTracebackItem *items, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    traceback_start_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(q, packed, amb_runs, hsps, n, items);
    return cudaGetLastError();
}

// ================================================================================================
// Per-HSP sequence work after the list logic, one thread per HSP:
//   Blast_HSPReevaluateWithAmbiguitiesGapped (core/blast_hits.c:350-516) + s_UpdateReevaluatedHSP (:311-348) for the
//   HSPs the reference re-evaluates (all of them after a greedy traceback; the ones trimmed by the common-endpoint
//   pass otherwise), then the identity count of Blast_HSPGetNumIdentities (:618-700) along the edit script.
// The edit script lives in ops[esp_off ..) and is modified in place exactly like the reference modifies esp->num[];
// the surviving part is ops[esp_off + first .. esp_off + last].
// ================================================================================================
__global__ void traceback_reevaluate_kernel(const DevQuery q, const uint8_t *packed, const int4 *amb_runs,
                                            const DevTracebackPost *items, int64_t n, int2 *ops, DevTracebackPostOut *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevTracebackPost h = items[i];
    const DevContext c = q.ctx[h.context];
    const uint8_t *Q = q.query + c.query_offset;
    const int64_t sb = h.byte_off * 4;
    const int32_t qlen = c.query_length, slen = h.seq_len;
    int2 *esp = ops + h.esp_off;
    const int32_t size = h.esp_n;
    DevTracebackPostOut o;
    o.deleted = 0; o.q_off = h.q_off; o.q_end = h.q_end; o.s_off = h.s_off; o.s_end = h.s_end; o.score = h.score;
    o.first = 0; o.last = size - 1; o.num_ident = 0; o.align_length = 0;
    auto qb = [&](int32_t p) -> int { return (int)__ldg(Q + p); };
    const SubjAmb amb = subj_amb(amb_runs, h.amb_first, h.amb_n, sb);
    auto sbs = [&](int32_t p) -> int { return subj_code(packed, amb, sb + p); };
    if (h.reevaluate && size > 0) {
        int32_t factor = 1, gap_open, gap_extend;
        if (q.gap_open == 0 && q.gap_extend == 0) {
            if (q.reward % 2 == 1) factor = 2;
            gap_open = 0;
            gap_extend = (q.reward - 2 * q.penalty) * factor / 2;
        } else { gap_open = q.gap_open; gap_extend = q.gap_extend; }
        const int32_t cutoff_score = c.gapped_cutoff;
        int32_t query = h.q_off, subject = h.s_off;
        int32_t score = 0, sum = 0;
        int32_t best_q_start = query, best_q_end = query, current_q_start = query;
        int32_t best_s_start = subject, best_s_end = subject, current_s_start = subject;
        int32_t best_start_esp_index = 0, best_end_esp_index = 0, current_start_esp_index = 0, best_end_esp_num = -1;
        // matches of unambiguous bases all score `reward`: a run of them inside a substitution block can be taken in one
        // step (sum only grows, so the reference's per-base "sum > score" update ends on the run's last base)
        const int32_t match_score = __ldg(q.matrix);
        const bool uniform = __ldg(q.matrix + 17) == match_score && __ldg(q.matrix + 34) == match_score &&
                             __ldg(q.matrix + 51) == match_score && match_score > 0;
        for (int32_t index = 0; index < size; index++) {
            const int32_t op = esp[index].x;
            int32_t num = esp[index].y;
            for (int32_t op_index = 0; op_index < num;) {
                if (op == 3) {
                    int32_t run = 0;
                    if (uniform && num - op_index >= 16) {
                        uint32_t qw, qa;
                        qwin(q, c.query_offset + query, qw, qa);
                        // an ambiguous subject base ends the run: it scores through the matrix, never `reward`
                        const uint32_t m = subj_mismatch16(q, packed, amb, c.query_offset + query, sb + subject, qw, qa, false);
                        run = m ? (__clz(m) >> 1) : 16;
                    }
                    if (run > 0) { sum += run * factor * match_score; query += run; subject += run; op_index += run; }
                    else {
                        sum += factor * __ldg(q.matrix + 16 * (qb(query) & 0x0f) + sbs(subject));
                        query++; subject++; op_index++;
                    }
                } else if (op == 0) {
                    sum -= gap_open + gap_extend * num;
                    subject += num; op_index += num;
                } else if (op == 6) {
                    sum -= gap_open + gap_extend * num;
                    query += num; op_index += num;
                } else op_index++;
                if (sum < 0) {
                    if (op_index < num) {
                        num -= op_index;
                        esp[index].y = num;
                        current_start_esp_index = index;
                        op_index = 0;
                    } else current_start_esp_index = index + 1;
                    sum = 0;
                    current_q_start = query; current_s_start = subject;
                    if (score < cutoff_score) {
                        best_q_start = query; best_s_start = subject;
                        score = 0;
                        best_start_esp_index = current_start_esp_index;
                        best_end_esp_index = current_start_esp_index;
                    }
                } else if (sum > score) {
                    score = sum;
                    best_q_start = current_q_start; best_s_start = current_s_start;
                    best_q_end = query; best_s_end = subject;
                    best_start_esp_index = current_start_esp_index;
                    best_end_esp_index = index;
                    best_end_esp_num = op_index;
                }
            }
        }
        score /= factor;
        if (best_start_esp_index < size && best_end_esp_index < size) {
            int32_t qp = best_q_start, sp = best_s_start, ext = 0;
            while (qp > 0 && sp > 0) {          // while (qp > 0 && sp > 0 && q[--qp] == s[--sp] && q[qp] < 4) ext++;
                --qp; --sp;
                if (!(qb(qp) == sbs(sp) && qb(qp) < 4)) break;
                ext++;
            }
            best_q_start -= ext; best_s_start -= ext;
            esp[best_start_esp_index].y += ext;
            if (best_end_esp_index == best_start_esp_index) best_end_esp_num += ext;
            score += ext * q.reward;
            qp = best_q_end; sp = best_s_end; ext = 0;
            while (qp < qlen && sp < slen && qb(qp) < 4) {   // ... && q[qp] < 4 && (q[qp++] == s[sp++])
                const bool eq = qb(qp) == sbs(sp);
                qp++; sp++;
                if (!eq) break;
                ext++;
            }
            best_q_end += ext; best_s_end += ext;
            esp[best_end_esp_index].y += ext;
            best_end_esp_num += ext;
            score += ext * q.reward;
        }
        // s_UpdateReevaluatedHSP
        o.score = score;
        if (score >= cutoff_score) {
            o.q_off = best_q_start; o.q_end = best_q_start + (best_q_end - best_q_start);
            o.s_off = best_s_start; o.s_end = best_s_start + (best_s_end - best_s_start);
            if (best_end_esp_index != size - 1 || best_start_esp_index > 0) { o.first = best_start_esp_index; o.last = best_end_esp_index; }
            esp[o.last].y = best_end_esp_num;
        } else o.deleted = 1;
    }
    if (!o.deleted) {           // Blast_HSPGetNumIdentities along the (possibly shortened) script
        int32_t qp = o.q_off, sp = o.s_off, ident = 0, alen = 0;
        for (int32_t index = o.first; index <= o.last; index++) {
            const int32_t num = esp[index].y, op = esp[index].x;
            alen += num;
            if (op == 3) {          // 16 bases per step: identical = neither different nor ambiguous
                for (int32_t k = 0; k < num; k += 16) {
                    const int32_t nb = min(16, num - k);
                    uint32_t qw, qa;
                    qwin(q, c.query_offset + qp + k, qw, qa);
                    // byte equality of blastna codes (core/blast_hits.c:640-660): the same ambiguity code on both sides counts
                    const uint32_t m = subj_mismatch16(q, packed, amb, c.query_offset + qp + k, sb + sp + k, qw, qa, true);
                    ident += nb - __popc(m >> (32 - 2 * nb));
                }
                qp += num; sp += num;
            } else if (op == 0) sp += num;
            else if (op == 6) qp += num;
            else { sp += num; qp += num; }
        }
        o.num_ident = ident; o.align_length = alen;
    }
    out[i] = o;
}

cudaError_t launch_traceback_reevaluate(const DevQuery &q, const uint8_t *packed, const int4 *amb_runs, const DevTracebackPost *items, int64_t n,
                                        int2 *ops, DevTracebackPostOut *out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    traceback_reevaluate_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(q, packed, amb_runs, items, n, ops, out);
    return cudaGetLastError();
}

cudaError_t launch_traceback_dp(const DevQuery &q, const TracebackLaunch &L, int blocks, cudaStream_t st)
{
    if (L.wide_ring) traceback_dp_kernel<TB_WIDE_CELLS, true><<<blocks, TB_WARPS * 32, 0, st>>>(q, L);
    else traceback_dp_kernel<TB_CELLS, false><<<blocks, TB_WARPS * 32, 0, st>>>(q, L);
    return cudaGetLastError();
}
int traceback_wide_cells() { return TB_WIDE_CELLS; }
int traceback_warps_per_block() { return TB_WARPS; }

}  // namespace bn
