// hostpost.h — host-side sequential replay that follows the GPU stages.
//
// The GPU extends every init-HSP speculatively; this code replays the reference's serial
// decisions over those precomputed results so the surviving HSP lists are identical:
//   containment filter   BLAST_GetGappedScore loop        core/blast_gapalign.c:3351-3547
//   interval tree        BlastIntervalTreeAddHSP/ContainsHSP core/blast_itree.c:545-1000
//   list post-processing s_BlastSearchEngineOneContext     core/blast_engine.c:503-540
//   E-values / reap      s_BlastSearchEngineCore           core/blast_engine.c:788-806
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/gblastn_b200.h"

namespace bn {

struct HostInit {          // one init-HSP with its speculative gapped result
    int32_t chunk;         // chunk table index
    int32_t q_off, s_off, q_start, s_start, length, score;
    uint32_t order;
    // gapped result
    int32_t g_q_start, g_q_stop, g_s_start, g_s_stop, g_score, g_q_seed, g_s_seed;
    int32_t g_status = 0;  // 3: the extension was set aside (expected to be contained), there is no result
};

struct HostChunk { int32_t oid, chunk_off, len; };

// Interval tree over (query range, subject range) of saved HSPs; faithful to the reference's
// node layout and traversal, including its common-endpoint pruning on insertion.
class IntervalTree {
public:
    IntervalTree(int32_t q_min, int32_t q_max, int32_t s_min, int32_t s_max, size_t expected_items = 0);
    struct Item { int32_t q_strand_start, q_off, q_end, s_off, s_end, score; };
    bool contains(const Item &in, int32_t min_diag_separation) const;
    // check_endpoints = false skips the common-endpoint search; only valid when the tree holds no item
    // of in's query strand (the search could not match anything)
    void add(const Item &in, bool check_endpoints = true);
private:
    struct Node { int32_t leftend, rightend, leftptr, midptr, rightptr, item; };
    std::vector<Node> nodes_;
    std::vector<Item> items_;
    int32_t s_min_, s_max_;
    int32_t new_root(int32_t lo, int32_t hi);
    int32_t new_child(int32_t parent, bool left);
    int32_t new_leaf(int32_t item, int32_t q_strand_start);
    bool mid_contains(int32_t root, const Item &in, int32_t mds) const;
    bool has_endpoint(const Item &in, bool right);
    bool mid_has_endpoint(int32_t root, const Item &in, bool right);
};

struct PostParams {
    const BnQueryBatch *batch;      // host copy (contexts, cutoffs, Karlin blocks, options)
};

// Ascending sort of packed integer keys (k0, k1, k2).  The host replay's sorts are short (hundreds to a few
// thousand elements) and random, where a comparison sort spends its time in mispredicted branches: k0 is
// sorted by a byte-wise LSD radix sort over the bytes that vary, ties on k0 by comparison.
struct SortKey { uint64_t k0, k1, k2; uint32_t idx, pad; };
void sort_keys(std::vector<SortKey> &keys);

// Sort key of Blast_InitHitListSortByScore (core/blast_extend.c:274-296) + emission order.
void sort_init_hits(std::vector<HostInit> &v);
// same order for the hits of ONE chunk
void sort_chunk_init_hits(HostInit *first, HostInit *last);

// Replays BLAST_GetGappedScore for one chunk; appends saved HSPs (chunk-relative subject coords).
// What the replay reads of a context, 16 bytes instead of the 80-byte BnContext (a 100 k-read batch has
// 200 k contexts; the replay touches them at random).
struct CtxLite { int32_t query_offset, gapped_cutoff, query_index, strand_ctx; };
std::vector<CtxLite> make_ctx_lite(const BnQueryBatch &b);

// lite: optional compact context table from make_ctx_lite (built per call when NULL)
// needed (optional): when the replay reaches an init-HSP that is NOT contained and whose extension was set aside
// (g_status 3), its position in init[] is stored there and the replay of the chunk stops (what follows depends on
// that extension); without `needed` such a record is an internal error and is skipped.
void replay_gapped(const BnQueryBatch &b, const HostChunk &ch, const HostInit *init, size_t n,
                   const int32_t *low_score, std::vector<BnHSP> &out, BnStats &stats,
                   const CtxLite *lite = nullptr, int64_t *needed = nullptr);

// the same with ONE tree for all strands, exactly as the reference lays it out (kept for bn_selftest_replay)
void replay_gapped_single_tree(const BnQueryBatch &b, const HostChunk &ch, const HostInit *init, size_t n,
                               const int32_t *low_score, std::vector<BnHSP> &out, BnStats &stats);

// seeded comparison of the two formulations; number of differing cases
int64_t selftest_replay(uint64_t seed, int32_t n_cases);

// purge common endpoints + odd-score rounding + sort (core/blast_engine.c:507-513)
void finish_chunk_list(const BnQueryBatch &b, std::vector<BnHSP> &list);

// Blast_HSPListsMerge for a split subject (core/blast_hits.c:2545-2716)
void merge_chunk_lists(std::vector<BnHSP> &combined, std::vector<BnHSP> &fresh, int32_t split_offset,
                       int32_t overlap);

// E-values + reap (core/blast_engine.c:788-806)
void evalues_and_reap(const BnQueryBatch &b, std::vector<BnHSP> &list);

// The hit-list bookkeeping behind hit_params->low_score (core/blast_engine.c:1313-1320,
// Blast_HitListUpdate core/blast_hits.c:2924-2981): one instance per search.
struct HitListKey { double best_evalue; int32_t best_score; int32_t oid; };
// A query's hit list while it is still filling up is a chain through one shared arena (no per-query
// allocation: most queries of a batch never fill their list); it is copied out into `full` only when
// the first replacement needs the heap.
struct HitListState {
    int32_t count = 0;
    int32_t head = -1;              // newest arena node
    int32_t full = -1;              // index into LowScoreTracker::full_ once heapified
    bool heapified = false;
    double worst_evalue = 0.0;
    int32_t low_score = INT32_MAX;
};
class LowScoreTracker {
public:
    // track_lists: keep the hit lists even when the low_score rule is off (low_score_perc == 0), for kept_oids()
    explicit LowScoreTracker(const BnQueryBatch &b, bool track_lists = false);
    const int32_t *low_score() const { return enabled_ ? low_.data() : nullptr; }
    // true when a search over n_subjects subjects can never raise a bound: a hit list has to be full
    // (hitlist_size_ subjects) before a further subject can displace anything
    bool bounds_stay_zero(int64_t n_subjects) const { return !enabled_ || n_subjects <= (int64_t)hitlist_size_; }
    void subject_done(const BnQueryBatch &b, const std::vector<BnHSP> &list);
    int32_t hitlist_size() const { return hitlist_size_; }
    // the subjects a query's hit list holds now (what survives prelim_hitlist_size and reaches the traceback
    // stage), ascending oids
    std::vector<int32_t> kept_oids(int32_t query_index) const;
private:
    bool enabled_, track_;
    int32_t hitlist_size_;
    double perc_;
    std::vector<int32_t> low_;
    std::vector<HitListState> states_;
    std::vector<int32_t> slot_, touched_;      // scratch of subject_done
    std::vector<HitListKey> keys_;
    struct ArenaNode { HitListKey key; int32_t prev; };
    std::vector<ArenaNode> arena_;
    std::vector<std::vector<HitListKey>> full_;
};

}  // namespace bn

// ---- traceback stage: list logic around the device alignments (Blast_TracebackFromHSPList,
// core/blast_traceback.c:336-790; the per-HSP sequence work runs on the device) ---------------------------
namespace bn {

struct TbCand {                 // one preliminary HSP with its speculative traceback alignment
    BnHSP pre;                  // as the preliminary stage left it (absolute subject coordinates)
    bool has_start;             // a start point exists (else the reference drops the HSP, :514-518)
    int32_t s_shift, q_start, s_start;      // AdjustSubjectRange shift, start point (subject: window-relative)
    BnTracebackResult res;      // alignment (subject coordinates window-relative)
    const BnEditOp *ops;        // res.esp_n operations
    int32_t num_ident, align_length;    // of the alignment as it stands (DP tracebacks with an identity / length filter)
};
// Blast_HSPTest (core/blast_hits.c:864-871): true = the HSP fails percent_identity / min_hit_length
inline bool hsp_fails_identity_or_length(const BnQueryBatch &b, int32_t num_ident, int32_t align_length)
{
    return (num_ident * 100.0 < align_length * b.percent_identity) || align_length < b.min_hit_length;
}
inline bool identity_filter_on(const BnQueryBatch &b) { return b.percent_identity > 0 || b.min_hit_length > 0; }

struct TbHsp {                  // an HSP of the traceback stage
    int32_t oid, context, q_off, q_end, s_off, s_end, score, q_gapped_start, s_gapped_start;
    int32_t num_ident;
    double evalue, bit_score;
    std::vector<BnEditOp> esp;
    bool alive;                 // false = the reference's NULL entry
    bool was_cut;               // trimmed by the common-endpoint pass: re-evaluated afterwards
};

// Loop of :444-676 over one list (containment replay in score order, Blast_HSPUpdateWithTraceback,
// Blast_HSPAdjustSubjectOffset, tree insertion), then Blast_HSPListPurgeHSPsWithCommonEndpoints(purge = FALSE)
// (core/blast_hits.c:2224-2300, s_CutOffGapEditScript :2155-2221).  arr = hsp_array afterwards (dead entries
// included, in the reference's positions); extra_start = the purge's return value.
void traceback_list_stage1(const BnQueryBatch &b, int32_t subject_length, const TbCand *cand, size_t n,
                           std::vector<TbHsp> &arr, size_t &extra_start);
// :720-790 after the per-HSP re-evaluation: purge NULLs, sort by score, second containment pass,
// s_HSPListPostTracebackUpdate (:278-335: odd-score rounding, E-values, reap, bit scores).
void traceback_list_stage2(const BnQueryBatch &b, int32_t subject_length, std::vector<TbHsp> &arr);
// s_EvalueCompareHSPLists order of one query's lists (Blast_HSPResultsSortByEvalue)
bool traceback_list_before(const std::vector<TbHsp> &a, const std::vector<TbHsp> &c);

}  // namespace bn
