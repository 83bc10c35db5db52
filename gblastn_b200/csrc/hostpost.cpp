// hostpost.cpp — see hostpost.h.  Host-side, sequential, touches only the handful of HSPs per
// subject that survive the GPU stages.
#include "hostpost.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace bn {

// ------------------------------------------------------------------------------------------------
// Interval tree (core/blast_itree.c).  Node fields keep the reference's meaning: internal nodes
// cover [leftend, rightend]; a leaf (item >= 0) stores the query-strand offset in leftptr and uses
// midptr as the "next" link of a midpoint list.
// ------------------------------------------------------------------------------------------------
IntervalTree::IntervalTree(int32_t q_min, int32_t q_max, int32_t s_min, int32_t s_max, size_t expected_items)
    : s_min_(s_min), s_max_(s_max)
{
    nodes_.reserve(16 + 6 * expected_items);
    items_.reserve(expected_items);
    new_root(q_min, q_max);
}

int32_t IntervalTree::new_root(int32_t lo, int32_t hi)
{
    nodes_.push_back(Node{lo, hi, 0, 0, 0, -1});
    return (int32_t)nodes_.size() - 1;
}

int32_t IntervalTree::new_child(int32_t parent, bool left)
{
    const Node p = nodes_[parent];
    const int32_t mid = (p.leftend + p.rightend) / 2;
    Node n{0, 0, 0, 0, 0, -1};
    if (left) { n.leftend = p.leftend; n.rightend = mid; }
    else { n.leftend = mid + 1; n.rightend = p.rightend; }
    nodes_.push_back(n);
    return (int32_t)nodes_.size() - 1;
}

int32_t IntervalTree::new_leaf(int32_t item, int32_t q_strand_start)
{
    nodes_.push_back(Node{0, 0, q_strand_start, 0, 0, item});
    return (int32_t)nodes_.size() - 1;
}

// s_HSPIsContained (core/blast_itree.c:815-853)
static bool item_contained(const IntervalTree::Item &in, const IntervalTree::Item &t, int32_t tree_q_start,
                           int32_t mds)
{
    if (in.q_strand_start != tree_q_start) return false;
    if (in.score <= t.score &&
        t.q_off <= in.q_off && t.q_end >= in.q_off && t.s_off <= in.s_off && t.s_end >= in.s_off &&
        t.q_off <= in.q_end && t.q_end >= in.q_end && t.s_off <= in.s_end && t.s_end >= in.s_end) {
        if (mds == 0) return true;
        if (std::abs((t.q_off - t.s_off) - (in.q_off - in.s_off)) < mds ||
            std::abs((t.q_end - t.s_end) - (in.q_end - in.s_end)) < mds)
            return true;
    }
    return false;
}

// s_MidpointTreeContainsHSP (core/blast_itree.c:871-932)
bool IntervalTree::mid_contains(int32_t root, const Item &in, int32_t mds) const
{
    const Node *node = &nodes_[root];
    const int32_t region_start = in.s_off, region_end = in.s_end;
    while (node->item < 0) {
        for (int32_t t = node->midptr; t != 0; t = nodes_[t].midptr) {
            const Node &ln = nodes_[t];
            if (item_contained(in, items_[ln.item], ln.leftptr, mds)) return true;
        }
        int32_t next = 0;
        const int32_t middle = (node->leftend + node->rightend) / 2;
        if (region_end < middle) next = node->leftptr;
        else if (region_start > middle) next = node->rightptr;
        if (next == 0) return false;
        node = &nodes_[next];
    }
    return item_contained(in, items_[node->item], node->leftptr, mds);
}

// BlastIntervalTreeContainsHSP (core/blast_itree.c:936-1000)
bool IntervalTree::contains(const Item &in, int32_t mds) const
{
    const Node *node = &nodes_[0];
    const int32_t region_start = in.q_strand_start + in.q_off;
    const int32_t region_end = in.q_strand_start + in.q_end;
    while (node->item < 0) {
        if (node->midptr > 0 && mid_contains(node->midptr, in, mds)) return true;
        int32_t next = 0;
        const int32_t middle = (node->leftend + node->rightend) / 2;
        if (region_end < middle) next = node->leftptr;
        else if (region_start > middle) next = node->rightptr;
        if (next == 0) return false;
        node = &nodes_[next];
    }
    return item_contained(in, items_[node->item], node->leftptr, mds);
}

// s_HSPsHaveCommonEndpoint (core/blast_itree.c:251-306): 0 = no match, 1 = tree item wins, 2 = new
static int common_endpoint(const IntervalTree::Item &in, const IntervalTree::Item &t, int32_t tree_q_start,
                           bool right)
{
    if (in.q_strand_start != tree_q_start) return 0;
    const bool match = right ? (in.q_end == t.q_end && in.s_end == t.s_end)
                             : (in.q_off == t.q_off && in.s_off == t.s_off);
    if (!match) return 0;
    if (in.score > t.score) return 2;
    if (in.score < t.score) return 1;
    const int32_t iq = in.q_end - in.q_off, tq = t.q_end - t.q_off;
    if (iq > tq) return 1;
    if (iq < tq) return 2;
    const int32_t is = in.s_end - in.s_off, ts = t.s_end - t.s_off;
    if (is > ts) return 1;
    if (is < ts) return 2;
    return 1;
}

// s_MidpointTreeHasHSPEndpoint (core/blast_itree.c:324-420)
bool IntervalTree::mid_has_endpoint(int32_t root, const Item &in, bool right)
{
    int32_t root_i = root;
    const int32_t target = right ? in.s_end : in.s_off;
    for (;;) {
        // walk the midpoint list; the reference advances its "previous" pointer onto a node it
        // has just unlinked, so of two adjacent losers only the first leaves the list
        int32_t list_i = root_i;
        int32_t tmp = nodes_[root_i].midptr;
        while (tmp != 0) {
            const int32_t next_i = tmp;
            const int r = common_endpoint(in, items_[nodes_[next_i].item], nodes_[next_i].leftptr, right);
            tmp = nodes_[next_i].midptr;
            if (r == 1) return true;
            if (r == 2) nodes_[list_i].midptr = tmp;
            list_i = next_i;
        }
        int32_t next = 0;
        const int32_t midpt = (nodes_[root_i].leftend + nodes_[root_i].rightend) / 2;
        if (target < midpt) next = nodes_[root_i].leftptr;
        else if (target > midpt) next = nodes_[root_i].rightptr;
        if (next == 0) return false;
        if (nodes_[next].item >= 0) {
            const int r = common_endpoint(in, items_[nodes_[next].item], nodes_[next].leftptr, right);
            if (r == 1) return true;
            if (r == 2) {
                if (target < midpt) nodes_[root_i].leftptr = 0;
                else if (target > midpt) nodes_[root_i].rightptr = 0;
                return false;
            }
            break;
        }
        root_i = next;
    }
    return false;
}

// s_IntervalTreeHasHSPEndpoint (core/blast_itree.c:438-510)
bool IntervalTree::has_endpoint(const Item &in, bool right)
{
    int32_t root_i = 0;
    const int32_t target = in.q_strand_start + (right ? in.q_end : in.q_off);
    for (;;) {
        const int32_t mid_tree = nodes_[root_i].midptr;
        if (mid_tree != 0 && mid_has_endpoint(mid_tree, in, right)) return true;
        int32_t next = 0;
        const int32_t midpt = (nodes_[root_i].leftend + nodes_[root_i].rightend) / 2;
        if (target < midpt) next = nodes_[root_i].leftptr;
        else if (target > midpt) next = nodes_[root_i].rightptr;
        if (next == 0) return false;
        if (nodes_[next].item >= 0) {
            const int r = common_endpoint(in, items_[nodes_[next].item], nodes_[next].leftptr, right);
            if (r == 1) return true;
            if (r == 2) {
                if (target < midpt) nodes_[root_i].leftptr = 0;
                else if (target > midpt) nodes_[root_i].rightptr = 0;
                return false;
            }
            break;
        }
        root_i = next;
    }
    return false;
}

// BlastIntervalTreeAddHSP, index_method == eQueryAndSubject (core/blast_itree.c:514-800)
void IntervalTree::add(const Item &in, bool check_endpoints)
{
    if (check_endpoints) {
        if (has_endpoint(in, false)) return;
        if (has_endpoint(in, true)) return;
    }

    items_.push_back(in);
    const int32_t item = (int32_t)items_.size() - 1;
    int32_t region_start = in.q_strand_start + in.q_off;
    int32_t region_end = in.q_strand_start + in.q_end;
    bool by_subject = false;
    int32_t root = 0;
    const int32_t leaf = new_leaf(item, in.q_strand_start);

    for (;;) {
        int32_t middle = (nodes_[root].leftend + nodes_[root].rightend) / 2;
        int32_t old;
        bool left_half;
        if (region_end < middle) {
            if (nodes_[root].leftptr == 0) { nodes_[root].leftptr = leaf; return; }
            old = nodes_[root].leftptr;
            if (nodes_[old].item < 0) { root = old; continue; }
            left_half = true;
        } else if (region_start > middle) {
            if (nodes_[root].rightptr == 0) { nodes_[root].rightptr = leaf; return; }
            old = nodes_[root].rightptr;
            if (nodes_[old].item < 0) { root = old; continue; }
            left_half = false;
        } else {
            if (by_subject) {
                nodes_[leaf].midptr = nodes_[root].midptr;
                nodes_[root].midptr = leaf;
                return;
            }
            by_subject = true;
            if (nodes_[root].midptr == 0) {
                const int32_t m = new_root(s_min_, s_max_);
                nodes_[root].midptr = m;
            }
            root = nodes_[root].midptr;
            region_start = in.s_off;
            region_end = in.s_end;
            continue;
        }
        // two leaves want the same slot: interpose an internal node and re-hang the old leaf
        const int32_t mid_index = new_child(root, left_half);
        const Item &old_item = items_[nodes_[old].item];
        if (left_half) nodes_[root].leftptr = mid_index; else nodes_[root].rightptr = mid_index;
        int32_t old_start, old_end;
        if (by_subject) { old_start = old_item.s_off; old_end = old_item.s_end; }
        else { old_start = nodes_[old].leftptr + old_item.q_off; old_end = nodes_[old].leftptr + old_item.q_end; }
        root = mid_index;
        middle = (nodes_[root].leftend + nodes_[root].rightend) / 2;
        if (old_end < middle) nodes_[mid_index].leftptr = old;
        else if (old_start > middle) nodes_[mid_index].rightptr = old;
        else if (by_subject) nodes_[mid_index].midptr = old;
        else {
            const int32_t m2 = new_root(s_min_, s_max_);
            nodes_[mid_index].midptr = m2;
            const int32_t middle2 = (nodes_[m2].leftend + nodes_[m2].rightend) / 2;
            if (old_item.s_end < middle2) nodes_[m2].leftptr = old;
            else if (old_item.s_off > middle2) nodes_[m2].rightptr = old;
            else nodes_[m2].midptr = old;
        }
    }
}

// ------------------------------------------------------------------------------------------------
static int32_t ctx_search(const BnQueryBatch &b, int32_t n)
{
    int32_t lo = 0, hi = b.num_contexts;
    while (lo < hi - 1) {
        const int32_t m = (lo + hi) / 2;
        if (b.contexts[m].query_offset > n) hi = m; else lo = m;
    }
    return lo;
}

// s_GetQueryStrandOffset (core/blast_itree.c:219-234)
static int32_t strand_offset(const BnQueryBatch &b, int32_t context)
{
    int32_t c = context;
    while (c) {
        const int f = b.contexts[c].frame, pf = b.contexts[c - 1].frame;
        const int sf = (f > 0) - (f < 0), spf = (pf > 0) - (pf < 0);
        if (f == 0 || sf != spf) break;
        c--;
    }
    return b.contexts[c].query_offset;
}

void sort_init_hits(std::vector<HostInit> &v)
{
    std::sort(v.begin(), v.end(), [](const HostInit &a, const HostInit &c) {
        if (a.chunk != c.chunk) return a.chunk < c.chunk;
        if (a.score != c.score) return a.score > c.score;
        if (a.s_start != c.s_start) return a.s_start < c.s_start;
        if (a.length != c.length) return a.length > c.length;
        if (a.q_start != c.q_start) return a.q_start < c.q_start;
        return a.order < c.order;      // glibc qsort is a stable merge sort: ties keep emission order
    });
}

void sort_keys(std::vector<SortKey> &v)
{
    const size_t n = v.size();
    if (n < 2) return;
    auto less = [](const SortKey &a, const SortKey &c) {
        if (a.k0 != c.k0) return a.k0 < c.k0;
        if (a.k1 != c.k1) return a.k1 < c.k1;
        return a.k2 < c.k2;
    };
    if (n < 48) { std::sort(v.begin(), v.end(), less); return; }
    // LSD radix sort on k0, least significant byte first, skipping the bytes that are the same in every key;
    // runs of equal k0 (short in practice) are finished by comparison
    uint64_t diff = 0;
    for (size_t i = 1; i < n; i++) diff |= v[i].k0 ^ v[0].k0;
    struct P { uint64_t k; uint32_t i, pad; };
    std::vector<P> a(n), t(n);
    for (size_t i = 0; i < n; i++) a[i] = P{v[i].k0, (uint32_t)i, 0u};
    for (int byte = 0; byte < 8; byte++) {
        const int sh = 8 * byte;
        if (!((diff >> sh) & 0xFFull)) continue;
        uint32_t count[257] = {0};
        for (size_t i = 0; i < n; i++) ++count[((a[i].k >> sh) & 0xFFull) + 1];
        for (int c = 0; c < 256; c++) count[c + 1] += count[c];
        for (size_t i = 0; i < n; i++) t[count[(a[i].k >> sh) & 0xFFull]++] = a[i];
        a.swap(t);
    }
    std::vector<SortKey> out(n);
    for (size_t i = 0; i < n; i++) out[i] = v[a[i].i];
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && out[j].k0 == out[i].k0) ++j;
        if (j - i > 1) std::sort(out.begin() + (ptrdiff_t)i, out.begin() + (ptrdiff_t)j, less);
        i = j;
    }
    v.swap(out);
}

void sort_chunk_init_hits(HostInit *first, HostInit *last)
{
    std::sort(first, last, [](const HostInit &a, const HostInit &c) {
        if (a.score != c.score) return a.score > c.score;
        if (a.s_start != c.s_start) return a.s_start < c.s_start;
        if (a.length != c.length) return a.length > c.length;
        if (a.q_start != c.q_start) return a.q_start < c.q_start;
        return a.order < c.order;
    });
}

void replay_gapped_single_tree(const BnQueryBatch &b, const HostChunk &ch, const HostInit *init, size_t n,
                   const int32_t *low_score, std::vector<BnHSP> &out, BnStats &stats)
{
    if (n == 0) return;
    IntervalTree tree(0, b.concat_len + 1, 0, ch.len + 1, n);
    // contexts of every init-HSP (BSearchContextInfo on the seed's query offset)
    std::vector<int32_t> ctx_of(n);
    for (size_t i = 0; i < n; i++) ctx_of[i] = ctx_search(b, init[i].q_off);
    std::vector<uint8_t> found_high;
    if (low_score) {
        found_high.assign((size_t)b.num_queries, 0);
        for (size_t i = 0; i < n; i++) {
            const int32_t qi = b.contexts[ctx_of[i]].query_index;
            if (init[i].score > low_score[qi]) found_high[qi] = 1;
        }
    }
    // Containment and common-endpoint tests can only succeed against an HSP of the same query
    // strand (s_HSPIsContained / s_HSPsHaveCommonEndpoint compare the strand offsets first), so they
    // are skipped - not approximated - while the tree holds nothing for that strand yet.  Insertion
    // itself always runs: it shapes the tree that later tests walk.
    std::vector<int32_t> strand_items((size_t)b.num_contexts, 0);
    for (size_t i = 0; i < n; i++) {
        const HostInit &h = init[i];
        const int32_t context = ctx_of[i];
        const BnContext &c = b.contexts[context];
        if (low_score && !found_high[c.query_index]) continue;
        IntervalTree::Item t;
        t.q_strand_start = strand_offset(b, context);
        t.q_off = h.q_start - c.query_offset;
        t.q_end = t.q_off + h.length;
        t.s_off = h.s_start;
        t.s_end = h.s_start + h.length;
        t.score = h.score;
        const bool strand_seen = strand_items[context] != 0;
        if (strand_seen && tree.contains(t, b.min_diag_separation)) continue;
        ++stats.gap_extensions;
        if (h.g_score >= c.gapped_cutoff) {
            BnHSP o;
            o.oid = ch.oid; o.context = context; o.chunk_off = ch.chunk_off;
            o.q_off = h.g_q_start; o.q_end = h.g_q_stop; o.s_off = h.g_s_start; o.s_end = h.g_s_stop;
            o.score = h.g_score; o.q_gapped_start = h.g_q_seed; o.s_gapped_start = h.g_s_seed;
            o.evalue = 0.0;
            out.push_back(o);
            IntervalTree::Item nt{t.q_strand_start, o.q_off, o.q_end, o.s_off, o.s_end, o.score};
            tree.add(nt, strand_seen);
            ++strand_items[context];
        }
    }
}

// s_GetQueryStrandOffset as a context index: first context of the run of same-sign frames
static int32_t strand_context(const BnQueryBatch &b, int32_t context)
{
    int32_t c = context;
    while (c) {
        const int f = b.contexts[c].frame, pf = b.contexts[c - 1].frame;
        const int sf = (f > 0) - (f < 0), spf = (pf > 0) - (pf < 0);
        if (f == 0 || sf != spf) break;
        c--;
    }
    return c;
}

// BLAST_GetGappedScore's containment filter over precomputed extensions, one interval tree PER QUERY STRAND.
//
// The reference keeps one tree per subject chunk for all queries.  What a test or an insertion does for an
// HSP of strand S depends only on the HSPs of S inserted before it, in their order:
//  * items of different strands occupy disjoint ranges of the concatenated query, and the tree is a FIXED
//    binary subdivision of [0, Qcat] (nodes are only materialised lazily): an item's home is the first node
//    on its path whose centre it contains, so two strands never share a node's subject tree or midpoint list;
//  * s_HSPIsContained / s_HSPsHaveCommonEndpoint compare the strand offsets first, so items of other strands
//    met on a walk (in an ancestor's lists, or as a not-yet-displaced leaf that ends the walk) change nothing,
//    and a leaf of another strand can only sit where no item of S lies below it;
//  * a not-yet-displaced leaf is the deepest materialised point of its path, so the order in which the items
//    of S are met (which is what the reference's unlink quirk in s_MidpointTreeHasHSPEndpoint depends on) is
//    the same whether or not other strands pushed it down earlier; list order is reverse insertion order.
// Each strand therefore gets its own tree WITH THE SAME ROOT RANGES (same node centres), created only when
// its second HSP arrives: most strands of a batch have a single init-HSP in a chunk and need no tree at all.
// replay_gapped_single_tree above is the one-tree formulation; bn_selftest_replay compares the two.
std::vector<CtxLite> make_ctx_lite(const BnQueryBatch &b)
{
    std::vector<CtxLite> v((size_t)b.num_contexts);
    for (int32_t c = 0; c < b.num_contexts; c++)
        v[(size_t)c] = CtxLite{b.contexts[c].query_offset, b.contexts[c].gapped_cutoff, b.contexts[c].query_index,
                               strand_context(b, c)};
    return v;
}

// BSearchContextInfo (core/blast_query_info.c:220-236) over the compact table
static int32_t ctx_search_lite(const CtxLite *L, int32_t n_ctx, int32_t n)
{
    // same answer as the reference's loop (the last context whose offset is <= n, context 0 below the first
    // offset), written without data-dependent branches: the probes of a replay are random, and a mispredicted
    // branch per level cost more than the rest of the per-HSP work
    const CtxLite *base = L;
    int32_t len = n_ctx;
    while (len > 1) {
        const int32_t half = len >> 1;
        base = (base[half].query_offset <= n) ? base + half : base;
        len -= half;
    }
    return (int32_t)(base - L);
}

void replay_gapped(const BnQueryBatch &b, const HostChunk &ch, const HostInit *init, size_t n,
                   const int32_t *low_score, std::vector<BnHSP> &out, BnStats &stats, const CtxLite *lite, int64_t *needed)
{
    if (needed) *needed = -1;
    if (n == 0) return;
    std::vector<CtxLite> local;
    if (!lite) { local = make_ctx_lite(b); lite = local.data(); }
    std::vector<int32_t> ctx_of(n);
    for (size_t i = 0; i < n; i++) ctx_of[i] = ctx_search_lite(lite, b.num_contexts, init[i].q_off);
    std::vector<uint8_t> found_high;
    if (low_score) {
        found_high.assign((size_t)b.num_queries, 0);
        for (size_t i = 0; i < n; i++) {
            const int32_t qi = lite[ctx_of[i]].query_index;
            if (init[i].score > low_score[qi]) found_high[qi] = 1;
        }
    }
    // per strand: -1 nothing saved yet; -2 - k: one saved HSP, parked in first_item[k]; >= 0: index of its tree
    std::vector<int32_t> state((size_t)b.num_contexts, -1);
    std::vector<IntervalTree::Item> first_item;
    std::vector<IntervalTree> trees;
    first_item.reserve(n);
    out.reserve(out.size() + n);
    for (size_t i = 0; i < n; i++) {
        const HostInit &h = init[i];
        const int32_t context = ctx_of[i];
        const CtxLite &c = lite[context];
        if (low_score && !found_high[c.query_index]) continue;
        const int32_t sc = c.strand_ctx;
        IntervalTree::Item t;
        t.q_strand_start = lite[sc].query_offset;
        t.q_off = h.q_start - c.query_offset;
        t.q_end = t.q_off + h.length;
        t.s_off = h.s_start;
        t.s_end = h.s_start + h.length;
        t.score = h.score;
        int32_t &st = state[(size_t)sc];
        if (st <= -2) {                       // second HSP of the strand: now the tree is needed
            trees.emplace_back(0, b.concat_len + 1, 0, ch.len + 1, 4);
            trees.back().add(first_item[(size_t)(-2 - st)], false);
            st = (int32_t)trees.size() - 1;
        }
        if (st >= 0 && trees[(size_t)st].contains(t, b.min_diag_separation)) continue;
        if (h.g_status == 3) {               // the reference extends here, and the extension was set aside
            if (needed) { *needed = (int64_t)i; return; }
            continue;
        }
        ++stats.gap_extensions;
        if (h.g_score >= c.gapped_cutoff) {
            BnHSP o;
            o.oid = ch.oid; o.context = context; o.chunk_off = ch.chunk_off;
            o.q_off = h.g_q_start; o.q_end = h.g_q_stop; o.s_off = h.g_s_start; o.s_end = h.g_s_stop;
            o.score = h.g_score; o.q_gapped_start = h.g_q_seed; o.s_gapped_start = h.g_s_seed;
            o.evalue = 0.0;
            out.push_back(o);
            const IntervalTree::Item nt{t.q_strand_start, o.q_off, o.q_end, o.s_off, o.s_end, o.score};
            if (st == -1) { first_item.push_back(nt); st = -2 - ((int32_t)first_item.size() - 1); }
            else trees[(size_t)st].add(nt, true);
        }
    }
}

// true when two HSPs of the list share (context, query start, subject start) or (context, query
// end, subject end) - the only situation in which the purge pass changes anything
static bool has_common_endpoints(const std::vector<BnHSP> &list)
{
    const size_t n = list.size();
    if (n < 2) return false;
    size_t cap = 64;
    while (cap < 4 * n) cap <<= 1;
    std::vector<uint64_t> tab(2 * cap, ~0ull);
    auto probe = [&](uint64_t *t, uint64_t a, uint64_t c) -> bool {
        // 96-bit key folded into a 64-bit value + verified on collision by storing both halves
        uint64_t h = (a * 0x9E3779B97F4A7C15ull) ^ (c * 0xC2B2AE3D27D4EB4Full);
        size_t i = (size_t)(h >> 20) & (cap - 1);
        for (;;) {
            if (t[i] == ~0ull) { t[i] = h; return false; }
            if (t[i] == h) return true;      // equal 64-bit hash: treat as a potential collision
            i = (i + 1) & (cap - 1);
        }
    };
    for (const BnHSP &h : list) {
        const uint64_t c = (uint64_t)(uint32_t)h.context;
        if (probe(tab.data(), ((uint64_t)(uint32_t)h.q_off << 32) | (uint32_t)h.s_off, c)) return true;
        if (probe(tab.data() + cap, ((uint64_t)(uint32_t)h.q_end << 32) | (uint32_t)h.s_end, c)) return true;
    }
    return false;
}

void finish_chunk_list(const BnQueryBatch &b, std::vector<BnHSP> &list)
{
    if (!has_common_endpoints(list)) {
        // Nothing to purge.  The reference still sorts by query offset, then by query end, then
        // (after the odd-score rounding) by score with a stable sort; two HSPs that tie on the score
        // comparator can only differ in context, and the query-end order puts the lower context
        // first, so a single sort with context as the last key gives the identical list.
        if (b.round_down)
            for (auto &h : list) h.score &= ~1;
        // the same order through packed integer keys (scores, offsets and contexts are non-negative here):
        // comparing three words and moving 32 bytes beats comparing six fields and moving 64
        const size_t n = list.size();
        std::vector<SortKey> keys(n);
        for (size_t i = 0; i < n; i++) {
            const BnHSP &h = list[i];
            keys[i] = SortKey{((uint64_t)(uint32_t)(INT32_MAX - h.score) << 32) | (uint32_t)h.s_off,
                              ((uint64_t)(uint32_t)(INT32_MAX - h.s_end) << 32) | (uint32_t)h.q_off,
                              ((uint64_t)(uint32_t)(INT32_MAX - h.q_end) << 32) | (uint32_t)h.context, (uint32_t)i, 0u};
        }
        sort_keys(keys);
        std::vector<BnHSP> sorted(n);
        for (size_t i = 0; i < n; i++) sorted[i] = list[keys[i].idx];
        list.swap(sorted);
        return;
    }
    // Blast_HSPListPurgeHSPsWithCommonEndpoints(purge = TRUE), core/blast_hits.c:2224-2300
    std::stable_sort(list.begin(), list.end(), [](const BnHSP &x, const BnHSP &y) {
        if (x.context != y.context) return x.context < y.context;
        if (x.q_off != y.q_off) return x.q_off < y.q_off;
        if (x.s_off != y.s_off) return x.s_off < y.s_off;
        if (x.score != y.score) return x.score > y.score;
        if (x.q_end != y.q_end) return x.q_end > y.q_end;
        if (x.s_end != y.s_end) return x.s_end > y.s_end;
        return false;
    });
    size_t o = 0;
    for (size_t i = 0; i < list.size(); i++) {
        if (o > 0 && list[o - 1].context == list[i].context && list[o - 1].q_off == list[i].q_off &&
            list[o - 1].s_off == list[i].s_off) continue;
        list[o++] = list[i];
    }
    list.resize(o);
    std::stable_sort(list.begin(), list.end(), [](const BnHSP &x, const BnHSP &y) {
        if (x.context != y.context) return x.context < y.context;
        if (x.q_end != y.q_end) return x.q_end < y.q_end;
        if (x.s_end != y.s_end) return x.s_end < y.s_end;
        if (x.score != y.score) return x.score > y.score;
        if (x.q_off != y.q_off) return x.q_off > y.q_off;
        if (x.s_off != y.s_off) return x.s_off > y.s_off;
        return false;
    });
    o = 0;
    for (size_t i = 0; i < list.size(); i++) {
        if (o > 0 && list[o - 1].context == list[i].context && list[o - 1].q_end == list[i].q_end &&
            list[o - 1].s_end == list[i].s_end) continue;
        list[o++] = list[i];
    }
    list.resize(o);
    // Blast_HSPListAdjustOddBlastnScores, core/blast_hits.c:2734-2749
    if (b.round_down)
        for (auto &h : list) h.score &= ~1;
    // Blast_HSPListSortByScore (ScoreCompareHSPs, core/blast_hits.c:1182-1210)
    std::stable_sort(list.begin(), list.end(), [](const BnHSP &x, const BnHSP &y) {
        if (x.score != y.score) return x.score > y.score;
        if (x.s_off != y.s_off) return x.s_off < y.s_off;
        if (x.s_end != y.s_end) return x.s_end > y.s_end;
        if (x.q_off != y.q_off) return x.q_off < y.q_off;
        if (x.q_end != y.q_end) return x.q_end > y.q_end;
        return false;
    });
}

static bool score_less(const BnHSP &x, const BnHSP &y)
{
    if (x.score != y.score) return x.score > y.score;
    if (x.s_off != y.s_off) return x.s_off < y.s_off;
    if (x.s_end != y.s_end) return x.s_end > y.s_end;
    if (x.q_off != y.q_off) return x.q_off < y.q_off;
    if (x.q_end != y.q_end) return x.q_end > y.q_end;
    return false;
}

void merge_chunk_lists(std::vector<BnHSP> &comb, std::vector<BnHSP> &fresh, int32_t split_offset,
                       int32_t overlap)
{
    if (fresh.empty()) return;
    if (comb.empty()) { comb.swap(fresh); return; }
    size_t n1 = 0, n2 = 0;
    for (size_t i = 0; i < comb.size(); i++)
        if (comb[i].s_end > split_offset) { std::swap(comb[n1], comb[i]); ++n1; }
    for (size_t i = 0; i < fresh.size(); i++)
        if (fresh[i].s_off < split_offset + overlap) { std::swap(fresh[n2], fresh[i]); ++n2; }
    if (n1 > 0 && n2 > 0) {
        std::vector<uint8_t> dead(fresh.size(), 0);
        for (size_t i = 0; i < n1; i++) {
            BnHSP &h1 = comb[i];
            for (size_t j = 0; j < n2; j++) {
                BnHSP &h2 = fresh[j];
                if (dead[j] || h1.context != h2.context) continue;
                // OVERLAP_DIAG_CLOSE 10; s_BlastMergeTwoHSPs core/blast_hits.c:1337-1375
                if (std::abs((h1.q_end - h1.s_end) - (h2.q_off - h2.s_off)) >= 10) continue;
                const bool c1 = h1.q_off <= h2.q_off && h1.q_end >= h2.q_off &&
                                h1.s_off <= h2.s_off && h1.s_end >= h2.s_off;
                const bool c2 = h1.q_off <= h2.q_end && h1.q_end >= h2.q_end &&
                                h1.s_off <= h2.s_end && h1.s_end >= h2.s_end;
                if (!(c1 || c2)) continue;
                h1.q_off = std::min(h1.q_off, h2.q_off); h1.s_off = std::min(h1.s_off, h2.s_off);
                h1.q_end = std::max(h1.q_end, h2.q_end); h1.s_end = std::max(h1.s_end, h2.s_end);
                if (h2.score > h1.score) {
                    h1.q_gapped_start = h2.q_gapped_start;
                    h1.s_gapped_start = h2.s_gapped_start;
                    h1.score = h2.score;
                }
                dead[j] = 1;
            }
        }
        size_t o = 0;
        for (size_t j = 0; j < fresh.size(); j++) if (!dead[j]) fresh[o++] = fresh[j];
        fresh.resize(o);
    }
    comb.insert(comb.end(), fresh.begin(), fresh.end());
    fresh.clear();
    std::stable_sort(comb.begin(), comb.end(), score_less);
}

void evalues_and_reap(const BnQueryBatch &b, std::vector<BnHSP> &list)
{
    size_t o = 0;
    for (size_t i = 0; i < list.size(); i++) {
        BnHSP h = list[i];
        const BnContext &c = b.contexts[h.context];
        // BLAST_KarlinStoE_simple, core/blast_stat.c:4111-4125
        h.evalue = (double)c.eff_searchsp * std::exp((double)(-c.gap_lambda * h.score) + c.gap_logK);
        if (h.evalue > b.evalue_cutoff) continue;
        list[o++] = h;
    }
    list.resize(o);
}

// ------------------------------------------------------------------------------------------------
// Hit-list model for the low_score feedback.
// ------------------------------------------------------------------------------------------------
namespace {
typedef HitListKey ListKey;

int fuzzy_cmp(double e1, double e2)     // s_FuzzyEvalueComp, FUZZY_EVALUE_COMPARE_FACTOR 1e-6
{
    if (e1 < (1 - 1e-6) * e2) return -1;
    if (e1 > (1 + 1e-6) * e2) return 1;
    return 0;
}
int list_cmp(const ListKey &a, const ListKey &c)   // s_EvalueCompareHSPLists
{
    int r = fuzzy_cmp(a.best_evalue, c.best_evalue);
    if (r) return r;
    if (a.best_score > c.best_score) return -1;
    if (a.best_score < c.best_score) return 1;
    return (c.oid > a.oid) ? 1 : (c.oid < a.oid ? -1 : 0);
}
// s_Heapify / s_CreateHeap (core/blast_hits.c:1470-1521): worst list at the root
void sift(std::vector<ListKey> &h, size_t base, size_t lim, size_t last)
{
    size_t left = 2 * base + 1;
    while (base <= lim) {
        size_t large;
        if (left == last) large = left;
        else large = list_cmp(h[left], h[left + 1]) >= 0 ? left : left + 1;
        if (list_cmp(h[base], h[large]) < 0) {
            std::swap(h[base], h[large]);
            base = large;
            left = 2 * base + 1;
        } else break;
    }
}
void make_heap(std::vector<ListKey> &h)
{
    const size_t n = h.size();
    if (n < 2) return;
    const size_t lim = (n - 2) / 2, last = n - 1;
    for (size_t i = n / 2; i > 0; i--) sift(h, i - 1, lim, last);
}
}  // namespace

LowScoreTracker::LowScoreTracker(const BnQueryBatch &b, bool track_lists)
{
    enabled_ = b.low_score_perc > 0.00001;
    track_ = track_lists;
    perc_ = b.low_score_perc;
    // BlastHSPCollectorParamsNew, core/hspfilter_collector.c:335-342 (gapped search)
    int32_t hs = b.hitlist_size > 0 ? b.hitlist_size : 500;
    hs = std::min(2 * hs, hs + 50);
    hitlist_size_ = std::max(hs, 10);
    low_.assign((size_t)std::max(b.num_queries, 1), 0);
    states_.resize((size_t)std::max(b.num_queries, 1));
}

void LowScoreTracker::subject_done(const BnQueryBatch &b, const std::vector<BnHSP> &list)
{
    if (!(enabled_ || track_) || list.empty()) return;
    // the collector splits the subject's list per query (core/hspfilter_collector.c:104-150);
    // `list` is sorted by score, so the first HSP of a query is its hsp_array[0]
    touched_.clear();
    keys_.clear();
    if (slot_.size() != (size_t)b.num_queries) slot_.assign((size_t)b.num_queries, -1);
    for (const BnHSP &h : list) {
        const int32_t qi = b.contexts[h.context].query_index;
        if (slot_[qi] < 0) {
            slot_[qi] = (int32_t)keys_.size();
            keys_.push_back(HitListKey{h.evalue, h.score, h.oid});
            touched_.push_back(qi);
        } else {
            HitListKey &k = keys_[slot_[qi]];
            k.best_evalue = std::min(k.best_evalue, h.evalue);
        }
    }
    std::sort(touched_.begin(), touched_.end());
    for (int32_t qi : touched_) {
        HitListState &S = states_[qi];
        const HitListKey &k = keys_[slot_[qi]];
        if (S.count < hitlist_size_) {
            arena_.push_back(ArenaNode{k, S.head});
            S.head = (int32_t)arena_.size() - 1;
            ++S.count;
            S.worst_evalue = std::max(k.best_evalue, S.worst_evalue);
            S.low_score = std::min(k.best_score, S.low_score);
        } else {
            const int order = fuzzy_cmp(k.best_evalue, S.worst_evalue);
            if (!(order > 0 || (order == 0 && k.best_score < S.low_score))) {
                if (!S.heapified) {
                    // copy the chain out in insertion order, then s_CreateHeap
                    full_.emplace_back((size_t)S.count);
                    S.full = (int32_t)full_.size() - 1;
                    std::vector<HitListKey> &L = full_.back();
                    int32_t at = S.head;
                    for (int32_t i = S.count - 1; i >= 0 && at >= 0; i--) { L[(size_t)i] = arena_[(size_t)at].key; at = arena_[(size_t)at].prev; }
                    make_heap(L);
                    S.heapified = true;
                }
                std::vector<HitListKey> &L = full_[(size_t)S.full];
                L[0] = k;
                if (L.size() >= 2) sift(L, 0, L.size() / 2 - 1, L.size() - 1);
                S.worst_evalue = L[0].best_evalue;
                S.low_score = L[0].best_score;
            }
        }
        // core/blast_engine.c:1313-1320 (only a query whose hit list changed can change its bound)
        if (S.heapified && enabled_) {
            const double v = perc_ * (double)S.low_score;
            if ((double)low_[qi] < v) low_[qi] = (int32_t)v;
        }
        slot_[qi] = -1;
    }
}

std::vector<int32_t> LowScoreTracker::kept_oids(int32_t qi) const
{
    std::vector<int32_t> out;
    if (qi < 0 || (size_t)qi >= states_.size()) return out;
    const HitListState &S = states_[(size_t)qi];
    if (S.heapified) for (const HitListKey &k : full_[(size_t)S.full]) out.push_back(k.oid);
    else for (int32_t at = S.head; at >= 0; at = arena_[(size_t)at].prev) out.push_back(arena_[(size_t)at].key.oid);
    std::sort(out.begin(), out.end());
    return out;
}

// ------------------------------------------------------------------------------------------------
// Self-test of the per-strand formulation against the one-tree formulation on seeded random inputs that
// are rich in what makes the tree interesting: several HSPs per strand, nested boxes, shared start / end
// points, equal scores.  Returns the number of cases whose outputs differ.
int64_t selftest_replay(uint64_t seed, int32_t n_cases)
{
    auto rnd = [&seed]() {                     // splitmix64
        uint64_t z = (seed += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    auto uni = [&](int64_t lo, int64_t hi) { return lo + (int64_t)(rnd() % (uint64_t)(hi - lo + 1)); };
    int64_t bad = 0;
    for (int32_t cs = 0; cs < n_cases; cs++) {
        const int nq = (int)uni(1, 6);
        std::vector<BnContext> ctx((size_t)(2 * nq));
        int32_t off = 0;
        for (int c = 0; c < 2 * nq; c++) {
            BnContext &x = ctx[(size_t)c];
            memset(&x, 0, sizeof x);
            x.query_length = (c % 2) ? ctx[(size_t)c - 1].query_length : (int32_t)uni(150, 2500);
            x.query_offset = off; x.query_index = c / 2; x.frame = (c % 2) ? -1 : 1; x.is_valid = 1;
            x.gapped_cutoff = (int32_t)uni(20, 60);
            off += x.query_length + 1;
        }
        BnQueryBatch b;
        memset(&b, 0, sizeof b);
        b.contexts = ctx.data(); b.num_contexts = 2 * nq; b.num_queries = nq; b.concat_len = off - 1;
        const int32_t mds_choices[3] = {0, 6, 50};
        b.min_diag_separation = mds_choices[uni(0, 2)];
        HostChunk ch{0, 0, (int32_t)uni(5000, 200000)};
        // a few "true alignments" per strand; init-HSPs sit on them and their gapped boxes snap to their ends
        struct Aln { int32_t ctx, q0, q1, s0, s1; };
        std::vector<Aln> alns;
        const int n_aln = (int)uni(1, 8);
        for (int a = 0; a < n_aln; a++) {
            Aln A;
            A.ctx = (int32_t)uni(0, 2 * nq - 1);
            const int32_t L = ctx[(size_t)A.ctx].query_length;
            A.q0 = (int32_t)uni(0, L - 60); A.q1 = (int32_t)uni(A.q0 + 40, L);
            A.s0 = (int32_t)uni(0, ch.len - (A.q1 - A.q0) - 10); A.s1 = A.s0 + (A.q1 - A.q0) + (int32_t)uni(-3, 3);
            alns.push_back(A);
        }
        const int n = (int)uni(1, 160);
        std::vector<HostInit> inits((size_t)n);
        for (int i = 0; i < n; i++) {
            const Aln &A = alns[(size_t)uni(0, n_aln - 1)];
            const BnContext &c = ctx[(size_t)A.ctx];
            HostInit h;
            memset(&h, 0, sizeof h);
            const int32_t len = (int32_t)uni(11, std::max<int32_t>(12, (A.q1 - A.q0) / 2));
            const int32_t qs = (int32_t)uni(A.q0, std::max(A.q0, A.q1 - len));
            h.chunk = 0; h.length = len; h.q_start = c.query_offset + qs;
            h.s_start = std::max<int32_t>(0, A.s0 + (qs - A.q0) + (int32_t)uni(-2, 2));
            h.q_off = h.q_start + (int32_t)uni(0, len - 1); h.s_off = h.s_start + (h.q_off - h.q_start);
            h.score = (int32_t)uni(15, 120); h.order = (uint32_t)i;
            // gapped box: usually the alignment's ends (shared endpoints), sometimes a private variation
            const int kind = (int)uni(0, 9);
            h.g_q_start = kind < 7 ? A.q0 : std::max(0, qs - (int32_t)uni(0, 30));
            h.g_s_start = kind < 7 ? A.s0 : std::max(0, h.s_start - (int32_t)uni(0, 30));
            h.g_q_stop = (kind < 6 || kind == 8) ? A.q1 : std::min(c.query_length, qs + len + (int32_t)uni(0, 30));
            h.g_s_stop = (kind < 6 || kind == 8) ? A.s1 : std::min(ch.len, h.s_start + len + (int32_t)uni(0, 30));
            h.g_score = (int32_t)uni(10, 200) & (kind == 9 ? ~0 : ~3);      // many equal scores
            h.g_q_seed = h.g_q_start; h.g_s_seed = h.g_s_start;
            inits[(size_t)i] = h;
        }
        sort_chunk_init_hits(inits.data(), inits.data() + inits.size());
        std::vector<BnHSP> o1, o2;
        BnStats s1, s2;
        memset(&s1, 0, sizeof s1); memset(&s2, 0, sizeof s2);
        replay_gapped_single_tree(b, ch, inits.data(), inits.size(), nullptr, o1, s1);
        replay_gapped(b, ch, inits.data(), inits.size(), nullptr, o2, s2);
        const bool same = o1.size() == o2.size() && s1.gap_extensions == s2.gap_extensions &&
                          (o1.empty() || memcmp(o1.data(), o2.data(), o1.size() * sizeof(BnHSP)) == 0);
        if (!same) ++bad;
    }
    return bad;
}

}  // namespace bn

// ================================================================================================
// Traceback stage: list logic (see hostpost.h)
// ================================================================================================
namespace bn {

static bool tb_score_less(const TbHsp &x, const TbHsp &y)      // ScoreCompareHSPs; NULLs last
{
    if (!x.alive || !y.alive) return x.alive && !y.alive;
    if (x.score != y.score) return x.score > y.score;
    if (x.s_off != y.s_off) return x.s_off < y.s_off;
    if (x.s_end != y.s_end) return x.s_end > y.s_end;
    if (x.q_off != y.q_off) return x.q_off < y.q_off;
    if (x.q_end != y.q_end) return x.q_end > y.q_end;
    return false;
}

// s_CutOffGapEditScript (core/blast_hits.c:2155-2221)
static void cut_off_edit_script(TbHsp &hsp, int32_t q_cut, int32_t s_cut, bool cut_begin)
{
    int32_t index, opid = 0, qid = 0, sid = 0;
    bool found = false;
    std::vector<BnEditOp> &esp = hsp.esp;
    const int32_t size = (int32_t)esp.size();
    q_cut -= hsp.q_off;
    s_cut -= hsp.s_off;
    for (index = 0; index < size; index++) {
        for (opid = 0; opid < esp[(size_t)index].num;) {
            if (esp[(size_t)index].op_type == 3) { qid++; sid++; opid++; }
            else if (esp[(size_t)index].op_type == 0) { sid += esp[(size_t)index].num; opid += esp[(size_t)index].num; }
            else if (esp[(size_t)index].op_type == 6) { qid += esp[(size_t)index].num; opid += esp[(size_t)index].num; }
            else opid++;
            if (qid >= q_cut && sid >= s_cut) found = true;
            if (found) break;
        }
        if (found) break;
    }
    if (!found) return;
    if (cut_begin) {
        int32_t new_index = 0;
        if (opid < esp[(size_t)index].num) {
            esp[0].op_type = esp[(size_t)index].op_type;
            esp[0].num = esp[(size_t)index].num - opid;
            new_index++;
        }
        ++index;
        for (; index < size; index++, new_index++) esp[(size_t)new_index] = esp[(size_t)index];
        esp.resize((size_t)new_index);
        hsp.q_off += qid;
        hsp.s_off += sid;
    } else {
        if (opid < esp[(size_t)index].num) esp[(size_t)index].num = opid;
        esp.resize((size_t)index + 1);
        hsp.q_end = hsp.q_off + qid;
        hsp.s_end = hsp.s_off + sid;
    }
}

void traceback_list_stage1(const BnQueryBatch &b, int32_t subject_length, const TbCand *cand, size_t n,
                           std::vector<TbHsp> &arr, size_t &extra_start)
{
    arr.clear();
    extra_start = 0;
    if (n == 0) return;
    // DP tracebacks are tested for identity / length right after the alignment and a failing HSP never enters the tree
    // (core/blast_traceback.c:658-676); greedy ones are tested after the re-evaluation (:727-735, the caller)
    const bool dp_filter = b.gap_algo != BN_GAP_GREEDY && identity_filter_on(b);
    if (n == 1) {       // a list of one HSP (the usual case for short reads): nothing to be contained in, nothing to purge
        const TbCand &c = cand[0];
        if (!c.has_start) return;
        if (dp_filter && hsp_fails_identity_or_length(b, c.num_ident, c.align_length)) return;
        TbHsp h{};
        h.alive = true; h.was_cut = false;
        h.oid = c.pre.oid; h.context = c.pre.context;
        h.score = c.res.score;
        h.q_off = c.res.query_start; h.q_end = c.res.query_stop;
        h.s_off = c.res.subject_start + c.s_shift; h.s_end = c.res.subject_stop + c.s_shift;
        h.q_gapped_start = c.q_start; h.s_gapped_start = c.s_start + c.s_shift;
        h.esp.assign(c.ops, c.ops + c.res.esp_n);
        arr.push_back(std::move(h));
        extra_start = 1;
        return;
    }
    IntervalTree tree(0, b.concat_len + 1, 0, subject_length + 1, n);
    for (size_t i = 0; i < n; i++) {
        const TbCand &c = cand[i];
        TbHsp h{};
        h.alive = false; h.was_cut = false;
        h.oid = c.pre.oid; h.context = c.pre.context;
        IntervalTree::Item t;
        t.q_strand_start = strand_offset(b, c.pre.context);
        t.q_off = c.pre.q_off; t.q_end = c.pre.q_end; t.s_off = c.pre.s_off; t.s_end = c.pre.s_end; t.score = c.pre.score;
        if (!tree.contains(t, b.min_diag_separation) && c.has_start &&
            !(dp_filter && hsp_fails_identity_or_length(b, c.num_ident, c.align_length))) {
            // Blast_HSPUpdateWithTraceback (:156-175) + Blast_HSPAdjustSubjectOffset (core/blast_hits.c:1168-1179)
            h.alive = true;
            h.score = c.res.score;
            h.q_off = c.res.query_start; h.q_end = c.res.query_stop;
            h.s_off = c.res.subject_start + c.s_shift; h.s_end = c.res.subject_stop + c.s_shift;
            h.q_gapped_start = c.q_start; h.s_gapped_start = c.s_start + c.s_shift;
            h.esp.assign(c.ops, c.ops + c.res.esp_n);
            IntervalTree::Item nt{t.q_strand_start, h.q_off, h.q_end, h.s_off, h.s_end, h.score};
            tree.add(nt, true);
        }
        arr.push_back(std::move(h));
    }
    // Blast_HSPListPurgeNullHSPs
    {
        size_t o = 0;
        for (size_t i = 0; i < arr.size(); i++) if (arr[i].alive) { if (o != i) arr[o] = std::move(arr[i]); o++; }
        arr.resize(o);
    }
    if (arr.empty()) return;
    // Blast_HSPListPurgeHSPsWithCommonEndpoints(eBlastTypeBlastn, list, FALSE): works on an array of pointers
    std::vector<int32_t> p(arr.size());
    for (size_t i = 0; i < p.size(); i++) p[i] = (int32_t)i;
    int32_t hsp_count = (int32_t)p.size();
    auto by_offset = [&](int32_t a, int32_t c) {      // s_QueryOffsetCompareHSPs
        const TbHsp &x = arr[(size_t)a], &y = arr[(size_t)c];
        if (x.context != y.context) return x.context < y.context;
        if (x.q_off != y.q_off) return x.q_off < y.q_off;
        if (x.s_off != y.s_off) return x.s_off < y.s_off;
        if (x.score != y.score) return x.score > y.score;
        if (x.q_end != y.q_end) return x.q_end > y.q_end;
        if (x.s_end != y.s_end) return x.s_end > y.s_end;
        return false;
    };
    auto by_end = [&](int32_t a, int32_t c) {         // s_QueryEndCompareHSPs
        const TbHsp &x = arr[(size_t)a], &y = arr[(size_t)c];
        if (x.context != y.context) return x.context < y.context;
        if (x.q_end != y.q_end) return x.q_end < y.q_end;
        if (x.s_end != y.s_end) return x.s_end < y.s_end;
        if (x.score != y.score) return x.score > y.score;
        if (x.q_off != y.q_off) return x.q_off > y.q_off;
        if (x.s_off != y.s_off) return x.s_off > y.s_off;
        return false;
    };
    std::stable_sort(p.begin(), p.begin() + hsp_count, by_offset);
    int32_t i = 0;
    while (i < hsp_count) {
        int32_t j = 1;
        while (i + j < hsp_count && p[(size_t)i] >= 0 && p[(size_t)(i + j)] >= 0 &&
               arr[(size_t)p[(size_t)i]].context == arr[(size_t)p[(size_t)(i + j)]].context &&
               arr[(size_t)p[(size_t)i]].q_off == arr[(size_t)p[(size_t)(i + j)]].q_off &&
               arr[(size_t)p[(size_t)i]].s_off == arr[(size_t)p[(size_t)(i + j)]].s_off) {
            hsp_count--;
            int32_t hp = p[(size_t)(i + j)];
            TbHsp &hsp = arr[(size_t)hp];
            const TbHsp &keep = arr[(size_t)p[(size_t)i]];
            if (hsp.q_end > keep.q_end) { cut_off_edit_script(hsp, keep.q_end, keep.s_end, true); hsp.was_cut = true; }
            else { hsp.alive = false; hp = -1; }
            for (int32_t k = i + j; k < hsp_count; k++) p[(size_t)k] = p[(size_t)(k + 1)];
            p[(size_t)hsp_count] = hp;
        }
        i += j;
    }
    std::stable_sort(p.begin(), p.begin() + hsp_count, by_end);
    i = 0;
    while (i < hsp_count) {
        int32_t j = 1;
        while (i + j < hsp_count && p[(size_t)i] >= 0 && p[(size_t)(i + j)] >= 0 &&
               arr[(size_t)p[(size_t)i]].context == arr[(size_t)p[(size_t)(i + j)]].context &&
               arr[(size_t)p[(size_t)i]].q_end == arr[(size_t)p[(size_t)(i + j)]].q_end &&
               arr[(size_t)p[(size_t)i]].s_end == arr[(size_t)p[(size_t)(i + j)]].s_end) {
            hsp_count--;
            int32_t hp = p[(size_t)(i + j)];
            TbHsp &hsp = arr[(size_t)hp];
            const TbHsp &keep = arr[(size_t)p[(size_t)i]];
            if (hsp.q_off < keep.q_off) { cut_off_edit_script(hsp, keep.q_off, keep.s_off, false); hsp.was_cut = true; }
            else { hsp.alive = false; hp = -1; }
            for (int32_t k = i + j; k < hsp_count; k++) p[(size_t)k] = p[(size_t)(k + 1)];
            p[(size_t)hsp_count] = hp;
        }
        i += j;
    }
    // materialise hsp_array in its new order (NULL entries stay as dead records)
    std::vector<TbHsp> out(p.size());
    for (size_t k = 0; k < p.size(); k++) {
        if (p[k] >= 0) out[k] = std::move(arr[(size_t)p[k]]);
        else { out[k] = TbHsp{}; out[k].alive = false; }
    }
    arr.swap(out);
    extra_start = (size_t)hsp_count;
}

void traceback_list_stage2(const BnQueryBatch &b, int32_t subject_length, std::vector<TbHsp> &arr)
{
    size_t o = 0;
    for (size_t i = 0; i < arr.size(); i++) if (arr[i].alive) { if (o != i) arr[o] = std::move(arr[i]); o++; }
    arr.resize(o);
    if (arr.empty()) return;
    std::stable_sort(arr.begin(), arr.end(), tb_score_less);
    if (arr.size() > 1) {   // containment among the final alignments (:741-760); a single HSP meets an empty tree
        IntervalTree tree(0, b.concat_len + 1, 0, subject_length + 1, arr.size());
        for (TbHsp &h : arr) {
            IntervalTree::Item t{strand_offset(b, h.context), h.q_off, h.q_end, h.s_off, h.s_end, h.score};
            if (tree.contains(t, b.min_diag_separation)) h.alive = false;
            else tree.add(t, true);
        }
        o = 0;
        for (size_t i = 0; i < arr.size(); i++) if (arr[i].alive) { if (o != i) arr[o] = std::move(arr[i]); o++; }
        arr.resize(o);
    }
    // s_HSPListPostTracebackUpdate
    if (b.round_down) {
        for (TbHsp &h : arr) h.score &= ~1;
        std::stable_sort(arr.begin(), arr.end(), tb_score_less);
    }
    o = 0;
    for (size_t i = 0; i < arr.size(); i++) {
        TbHsp &h = arr[i];
        const BnContext &c = b.contexts[h.context];
        h.evalue = (double)c.eff_searchsp * std::exp((double)(-c.gap_lambda * h.score) + c.gap_logK);
        if (h.evalue > b.evalue_cutoff) continue;
        // Blast_HSPListGetBitScores (core/blast_hits.c): (Lambda * score - logK) / ln 2
        // the reference is built with -ffast-math (core/Makefile.blast.lib:19): the division by the constant NCBIMATH_LN2
        // is a multiplication by its reciprocal there
        h.bit_score = (h.score * c.gap_lambda - c.gap_logK) * (1.0 / 0.69314718055994530941723212145818);
        if (o != i) arr[o] = std::move(arr[i]);
        o++;
    }
    arr.resize(o);
}

bool traceback_list_before(const std::vector<TbHsp> &a, const std::vector<TbHsp> &c)
{
    double ea = a[0].evalue, ec = c[0].evalue;            // best_evalue = minimum over the list
    for (const TbHsp &h : a) ea = std::min(ea, h.evalue);
    for (const TbHsp &h : c) ec = std::min(ec, h.evalue);
    const int f = fuzzy_cmp(ea, ec);
    if (f != 0) return f < 0;
    if (a[0].score != c[0].score) return a[0].score > c[0].score;
    return a[0].oid > c[0].oid;                           // BLAST_CMP(h2->oid, h1->oid)
}

}  // namespace bn
